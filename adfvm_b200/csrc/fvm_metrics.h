// Mesh metric build on the device: what the reference computes once per mesh on the CPU in
// adFVM/cpp/cmesh.cpp:65-270 (+ the ghost-cell centres of adFVM/mesh.py:758-819) - face normals / centres / areas,
// cell centres / volumes, deltas, interpolation and reconstruction weights. Inputs and outputs are the reference's
// AoS arrays ([n][d] row-major, point / face / cell numbering of the case); one-off kernels, one thread per face or
// cell. The integer connectivity (cellFaces: owned faces ascending, then neighbour faces, cmesh.cpp:9-19) is a sort
// and stays with the host layer.
#pragma once
#include "fvm_math.h"

namespace fvm {

template <typename R> FVM_HD void cross3(const R* a, const R* b, R* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename R> FVM_HD R norm3(const R* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// cmesh.cpp:69-121: unit normal from the face's first three points ((a-b) x (b-c)), centre and area by a triangle fan
// about the mean of the four vertices
template <typename R> struct FaceGeomBody {
    static constexpr const char* kName = "mesh_face_geom";
    const R* points; const int* faces;          // [nP][3], [nF][4]
    R *normals, *centres, *areas;               // [nF][3], [nF][3], [nF]
    FVM_HD void operator()(int f) const {
        R P[4][3];
        for (int j = 0; j < 4; j++) for (int k = 0; k < 3; k++) P[j][k] = points[(long)faces[(long)f * 4 + j] * 3 + k];
        R ab[3], bc[3], n[3];
        for (int k = 0; k < 3; k++) { ab[k] = P[0][k] - P[1][k]; bc[k] = P[1][k] - P[2][k]; }
        cross3(ab, bc, n);
        const R nn = norm3(n);
        for (int k = 0; k < 3; k++) normals[(long)f * 3 + k] = n[k] / nn;
        R fc0[3];
        for (int k = 0; k < 3; k++) fc0[k] = (((P[0][k] + P[1][k]) + P[2][k]) + P[3][k]) / R(4);
        R area = R(0), sumC[3] = {R(0), R(0), R(0)};
        for (int j = 0; j < 4; j++) {
            const R* p0 = P[j]; const R* p1 = P[(j + 1) % 4];
            R e0[3], e1[3], N[3];
            for (int k = 0; k < 3; k++) { e0[k] = p1[k] - p0[k]; e1[k] = fc0[k] - p0[k]; }
            cross3(e0, e1, N);
            const R Ns = norm3(N);
            area = area + Ns / R(2);
            for (int k = 0; k < 3; k++) sumC[k] = sumC[k] + (Ns * ((p0[k] + p1[k] + fc0[k]) / R(3))) / R(2);
        }
        areas[f] = area;
        for (int k = 0; k < 3; k++) centres[(long)f * 3 + k] = sumC[k] / area;
    }
};

// cmesh.cpp:126-162: cell centre and volume from six pyramids about the mean of the face centres
template <typename R> struct CellGeomBody {
    static constexpr const char* kName = "mesh_cell_geom";
    const int* cellFaces;                       // [nIC][6]
    const R *normals, *faceCentres, *areas;
    R *cellCentres, *volumes;                   // [nCells][3] (first nIC rows), [nIC]
    FVM_HD void operator()(int c) const {
        R cc0[3] = {R(0), R(0), R(0)};
        for (int j = 0; j < 6; j++) { const long f = cellFaces[(long)c * 6 + j]; for (int k = 0; k < 3; k++) cc0[k] = cc0[k] + faceCentres[f * 3 + k]; }
        for (int k = 0; k < 3; k++) cc0[k] = cc0[k] / R(6);
        R vol = R(0), sumCC[3] = {R(0), R(0), R(0)};
        for (int j = 0; j < 6; j++) {
            const long f = cellFaces[(long)c * 6 + j];
            R an[3], h[3];
            for (int k = 0; k < 3; k++) { an[k] = areas[f] * normals[f * 3 + k]; h[k] = cc0[k] - faceCentres[f * 3 + k]; }
            R v = (an[0] * h[0] + an[1] * h[1]) + an[2] * h[2];
            v = fabs(v / R(3));
            vol = vol + v;
            for (int k = 0; k < 3; k++) sumCC[k] = sumCC[k] + v * (R(3) / R(4) * faceCentres[f * 3 + k] + R(1) / R(4) * cc0[k]);
        }
        volumes[c] = vol;
        for (int k = 0; k < 3; k++) cellCentres[(long)c * 3 + k] = sumCC[k] / vol;
    }
};

// adFVM/mesh.py:758-819: centre of the ghost cell of boundary face b. kind 0: the face centre; 1 (cyclic): the partner
// face's owner centre shifted by the offset between the two patches' first faces; 2 (processor): supplied by the caller
struct MetricPatch { int startFace, nFaces, kind, nbrStartFace; };
template <typename R> struct GhostCentreBody {
    static constexpr const char* kName = "mesh_ghost_centre";
    const MetricPatch* patches; const unsigned char* bpatch;    // patch of every boundary face
    const int* owner; const R* faceCentres; const R* remote;     // remote: [nBoundaryFaces][3] or NULL
    int nIF, nIC; R* cellCentres;
    FVM_HD void operator()(int b) const {
        const MetricPatch& P = patches[bpatch[b]];
        const long f = nIF + b, g = (long)nIC + b;
        if (P.kind == 1) {
            const long i = f - P.startFace, s = P.startFace, ns = P.nbrStartFace, o = owner[ns + i];
            for (int k = 0; k < 3; k++) cellCentres[g * 3 + k] = (faceCentres[s * 3 + k] - faceCentres[ns * 3 + k]) + cellCentres[o * 3 + k];
        } else if (P.kind == 2) {
            for (int k = 0; k < 3; k++) cellCentres[g * 3 + k] = remote[(long)b * 3 + k];
        } else {
            for (int k = 0; k < 3; k++) cellCentres[g * 3 + k] = faceCentres[f * 3 + k];
        }
    }
};

// cmesh.cpp:168-233: deltas, deltasUnit, interpolation weight (of the OWNER value), linear / quadratic weights of the
// second-order reconstruction
template <typename R> struct FaceWeightBody {
    static constexpr const char* kName = "mesh_face_weights";
    const int *owner, *neighbour;               // neighbour: all faces (ghost cells for boundary faces)
    const R *cellCentres, *faceCentres, *normals;
    R *deltas, *deltasUnit, *weights, *linW, *quadW;     // [nF], [nF][3], [nF], [nF][2], [nF][2][3]
    FVM_HD void operator()(int f) const {
        const long o = owner[f], n = neighbour[f];
        R delta[3], nFv[3], pFv[3], nrm[3];
        for (int k = 0; k < 3; k++) {
            delta[k] = cellCentres[o * 3 + k] - cellCentres[n * 3 + k];
            nFv[k] = faceCentres[(long)f * 3 + k] - cellCentres[n * 3 + k];
            pFv[k] = faceCentres[(long)f * 3 + k] - cellCentres[o * 3 + k];
            nrm[k] = normals[(long)f * 3 + k];
        }
        const R d = norm3(delta);
        deltas[f] = d;
        for (int k = 0; k < 3; k++) deltasUnit[(long)f * 3 + k] = -delta[k] / d;
        const R nD = fabs((nFv[0] * nrm[0] + nFv[1] * nrm[1]) + nFv[2] * nrm[2]);
        const R pD = fabs((pFv[0] * nrm[0] + pFv[1] * nrm[1]) + pFv[2] * nrm[2]);
        weights[f] = nD / (nD + pD);
        R w1 = ((-delta[0] * pFv[0]) + (-delta[1] * pFv[1])) + (-delta[2] * pFv[2]);
        R w2 = ((delta[0] * nFv[0]) + (delta[1] * nFv[1])) + (delta[2] * nFv[2]);
        const R d2 = (delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2];
        w1 = w1 / d2; w2 = w2 / d2;
        linW[(long)f * 2] = w1 / R(3); linW[(long)f * 2 + 1] = w2 / R(3);
        for (int k = 0; k < 3; k++) {
            quadW[(long)f * 6 + k] = R(2) / R(3) * pFv[k] + R(1) / R(3) * (pFv[k] + w1 * delta[k]);
            quadW[(long)f * 6 + 3 + k] = R(2) / R(3) * nFv[k] + R(1) / R(3) * (nFv[k] - w2 * delta[k]);
        }
    }
};

}  // namespace fvm
