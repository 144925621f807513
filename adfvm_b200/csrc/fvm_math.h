// Per-face / per-cell arithmetic of the adFVM residual and its hand-derived reverse mode.
// Scalar, branch-light, __host__ __device__ so that the very same source is compiled by nvcc for the
// sm_100a kernels and by g++ for the CPU-side test simulator (tests/hostsim) that checks it against
// the oracle without a GPU.
//
// Reference formulas restated here (paths relative to /root/reference):
//   conservative/primitive  adFVM/density.py:162-181
//   secondOrder             adFVM/interp.py:20-28
//   eulerRoe                adFVM/riemann.py:21-70
//   eulerLaxFriedrichs      adFVM/riemann.py:6-18
//   viscousFlux             adFVM/density.py:192-222
//   flux / boundaryFlux     adFVM/density.py:253-331
//   AD rules                adpy/adpy/scalar.py:157-320  (abs: x<0 ? -g : g ; switch: no grad to cond ;
//                                                        max-reduce: no grad)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FVM_HD __host__ __device__ __forceinline__
#define FVM_RESTRICT __restrict__
#else
#define FVM_HD inline
#define FVM_RESTRICT __restrict__
#endif

namespace fvm {

enum MuLaw { MU_CONSTANT = 0, MU_SUTHERLAND = 1 };
enum Riemann { RIEMANN_ROE = 0, RIEMANN_LAXFRIEDRICHS = 1 };
// how a face's flux is evaluated (adFVM/density.py:386-403)
enum FaceKind { FACE_COUPLED = 0,        // internal, cyclic, processor: reconstruct both sides + riemannSolver
                FACE_CHARACTERISTIC = 1, // left reconstruction, right = ghost state, boundaryRiemannSolver
                FACE_BOUNDARY = 2 };     // analytic Euler flux of the ghost state

template <typename R> struct Phys {
    R gamma, Cp, Pr, Cv, small;   // small = config.SMALL (1e-30 fp64 / 1e-9 fp32, adFVM/config.py:171-178)
    R mu_value;
    R gm1, iCv, iCvgm1, g_gm1, CpPr;   // gamma-1, 1/Cv, 1/(Cv (gamma-1)), gamma/(gamma-1), Cp/Pr (set_physics)
    int mu_law, riemann, boundary_riemann;
};

// Reciprocal and reciprocal square root without the IEEE-division slow path. The flux kernels are bound by
// instruction issue, and an fp64 `a/b` costs ~18 SASS instructions + a divergent fix-up branch; MUFU.RCP64H /
// MUFU.RSQ64H seeds (rel. error 1e-6, measured) + two Newton steps reach 1.1e-16 / 3.7e-16 (tools/rcp_test.cu),
// i.e. the last bit or two, in 5 / 9 instructions. fp32 uses the MUFU result directly (1e-7, within the stated
// fp32 tolerance). Arguments on this path are physical quantities (rho, T, p, a^2, V, ...): positive, normal.
// The CPU build (tests/hostsim) uses plain division.
FVM_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0); y = fma(y, e, y);
    e = fma(-x, y, 1.0); y = fma(y, e, y);
    return y;
#else
    return 1.0 / x;
#endif
}
FVM_HD float rcp(float x) {
#if defined(__CUDA_ARCH__)
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return 1.0f / x;
#endif
}
FVM_HD double rsqrt_fast(double x) {
#if defined(__CUDA_ARCH__)
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0); y = fma(0.5 * y, e, y);
    e = fma(-x * y, y, 1.0); y = fma(0.5 * y, e, y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}
FVM_HD float rsqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
    return 1.0f / sqrtf(x);
#endif
}

template <typename R> struct Prim  { R U[3], T, p; };
template <typename R> struct Grad  { R U[9], T[3], p[3]; };   // U[3*i+j] = dU_i/dx_j
template <typename R> struct Geom  { R area, n[3], idelta, d[3], lw[2], qw[2][3]; };   // idelta = 1/deltas
template <typename R> struct Flux5 { R rho, rhoU[3], rhoE; };

template <typename R> FVM_HD R sgn_ref(R x) { return x < R(0) ? R(-1) : R(1); }   // scalar.py:209-212
template <typename R> FVM_HD R dot3(const R* a, const R* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename R> FVM_HD void zero(Prim<R>& q) { q.U[0] = q.U[1] = q.U[2] = q.T = q.p = R(0); }
template <typename R> FVM_HD void zero(Grad<R>& g) {
    for (int i = 0; i < 9; i++) g.U[i] = R(0);
    for (int i = 0; i < 3; i++) g.T[i] = g.p[i] = R(0);
}

template <typename R> FVM_HD R viscosity(const Phys<R>& ph, R T) {
    if (ph.mu_law == MU_SUTHERLAND) return R(1.4792e-06) * T * (T * rsqrt_fast(T)) * rcp(T + R(116.));   // density.py:27
    return ph.mu_value;
}
template <typename R> FVM_HD R viscosity_dT(const Phys<R>& ph, R T, R mu) {
    if (ph.mu_law == MU_SUTHERLAND) return mu * (R(1.5) * rcp(T) - rcp(T + R(116.)));
    return R(0);
}

// ---------------------------------------------------------------- cell: conservative -> primitive
template <typename R> FVM_HD void primitive(const Phys<R>& ph, R rho, const R* rhoU, R rhoE, Prim<R>& q) {
    R ir = rcp(rho);
    q.U[0] = rhoU[0] * ir; q.U[1] = rhoU[1] * ir; q.U[2] = rhoU[2] * ir;
    R e = rhoE * ir - R(0.5) * dot3(q.U, q.U);
    q.p = ph.gm1 * rho * e;
    q.T = e * ph.iCv;
}
// reverse: given qb (adjoint of U,T,p) accumulate into (rhob, rhoUb, rhoEb)
template <typename R> FVM_HD void primitive_vjp(const Phys<R>& ph, R rho, const R* rhoU, R rhoE, const Prim<R>& qb,
                                                R& rhob, R* rhoUb, R& rhoEb) {
    R ir = rcp(rho);
    R U[3] = {rhoU[0] * ir, rhoU[1] * ir, rhoU[2] * ir};
    R E = rhoE * ir;
    R e = E - R(0.5) * dot3(U, U);
    R eb = qb.p * ph.gm1 * rho + qb.T * ph.iCv;
    R rb = qb.p * ph.gm1 * e;
    R Eb = eb;
    R Ub[3];
    for (int i = 0; i < 3; i++) Ub[i] = qb.U[i] - eb * U[i];
    rhoEb += Eb * ir;
    rb += -Eb * E * ir;
    for (int i = 0; i < 3; i++) { rhoUb[i] += Ub[i] * ir; rb += -Ub[i] * U[i] * ir; }
    rhob += rb;
}

// ---------------------------------------------------------------- face: primitive -> conservative
template <typename R> struct Cons { R rho, rhoU[3], rhoE, iT, irho; };   // iT = 1/T, irho = 1/rho (shared reciprocals)
template <typename R> FVM_HD void conservative(const Phys<R>& ph, const Prim<R>& q, Cons<R>& w) {
    R e = ph.Cv * q.T;
    w.iT = rcp(q.T);
    w.rho = q.p * w.iT * ph.iCvgm1;                  // p / (Cv T (gamma-1))
    w.irho = rcp(w.rho);
    w.rhoE = w.rho * (e + R(0.5) * dot3(q.U, q.U));
    for (int i = 0; i < 3; i++) w.rhoU[i] = q.U[i] * w.rho;
}
template <typename R> FVM_HD void conservative_vjp(const Phys<R>& ph, const Prim<R>& q, const Cons<R>& w,
                                                   const Cons<R>& wb, Prim<R>& qb) {
    R e = ph.Cv * q.T;
    R rb = wb.rho + wb.rhoE * (e + R(0.5) * dot3(q.U, q.U)) + dot3(wb.rhoU, q.U);
    R eb = wb.rhoE * w.rho;
    for (int i = 0; i < 3; i++) qb.U[i] += wb.rhoE * w.rho * q.U[i] + wb.rhoU[i] * w.rho;
    // rho = p/(e (g-1))
    qb.p += rb * w.iT * ph.iCvgm1;
    eb += -rb * w.rho * w.iT * ph.iCv;
    qb.T += eb * ph.Cv;
}

// ---------------------------------------------------------------- reconstruction (interp.py:20-28)
// phiF = phiC + (phiD - phiC)*lw + qw . grad(phiC)
template <typename R> FVM_HD void reconstruct(const Prim<R>& C, const Prim<R>& D, const Grad<R>& gC, R lw, const R* qw, Prim<R>& F) {
    for (int i = 0; i < 3; i++) F.U[i] = C.U[i] + (D.U[i] - C.U[i]) * lw + (qw[0] * gC.U[3 * i] + qw[1] * gC.U[3 * i + 1] + qw[2] * gC.U[3 * i + 2]);
    F.T = C.T + (D.T - C.T) * lw + dot3(qw, gC.T);
    F.p = C.p + (D.p - C.p) * lw + dot3(qw, gC.p);
}
template <typename R> FVM_HD void reconstruct_vjp(const Prim<R>& Fb, R lw, const R* qw, Prim<R>& Cb, Prim<R>& Db, Grad<R>& gCb) {
    R a = R(1) - lw;
    for (int i = 0; i < 3; i++) {
        Cb.U[i] += Fb.U[i] * a; Db.U[i] += Fb.U[i] * lw;
        for (int j = 0; j < 3; j++) gCb.U[3 * i + j] += qw[j] * Fb.U[i];
    }
    Cb.T += Fb.T * a; Db.T += Fb.T * lw;
    Cb.p += Fb.p * a; Db.p += Fb.p * lw;
    for (int j = 0; j < 3; j++) { gCb.T[j] += qw[j] * Fb.T; gCb.p[j] += qw[j] * Fb.p; }
}

// ---------------------------------------------------------------- Roe flux (riemann.py:21-70)
// Forward intermediates needed by the reverse sweep are kept in this struct.
template <typename R> struct RoeTmp {
    R unL, unR, mL, mR, hL, hR, sL, sR, rsL, rsR, idv, Ut[3], ht, qt, a2, a, ia, unt, dr, dU[3], dE;
    R l1, l2, l3, e1, e2, cL, cR, icL, icR, eps, ie, l1p, l2p, l3p; bool c1, c2, c3;
    R b1, b2, b3, b4, b5, b6, b7;
};

template <typename R>
FVM_HD void roe_forward(const Phys<R>& ph, const Prim<R>& L, const Prim<R>& Rr, const Cons<R>& wL, const Cons<R>& wR,
                        const R* N, Flux5<R>& F, RoeTmp<R>& t) {
    const R g = ph.gamma, gm1 = ph.gm1;
    const R rL = wL.rho, rR = wR.rho, irL = wL.irho, irR = wR.irho;
    t.unL = dot3(L.U, N); t.unR = dot3(Rr.U, N);
    t.mL = rL * t.unL; t.mR = rR * t.unR;
    t.hL = ph.g_gm1 * L.p * irL + R(0.5) * dot3(L.U, L.U);
    t.hR = ph.g_gm1 * Rr.p * irR + R(0.5) * dot3(Rr.U, Rr.U);
    F.rho = R(0.5) * (t.mL + t.mR);
    for (int i = 0; i < 3; i++) F.rhoU[i] = R(0.5) * (t.mL * L.U[i] + t.mR * Rr.U[i] + (L.p + Rr.p) * N[i]);
    F.rhoE = R(0.5) * (t.mL * t.hL + t.mR * t.hR);
    t.rsL = rsqrt_fast(rL); t.rsR = rsqrt_fast(rR);
    t.sL = rL * t.rsL; t.sR = rR * t.rsR;
    t.idv = rcp(t.sL + t.sR);
    for (int i = 0; i < 3; i++) t.Ut[i] = (L.U[i] * t.sL + Rr.U[i] * t.sR) * t.idv;
    t.ht = (t.hL * t.sL + t.hR * t.sR) * t.idv;
    t.qt = R(0.5) * dot3(t.Ut, t.Ut);
    t.a2 = gm1 * (t.ht - t.qt);
    t.ia = rsqrt_fast(t.a2);
    t.a = t.a2 * t.ia;
    t.unt = dot3(t.Ut, N);
    t.dr = rR - rL;
    for (int i = 0; i < 3; i++) t.dU[i] = rR * Rr.U[i] - rL * L.U[i];
    t.dE = (t.hR * rR - Rr.p) - (t.hL * rL - L.p);
    t.l1 = fabs(t.unt); t.l2 = fabs(t.unt + t.a); t.l3 = fabs(t.unt - t.a);
    t.e1 = t.mL * irL - t.mR * irR;
    const R xL = g * L.p * irL, xR = g * Rr.p * irR;
    t.icL = rsqrt_fast(xL); t.icR = rsqrt_fast(xR);
    t.cL = xL * t.icL; t.cR = xR * t.icR;
    t.e2 = t.cL - t.cR;
    R eps = R(0.5) * fabs(t.e1) + R(0.5) * fabs(t.e2);
    t.eps = (eps < R(0)) ? eps - ph.small : eps + ph.small;            // Tensor.stabilise
    t.c1 = t.l1 < R(2) * t.eps; t.c2 = t.l2 < R(2) * t.eps; t.c3 = t.l3 < R(2) * t.eps;
    t.ie = (t.c1 || t.c2 || t.c3) ? rcp(t.eps) : R(0);                 // Harten fix is rarely active
    t.l1p = t.c1 ? R(.25) * t.l1 * t.l1 * t.ie + t.eps : t.l1;
    t.l2p = t.c2 ? R(.25) * t.l2 * t.l2 * t.ie + t.eps : t.l2;
    t.l3p = t.c3 ? R(.25) * t.l3 * t.l3 * t.ie + t.eps : t.l3;
    t.b1 = R(0.5) * (t.l2p + t.l3p);
    t.b2 = R(0.5) * (t.l2p - t.l3p);
    t.b3 = t.b1 - t.l1p;
    t.b4 = gm1 * (t.qt * t.dr - dot3(t.Ut, t.dU) + t.dE);
    t.b5 = t.unt * t.dr - dot3(t.dU, N);
    t.b6 = t.b3 * t.b4 * (t.ia * t.ia) - t.b2 * t.b5 * t.ia;
    t.b7 = t.b3 * t.b5 - t.b2 * t.b4 * t.ia;
    F.rho -= R(0.5) * (t.l1p * t.dr + t.b6);
    for (int i = 0; i < 3; i++) F.rhoU[i] -= R(0.5) * (t.l1p * t.dU[i] + t.Ut[i] * t.b6 - t.b7 * N[i]);
    F.rhoE -= R(0.5) * (t.l1p * t.dE + t.ht * t.b6 - t.unt * t.b7);
}

// reverse: accumulates into Lb, Rb (U, p only; T untouched) and rLb, rRb (adjoint of rho_L, rho_R)
template <typename R>
FVM_HD void roe_reverse(const Phys<R>& ph, const Prim<R>& L, const Prim<R>& Rr, const Cons<R>& wL, const Cons<R>& wR,
                        const R* N, const RoeTmp<R>& t, const Flux5<R>& Fb, Prim<R>& Lb, Prim<R>& Rb, R& rLb_out, R& rRb_out,
                        R* Nb = nullptr) {       // Nb (optional): accumulates the adjoint of the face normal (mesh sensitivities)
    const R g = ph.gamma, gm1 = ph.gm1;
    const R rL = wL.rho, rR = wR.rho, irL = wL.irho, irR = wR.irho;
    const R h = R(0.5);
    R rLb = R(0), rRb = R(0), pLb = R(0), pRb = R(0), ULb[3] = {0, 0, 0}, URb[3] = {0, 0, 0};
    // ---- dissipation
    R FUbdU = dot3(Fb.rhoU, t.dU), FUbUt = dot3(Fb.rhoU, t.Ut), FUbN = dot3(Fb.rhoU, N);
    R l1pb = -h * (Fb.rho * t.dr + FUbdU + Fb.rhoE * t.dE);
    R drb = -h * Fb.rho * t.l1p;
    R dUb[3], Utb[3];
    for (int i = 0; i < 3; i++) { dUb[i] = -h * t.l1p * Fb.rhoU[i]; Utb[i] = -h * t.b6 * Fb.rhoU[i]; }
    R dEb = -h * t.l1p * Fb.rhoE;
    R b6b = -h * (Fb.rho + FUbUt + Fb.rhoE * t.ht);
    R b7b = h * FUbN + h * Fb.rhoE * t.unt;
    R htb = -h * Fb.rhoE * t.b6;
    R untb = h * Fb.rhoE * t.b7;
    // b7 = b3*b5 - b2*b4/a ; b6 = b3*b4/a2 - b2*b5/a
    R ia = t.ia, ia2 = t.ia * t.ia;
    R b3b = b7b * t.b5 + b6b * t.b4 * ia2;
    R b5b = b7b * t.b3 - b6b * t.b2 * ia;
    R b2b = -b7b * t.b4 * ia - b6b * t.b5 * ia;
    R b4b = -b7b * t.b2 * ia + b6b * t.b3 * ia2;
    R ab = (b7b * t.b2 * t.b4 + b6b * t.b2 * t.b5) * ia * ia;
    R a2b = -b6b * t.b3 * t.b4 * ia2 * ia2;
    // b5 = unt*dr - dU.N
    untb += b5b * t.dr; drb += b5b * t.unt;
    for (int i = 0; i < 3; i++) dUb[i] += -b5b * N[i];
    // b4 = gm1*(qt*dr - Ut.dU + dE)
    R tb = gm1 * b4b;
    R qtb = tb * t.dr; drb += tb * t.qt; dEb += tb;
    for (int i = 0; i < 3; i++) { Utb[i] += -tb * t.dU[i]; dUb[i] += -tb * t.Ut[i]; }
    // b3 = b1 - l1p ; b1 = .5(l2p+l3p) ; b2 = .5(l2p-l3p)
    l1pb += -b3b;
    R l2pb = h * (b3b + b2b), l3pb = h * (b3b - b2b);
    // entropy fix
    R epsb = R(0), l1b, l2b, l3b;
    const R ie = t.ie;
    if (t.c1) { l1b = l1pb * h * t.l1 * ie; epsb += l1pb * (R(1) - R(.25) * t.l1 * t.l1 * ie * ie); } else l1b = l1pb;
    if (t.c2) { l2b = l2pb * h * t.l2 * ie; epsb += l2pb * (R(1) - R(.25) * t.l2 * t.l2 * ie * ie); } else l2b = l2pb;
    if (t.c3) { l3b = l3pb * h * t.l3 * ie; epsb += l3pb * (R(1) - R(.25) * t.l3 * t.l3 * ie * ie); } else l3b = l3pb;
    // eps = .5|e1| + .5|e2| (+- small)
    R e1b = h * epsb * sgn_ref(t.e1), e2b = h * epsb * sgn_ref(t.e2);
    // cL = sqrt(g pL/rL)
    pLb += e2b * g * h * t.icL * irL; rLb += -e2b * t.cL * h * irL;
    pRb += -e2b * g * h * t.icR * irR; rRb += e2b * t.cR * h * irR;
    // e1 = mL/rL - mR/rR
    R mLb = e1b * irL, mRb = -e1b * irR;
    rLb += -e1b * t.mL * irL * irL; rRb += e1b * t.mR * irR * irR;
    // l1 = |unt|, l2 = |unt+a|, l3 = |unt-a|
    R s1 = sgn_ref(t.unt), s2 = sgn_ref(t.unt + t.a), s3 = sgn_ref(t.unt - t.a);
    untb += l1b * s1 + l2b * s2 + l3b * s3;
    ab += l2b * s2 - l3b * s3;
    // dE = (hR*rR - pR) - (hL*rL - pL)
    R hRb = dEb * rR, hLb = -dEb * rL;
    rRb += dEb * t.hR; pRb += -dEb; rLb += -dEb * t.hL; pLb += dEb;
    // dU = rR*UR - rL*UL ; dr = rR - rL
    rRb += dot3(dUb, Rr.U) + drb; rLb += -dot3(dUb, L.U) - drb;
    for (int i = 0; i < 3; i++) { URb[i] += dUb[i] * rR; ULb[i] += -dUb[i] * rL; }
    // unt = Ut.N ; a = sqrt(a2) ; a2 = gm1 (ht - qt) ; qt = .5 Ut.Ut
    a2b += ab * h * ia;
    htb += gm1 * a2b; qtb += -gm1 * a2b;
    for (int i = 0; i < 3; i++) Utb[i] += untb * N[i] + qtb * t.Ut[i];
    // ht = (hL sL + hR sR)/dv ; Ut = (UL sL + UR sR)/dv
    R idv = t.idv;
    hLb += htb * t.sL * idv; hRb += htb * t.sR * idv;
    R sLb = htb * t.hL * idv + dot3(Utb, L.U) * idv;
    R sRb = htb * t.hR * idv + dot3(Utb, Rr.U) * idv;
    R dvb = -(htb * t.ht + dot3(Utb, t.Ut)) * idv;
    for (int i = 0; i < 3; i++) { ULb[i] += Utb[i] * t.sL * idv; URb[i] += Utb[i] * t.sR * idv; }
    sLb += dvb; sRb += dvb;
    rLb += sLb * h * t.rsL; rRb += sRb * h * t.rsR;
    // ---- central part
    mLb += h * Fb.rhoE * t.hL + h * dot3(Fb.rhoU, L.U) + h * Fb.rho;
    mRb += h * Fb.rhoE * t.hR + h * dot3(Fb.rhoU, Rr.U) + h * Fb.rho;
    hLb += h * Fb.rhoE * t.mL; hRb += h * Fb.rhoE * t.mR;
    for (int i = 0; i < 3; i++) { ULb[i] += h * t.mL * Fb.rhoU[i]; URb[i] += h * t.mR * Fb.rhoU[i]; }
    pLb += h * FUbN; pRb += h * FUbN;
    // hL = g pL/(gm1 rL) + .5 UL.UL
    pLb += hLb * ph.g_gm1 * irL; rLb += -hLb * ph.g_gm1 * L.p * irL * irL;
    pRb += hRb * ph.g_gm1 * irR; rRb += -hRb * ph.g_gm1 * Rr.p * irR * irR;
    for (int i = 0; i < 3; i++) { ULb[i] += hLb * L.U[i]; URb[i] += hRb * Rr.U[i]; }
    // mL = rL*unL ; unL = UL.N
    rLb += mLb * t.unL; rRb += mRb * t.unR;
    for (int i = 0; i < 3; i++) { ULb[i] += mLb * rL * N[i]; URb[i] += mRb * rR * N[i]; }
    for (int i = 0; i < 3; i++) { Lb.U[i] += ULb[i]; Rb.U[i] += URb[i]; }
    Lb.p += pLb; Rb.p += pRb;
    rLb_out += rLb; rRb_out += rRb;
    if (Nb) {   // unL = UL.N, unR = UR.N, central (pL+pR) N, unt = Ut.N, b5 = unt dr - dU.N, dissipation + b7 N
        for (int i = 0; i < 3; i++)
            Nb[i] += mLb * rL * L.U[i] + mRb * rR * Rr.U[i] + h * (L.p + Rr.p) * Fb.rhoU[i] + untb * t.Ut[i] - b5b * t.dU[i] + h * t.b7 * Fb.rhoU[i];
    }
}

// ---------------------------------------------------------------- Lax-Friedrichs (riemann.py:6-18)
template <typename R>
FVM_HD void lf_forward(const Phys<R>& ph, const Prim<R>& L, const Prim<R>& Rr, const Cons<R>& wL, const Cons<R>& wR,
                       const R* N, Flux5<R>& F) {
    const R g = ph.gamma, h = R(0.5);
    R unL = dot3(L.U, N), unR = dot3(Rr.U, N);
    const R xL = g * L.p * wL.irho, xR = g * Rr.p * wR.irho;
    R cL = xL * rsqrt_fast(xL), cR = xR * rsqrt_fast(xR);
    R x1 = fabs(unL) + cL, x2 = fabs(unR) + cR;
    R aF = (x1 > x2) ? x1 : x2;
    F.rho = h * (wL.rho * unL + wR.rho * unR) - h * aF * (wR.rho - wL.rho);
    for (int i = 0; i < 3; i++)
        F.rhoU[i] = h * (wL.rhoU[i] * unL + wR.rhoU[i] * unR + (L.p + Rr.p) * N[i]) - h * aF * (wR.rhoU[i] - wL.rhoU[i]);
    F.rhoE = h * ((wL.rhoE + L.p) * unL + (wR.rhoE + Rr.p) * unR) - h * aF * (wR.rhoE - wL.rhoE);
}
// reverse: accumulates into Lb/Rb (U,p) and wLb/wRb (all conservative components)
template <typename R>
FVM_HD void lf_reverse(const Phys<R>& ph, const Prim<R>& L, const Prim<R>& Rr, const Cons<R>& wL, const Cons<R>& wR,
                       const R* N, const Flux5<R>& Fb, Prim<R>& Lb, Prim<R>& Rb, Cons<R>& wLb, Cons<R>& wRb, R* Nb = nullptr) {
    const R g = ph.gamma, h = R(0.5);
    R unL = dot3(L.U, N), unR = dot3(Rr.U, N);
    const R xL = g * L.p * wL.irho, xR = g * Rr.p * wR.irho;
    const R icL = rsqrt_fast(xL), icR = rsqrt_fast(xR);
    R cL = xL * icL, cR = xR * icR;
    R x1 = fabs(unL) + cL, x2 = fabs(unR) + cR;
    bool first = x1 > x2;
    R aF = first ? x1 : x2;
    R FUbN = dot3(Fb.rhoU, N);
    R aFb = -h * (Fb.rho * (wR.rho - wL.rho) + Fb.rhoE * (wR.rhoE - wL.rhoE));
    for (int i = 0; i < 3; i++) aFb += -h * Fb.rhoU[i] * (wR.rhoU[i] - wL.rhoU[i]);
    R unLb = h * (Fb.rho * wL.rho + dot3(Fb.rhoU, wL.rhoU) + Fb.rhoE * (wL.rhoE + L.p));
    R unRb = h * (Fb.rho * wR.rho + dot3(Fb.rhoU, wR.rhoU) + Fb.rhoE * (wR.rhoE + Rr.p));
    wLb.rho += h * Fb.rho * (unL + aF); wRb.rho += h * Fb.rho * (unR - aF);
    for (int i = 0; i < 3; i++) { wLb.rhoU[i] += h * Fb.rhoU[i] * (unL + aF); wRb.rhoU[i] += h * Fb.rhoU[i] * (unR - aF); }
    wLb.rhoE += h * Fb.rhoE * (unL + aF); wRb.rhoE += h * Fb.rhoE * (unR - aF);
    Lb.p += h * FUbN + h * Fb.rhoE * unL; Rb.p += h * FUbN + h * Fb.rhoE * unR;
    if (first) {
        unLb += aFb * sgn_ref(unL);
        Lb.p += aFb * g * h * icL * wL.irho; wLb.rho += -aFb * cL * h * wL.irho;
    } else {
        unRb += aFb * sgn_ref(unR);
        Rb.p += aFb * g * h * icR * wR.irho; wRb.rho += -aFb * cR * h * wR.irho;
    }
    for (int i = 0; i < 3; i++) { Lb.U[i] += unLb * N[i]; Rb.U[i] += unRb * N[i]; }
    if (Nb) for (int i = 0; i < 3; i++) Nb[i] += unLb * L.U[i] + unRb * Rr.U[i] + h * (L.p + Rr.p) * Fb.rhoU[i];
}

// ---------------------------------------------------------------- viscous flux (density.py:192-222)
// adds -sigma to F.rhoU and -(q + sigma.UF) to F.rhoE.  TL,TR,UL,UR are CELL values (snGrad), TF/UF face values,
// gTF[3], gUF[9] the face gradients before the "charles" normal-derivative correction.
template <typename R>
FVM_HD void viscous_forward(const Phys<R>& ph, const Geom<R>& gm, R TL, R TR, const R* UL, const R* UR, R TF, const R* UF,
                            const R* gTF, const R* gUF, Flux5<R>& F) {
    R mu = viscosity(ph, TF);
    R kappa = mu * ph.CpPr;
    const R* D = gm.d; const R* N = gm.n;
    R idel = gm.idelta;
    R gT[3], gU[9];
    R snT = (TR - TL) * idel, gTD = dot3(gTF, D);
    for (int j = 0; j < 3; j++) gT[j] = gTF[j] + snT * D[j] - gTD * D[j];
    for (int i = 0; i < 3; i++) {
        R snU = (UR[i] - UL[i]) * idel;
        R gUD = gUF[3 * i] * D[0] + gUF[3 * i + 1] * D[1] + gUF[3 * i + 2] * D[2];
        for (int j = 0; j < 3; j++) gU[3 * i + j] = gUF[3 * i + j] + snU * D[j] - gUD * D[j];
    }
    R qF = kappa * dot3(gT, N);
    R tr = gU[0] + gU[4] + gU[8];
    R sig[3];
    for (int i = 0; i < 3; i++) {
        R tmp2 = (gU[3 * i] + gU[i]) * N[0] + (gU[3 * i + 1] + gU[3 + i]) * N[1] + (gU[3 * i + 2] + gU[6 + i]) * N[2];
        sig[i] = mu * (tmp2 - R(2. / 3) * tr * N[i]);
        F.rhoU[i] += -sig[i];
    }
    F.rhoE += -(qF + dot3(sig, UF));
}
// reverse: accumulates TLb,TRb,ULb,URb (cell values), TFb, UFb, gTFb[3], gUFb[9]
template <typename R>
FVM_HD void viscous_reverse(const Phys<R>& ph, const Geom<R>& gm, R TL, R TR, const R* UL, const R* UR, R TF, const R* UF,
                            const R* gTF, const R* gUF, const Flux5<R>& Fb,
                            R& TLb, R& TRb, R* ULb, R* URb, R& TFb, R* UFb, R* gTFb, R* gUFb,
                            R* Db = nullptr, R* Nb = nullptr, R* idelb = nullptr) {   // optional: adjoints of deltasUnit, normal, 1/delta
    R mu = viscosity(ph, TF);
    R CpPr = ph.CpPr;
    R kappa = mu * CpPr;
    const R* D = gm.d; const R* N = gm.n;
    R idel = gm.idelta;
    // recompute forward
    R gT[3], gU[9];
    R snT = (TR - TL) * idel, gTD = dot3(gTF, D);
    for (int j = 0; j < 3; j++) gT[j] = gTF[j] + snT * D[j] - gTD * D[j];
    for (int i = 0; i < 3; i++) {
        R snU = (UR[i] - UL[i]) * idel;
        R gUD = gUF[3 * i] * D[0] + gUF[3 * i + 1] * D[1] + gUF[3 * i + 2] * D[2];
        for (int j = 0; j < 3; j++) gU[3 * i + j] = gUF[3 * i + j] + snU * D[j] - gUD * D[j];
    }
    R gTN = dot3(gT, N);
    R tr = gU[0] + gU[4] + gU[8];
    R tmp2[3], sig[3];
    for (int i = 0; i < 3; i++) {
        tmp2[i] = (gU[3 * i] + gU[i]) * N[0] + (gU[3 * i + 1] + gU[3 + i]) * N[1] + (gU[3 * i + 2] + gU[6 + i]) * N[2];
        sig[i] = mu * (tmp2[i] - R(2. / 3) * tr * N[i]);
    }
    // reverse
    R sigb[3], tmp2b[3];
    R mub = R(0), trb = R(0);
    for (int i = 0; i < 3; i++) {
        sigb[i] = -Fb.rhoU[i] - Fb.rhoE * UF[i];
        UFb[i] += -Fb.rhoE * sig[i];
        mub += sigb[i] * (tmp2[i] - R(2. / 3) * tr * N[i]);
        tmp2b[i] = mu * sigb[i];
        trb += -R(2. / 3) * mu * sigb[i] * N[i];
    }
    R qFb = -Fb.rhoE;
    mub += qFb * gTN * CpPr;
    R gTb[3];
    for (int j = 0; j < 3; j++) gTb[j] = qFb * kappa * N[j];
    R gUb[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) gUb[3 * i + j] = tmp2b[i] * N[j] + tmp2b[j] * N[i] + (i == j ? trb : R(0));
    for (int i = 0; i < 3; i++) {
        R gD = gUb[3 * i] * D[0] + gUb[3 * i + 1] * D[1] + gUb[3 * i + 2] * D[2];
        for (int j = 0; j < 3; j++) gUFb[3 * i + j] += gUb[3 * i + j] - gD * D[j];
        URb[i] += gD * idel; ULb[i] -= gD * idel;
        if (Db) {   // gU[3i+j] = gUF[3i+j] + (snU_i - gUF_i.D) D[j]
            const R snU = (UR[i] - UL[i]) * idel;
            const R gUD = gUF[3 * i] * D[0] + gUF[3 * i + 1] * D[1] + gUF[3 * i + 2] * D[2];
            for (int j = 0; j < 3; j++) Db[j] += gUb[3 * i + j] * (snU - gUD) - gD * gUF[3 * i + j];
            *idelb += gD * (UR[i] - UL[i]);
        }
    }
    R gTbD = dot3(gTb, D);
    for (int j = 0; j < 3; j++) gTFb[j] += gTb[j] - gTbD * D[j];
    TRb += gTbD * idel; TLb -= gTbD * idel;
    if (Db) {
        for (int j = 0; j < 3; j++) Db[j] += gTb[j] * (snT - gTD) - gTbD * gTF[j];
        *idelb += gTbD * (TR - TL);
        // qF = kappa gT.N ; sig_i = mu (tmp2_i - 2/3 tr N_i), tmp2_i = sum_j (gU[3i+j] + gU[3j+i]) N_j
        for (int j = 0; j < 3; j++) {
            R v = qFb * kappa * gT[j] - R(2. / 3) * mu * tr * sigb[j];
            for (int i = 0; i < 3; i++) v += mu * sigb[i] * (gU[3 * i + j] + gU[3 * j + i]);
            Nb[j] += v;
        }
    }
    TFb += mub * viscosity_dT(ph, TF, mu);
}

// ---------------------------------------------------------------- full face flux, three kinds
// Returns the flux per unit area (rho, rhoU, rhoE) and the CFL wave speed (|UF.n| + sqrt(Cp TF (g-1))).
template <typename R>
FVM_HD void face_flux(const Phys<R>& ph, int kind, const Geom<R>& gm, const Prim<R>& qL, const Grad<R>& gL,
                      const Prim<R>& qR, const Grad<R>& gR, Flux5<R>& F, R& wave) {
    const R h = R(0.5);
    if (kind == FACE_BOUNDARY) {                      // density.py:307-331 (+ getFlux :184-190)
        Cons<R> w; conservative(ph, qR, w);
        R un = dot3(qR.U, gm.n);
        F.rho = w.rho * un;
        for (int i = 0; i < 3; i++) F.rhoU[i] = w.rhoU[i] * un + qR.p * gm.n[i];
        F.rhoE = (w.rhoE + qR.p) * un;
        viscous_forward(ph, gm, qL.T, qR.T, qL.U, qR.U, qR.T, qR.U, gR.T, gR.U, F);
        { const R x = ph.Cp * qR.T * ph.gm1; wave = fabs(un) + x * rsqrt_fast(x); }
        return;
    }
    Prim<R> LF, RF;
    reconstruct(qL, qR, gL, gm.lw[0], gm.qw[0], LF);
    Cons<R> wL, wR;
    conservative(ph, LF, wL);
    R TF, UF[3], gTF[3], gUF[9];
    int solver;
    if (kind == FACE_CHARACTERISTIC) {                // density.py:266-274
        RF = qR;
        solver = ph.boundary_riemann;
        TF = qR.T;
        for (int i = 0; i < 3; i++) { UF[i] = qR.U[i]; gTF[i] = gR.T[i]; }
        for (int i = 0; i < 9; i++) gUF[i] = gR.U[i];
    } else {                                          // density.py:275-291
        reconstruct(qR, qL, gR, gm.lw[1], gm.qw[1], RF);
        solver = ph.riemann;
        TF = h * (LF.T + RF.T);
        for (int i = 0; i < 3; i++) { UF[i] = h * (LF.U[i] + RF.U[i]); gTF[i] = h * (gL.T[i] + gR.T[i]); }
        for (int i = 0; i < 9; i++) gUF[i] = h * (gL.U[i] + gR.U[i]);
    }
    conservative(ph, RF, wR);
    if (solver == RIEMANN_ROE) { RoeTmp<R> t; roe_forward(ph, LF, RF, wL, wR, gm.n, F, t); }
    else lf_forward(ph, LF, RF, wL, wR, gm.n, F);
    viscous_forward(ph, gm, qL.T, qR.T, qL.U, qR.U, TF, UF, gTF, gUF, F);
    { const R x = ph.Cp * TF * ph.gm1; wave = fabs(dot3(UF, gm.n)) + x * rsqrt_fast(x); }
}

// Reverse of face_flux with respect to the FACE METRICS (mesh sensitivities, parameters = 'mesh' of the reference,
// apps/adjoint.py:105-107): accumulates into gb.n, gb.d, gb.idelta, gb.lw, gb.qw for a given flux adjoint Fb (per unit
// area; the scatter weights A/V are handled by the caller). Same three face kinds as face_flux.
template <typename R>
FVM_HD void face_flux_metric_vjp(const Phys<R>& ph, int kind, const Geom<R>& gm, const Prim<R>& qL, const Grad<R>& gL,
                                 const Prim<R>& qR, const Grad<R>& gR, const Flux5<R>& Fb, Geom<R>& gb) {
    const R h = R(0.5);
    Prim<R> dumL, dumR; zero(dumL); zero(dumR);
    R TFb = R(0), UFb[3] = {0, 0, 0}, gTFb[3] = {0, 0, 0}, gUFb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (kind == FACE_BOUNDARY) {
        Cons<R> w; conservative(ph, qR, w);
        const R unb = Fb.rho * w.rho + dot3(Fb.rhoU, w.rhoU) + Fb.rhoE * (w.rhoE + qR.p);
        for (int i = 0; i < 3; i++) gb.n[i] += unb * qR.U[i] + qR.p * Fb.rhoU[i];
        viscous_reverse(ph, gm, qL.T, qR.T, qL.U, qR.U, qR.T, qR.U, gR.T, gR.U, Fb,
                        dumL.T, dumR.T, dumL.U, dumR.U, TFb, UFb, gTFb, gUFb, gb.d, gb.n, &gb.idelta);
        return;
    }
    Prim<R> LF, RF;
    reconstruct(qL, qR, gL, gm.lw[0], gm.qw[0], LF);
    Cons<R> wL, wR;
    conservative(ph, LF, wL);
    Prim<R> LFb, RFb; zero(LFb); zero(RFb);
    Cons<R> wLb, wRb;
    wLb.rho = wRb.rho = wLb.rhoE = wRb.rhoE = R(0);
    for (int i = 0; i < 3; i++) wLb.rhoU[i] = wRb.rhoU[i] = R(0);
    const bool ch = (kind == FACE_CHARACTERISTIC);
    R TF, UF[3], gTF[3], gUF[9];
    if (ch) {
        RF = qR; TF = qR.T;
        for (int i = 0; i < 3; i++) { UF[i] = qR.U[i]; gTF[i] = gR.T[i]; }
        for (int i = 0; i < 9; i++) gUF[i] = gR.U[i];
    } else {
        reconstruct(qR, qL, gR, gm.lw[1], gm.qw[1], RF);
        TF = h * (LF.T + RF.T);
        for (int i = 0; i < 3; i++) { UF[i] = h * (LF.U[i] + RF.U[i]); gTF[i] = h * (gL.T[i] + gR.T[i]); }
        for (int i = 0; i < 9; i++) gUF[i] = h * (gL.U[i] + gR.U[i]);
    }
    conservative(ph, RF, wR);
    viscous_reverse(ph, gm, qL.T, qR.T, qL.U, qR.U, TF, UF, gTF, gUF, Fb,
                    dumL.T, dumR.T, dumL.U, dumR.U, TFb, UFb, gTFb, gUFb, gb.d, gb.n, &gb.idelta);
    const int solver = ch ? ph.boundary_riemann : ph.riemann;
    if (solver == RIEMANN_ROE) {
        Flux5<R> F; RoeTmp<R> t;
        roe_forward(ph, LF, RF, wL, wR, gm.n, F, t);
        roe_reverse(ph, LF, RF, wL, wR, gm.n, t, Fb, LFb, RFb, wLb.rho, wRb.rho, gb.n);
    } else {
        lf_reverse(ph, LF, RF, wL, wR, gm.n, Fb, LFb, RFb, wLb, wRb, gb.n);
    }
    conservative_vjp(ph, LF, wL, wLb, LFb);
    conservative_vjp(ph, RF, wR, wRb, RFb);
    if (!ch) {
        LFb.T += h * TFb; RFb.T += h * TFb;
        for (int i = 0; i < 3; i++) { LFb.U[i] += h * UFb[i]; RFb.U[i] += h * UFb[i]; }
    }
    // reconstruction (interp.py:20-28): phiF = phiC + (phiD - phiC) lw + qw . grad(phiC)
    {
        R s = (qR.T - qL.T) * LFb.T + (qR.p - qL.p) * LFb.p;
        for (int i = 0; i < 3; i++) s += (qR.U[i] - qL.U[i]) * LFb.U[i];
        gb.lw[0] += s;
        for (int j = 0; j < 3; j++) {
            R v = gL.T[j] * LFb.T + gL.p[j] * LFb.p;
            for (int i = 0; i < 3; i++) v += gL.U[3 * i + j] * LFb.U[i];
            gb.qw[0][j] += v;
        }
    }
    if (!ch) {
        R s = (qL.T - qR.T) * RFb.T + (qL.p - qR.p) * RFb.p;
        for (int i = 0; i < 3; i++) s += (qL.U[i] - qR.U[i]) * RFb.U[i];
        gb.lw[1] += s;
        for (int j = 0; j < 3; j++) {
            R v = gR.T[j] * RFb.T + gR.p[j] * RFb.p;
            for (int i = 0; i < 3; i++) v += gR.U[3 * i + j] * RFb.U[i];
            gb.qw[1][j] += v;
        }
    }
}

// Compact form of the reverse of a COUPLED face (both sides reconstructed): everything the 40 input adjoints are
// made of. LF/RF: adjoints of the reconstructed states (face means of T,U folded in), dU/dT: the normal-derivative
// (snGrad) share of the owner-side cell values (the neighbour side gets the opposite), hgU/hgT: half the adjoint
// of the face gradient (both sides get it).
template <typename R> struct FaceAdj { Prim<R> LF, RF; R dU[3], dT, hgU[9], hgT[3]; };

template <typename R>
FVM_HD void face_flux_vjp_coupled(const Phys<R>& ph, const Geom<R>& gm, const Prim<R>& qL, const Grad<R>& gL,
                                  const Prim<R>& qR, const Grad<R>& gR, const Flux5<R>& Fb, FaceAdj<R>& c) {
    const R h = R(0.5);
    // ---- recompute forward
    Prim<R> LF, RF;
    reconstruct(qL, qR, gL, gm.lw[0], gm.qw[0], LF);
    reconstruct(qR, qL, gR, gm.lw[1], gm.qw[1], RF);
    Cons<R> wL, wR;
    conservative(ph, LF, wL);
    conservative(ph, RF, wR);
    R UF[3], gTF[3], gUF[9];
    const R TF = h * (LF.T + RF.T);
    for (int i = 0; i < 3; i++) { UF[i] = h * (LF.U[i] + RF.U[i]); gTF[i] = h * (gL.T[i] + gR.T[i]); }
    for (int i = 0; i < 9; i++) gUF[i] = h * (gL.U[i] + gR.U[i]);
    // ---- reverse
    zero(c.LF); zero(c.RF);
    Cons<R> wLb, wRb;
    wLb.rho = wRb.rho = wLb.rhoE = wRb.rhoE = R(0);
    for (int i = 0; i < 3; i++) wLb.rhoU[i] = wRb.rhoU[i] = R(0);
    R TFb = R(0), UFb[3] = {0, 0, 0}, gTFb[3] = {0, 0, 0}, gUFb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    R TLb = R(0), TRb = R(0), ULb[3] = {0, 0, 0}, URb[3] = {0, 0, 0};
    viscous_reverse(ph, gm, qL.T, qR.T, qL.U, qR.U, TF, UF, gTF, gUF, Fb, TLb, TRb, ULb, URb, TFb, UFb, gTFb, gUFb);
    if (ph.riemann == RIEMANN_ROE) {
        Flux5<R> F; RoeTmp<R> t;
        roe_forward(ph, LF, RF, wL, wR, gm.n, F, t);
        roe_reverse(ph, LF, RF, wL, wR, gm.n, t, Fb, c.LF, c.RF, wLb.rho, wRb.rho);
    } else {
        lf_reverse(ph, LF, RF, wL, wR, gm.n, Fb, c.LF, c.RF, wLb, wRb);
    }
    conservative_vjp(ph, LF, wL, wLb, c.LF);
    conservative_vjp(ph, RF, wR, wRb, c.RF);
    c.LF.T += h * TFb; c.RF.T += h * TFb;
    for (int i = 0; i < 3; i++) { c.LF.U[i] += h * UFb[i]; c.RF.U[i] += h * UFb[i]; c.dU[i] = ULb[i]; c.hgT[i] = h * gTFb[i]; }
    c.dT = TLb;
    for (int i = 0; i < 9; i++) c.hgU[i] = h * gUFb[i];
}
// component k (0-4 U,T,p; 5-13 gradU; 14-16 gradT; 17-19 gradp) of the owner-side / neighbour-side input adjoints
// (reverse of the two reconstructions, interp.py:20-28, + the shares above)
template <typename R> FVM_HD R face_adj_owner(const FaceAdj<R>& c, const Geom<R>& gm, int k) {
    const R a = R(1) - gm.lw[0], b = gm.lw[1];
    if (k < 3) return c.dU[k] + c.LF.U[k] * a + c.RF.U[k] * b;
    if (k == 3) return c.dT + c.LF.T * a + c.RF.T * b;
    if (k == 4) return c.LF.p * a + c.RF.p * b;
    if (k < 14) return c.hgU[k - 5] + gm.qw[0][(k - 5) % 3] * c.LF.U[(k - 5) / 3];
    if (k < 17) return c.hgT[k - 14] + gm.qw[0][k - 14] * c.LF.T;
    return gm.qw[0][k - 17] * c.LF.p;
}
template <typename R> FVM_HD R face_adj_neighbour(const FaceAdj<R>& c, const Geom<R>& gm, int k) {
    const R a = R(1) - gm.lw[1], b = gm.lw[0];
    if (k < 3) return -c.dU[k] + c.RF.U[k] * a + c.LF.U[k] * b;
    if (k == 3) return -c.dT + c.RF.T * a + c.LF.T * b;
    if (k == 4) return c.RF.p * a + c.LF.p * b;
    if (k < 14) return c.hgU[k - 5] + gm.qw[1][(k - 5) % 3] * c.RF.U[(k - 5) / 3];
    if (k < 17) return c.hgT[k - 14] + gm.qw[1][k - 14] * c.RF.T;
    return gm.qw[1][k - 17] * c.RF.p;
}
template <typename R> FVM_HD void add20(Prim<R>& q, Grad<R>& g, const R* v) {
    q.U[0] += v[0]; q.U[1] += v[1]; q.U[2] += v[2]; q.T += v[3]; q.p += v[4];
    for (int k = 0; k < 9; k++) g.U[k] += v[5 + k];
    for (int k = 0; k < 3; k++) { g.T[k] += v[14 + k]; g.p[k] += v[17 + k]; }
}

// Reverse of face_flux w.r.t. (qL, gL, qR, gR) for a given flux adjoint Fb (the wave speed only feeds a
// max-reduction, which carries no gradient). ACCUMULATES into qLb, gLb, qRb, gRb.
template <typename R>
FVM_HD void face_flux_vjp(const Phys<R>& ph, int kind, const Geom<R>& gm, const Prim<R>& qL, const Grad<R>& gL,
                          const Prim<R>& qR, const Grad<R>& gR, const Flux5<R>& Fb,
                          Prim<R>& qLb, Grad<R>& gLb, Prim<R>& qRb, Grad<R>& gRb) {
    const R h = R(0.5);
    if (kind == FACE_BOUNDARY) {
        Cons<R> w; conservative(ph, qR, w);
        R un = dot3(qR.U, gm.n);
        R FUbN = dot3(Fb.rhoU, gm.n);
        Cons<R> wb;
        wb.rho = Fb.rho * un; wb.rhoE = Fb.rhoE * un;
        for (int i = 0; i < 3; i++) wb.rhoU[i] = Fb.rhoU[i] * un;
        R unb = Fb.rho * w.rho + dot3(Fb.rhoU, w.rhoU) + Fb.rhoE * (w.rhoE + qR.p);
        qRb.p += FUbN + Fb.rhoE * un;
        for (int i = 0; i < 3; i++) qRb.U[i] += unb * gm.n[i];
        conservative_vjp(ph, qR, w, wb, qRb);
        R TFb = R(0), UFb[3] = {0, 0, 0};
        viscous_reverse(ph, gm, qL.T, qR.T, qL.U, qR.U, qR.T, qR.U, gR.T, gR.U, Fb,
                        qLb.T, qRb.T, qLb.U, qRb.U, TFb, UFb, gRb.T, gRb.U);
        qRb.T += TFb;
        for (int i = 0; i < 3; i++) qRb.U[i] += UFb[i];
        return;
    }
    if (kind == FACE_COUPLED) {
        FaceAdj<R> c; face_flux_vjp_coupled(ph, gm, qL, gL, qR, gR, Fb, c);
        R vo[20], vn[20];
        for (int k = 0; k < 20; k++) { vo[k] = face_adj_owner(c, gm, k); vn[k] = face_adj_neighbour(c, gm, k); }
        add20(qLb, gLb, vo); add20(qRb, gRb, vn);
        return;
    }
    // ---- characteristic face (density.py:266-274): left reconstruction, right = ghost state
    Prim<R> LF;
    reconstruct(qL, qR, gL, gm.lw[0], gm.qw[0], LF);
    Cons<R> wL, wR;
    conservative(ph, LF, wL);
    conservative(ph, qR, wR);
    Prim<R> LFb, RFb; zero(LFb); zero(RFb);
    Cons<R> wLb, wRb;
    wLb.rho = wRb.rho = wLb.rhoE = wRb.rhoE = R(0);
    for (int i = 0; i < 3; i++) wLb.rhoU[i] = wRb.rhoU[i] = R(0);
    R TFb = R(0), UFb[3] = {0, 0, 0};
    viscous_reverse(ph, gm, qL.T, qR.T, qL.U, qR.U, qR.T, qR.U, gR.T, gR.U, Fb,
                    qLb.T, qRb.T, qLb.U, qRb.U, TFb, UFb, gRb.T, gRb.U);
    if (ph.boundary_riemann == RIEMANN_ROE) {
        Flux5<R> F; RoeTmp<R> t;
        roe_forward(ph, LF, qR, wL, wR, gm.n, F, t);
        roe_reverse(ph, LF, qR, wL, wR, gm.n, t, Fb, LFb, RFb, wLb.rho, wRb.rho);
    } else {
        lf_reverse(ph, LF, qR, wL, wR, gm.n, Fb, LFb, RFb, wLb, wRb);
    }
    conservative_vjp(ph, LF, wL, wLb, LFb);
    conservative_vjp(ph, qR, wR, wRb, RFb);
    qRb.T += RFb.T + TFb; qRb.p += RFb.p;
    for (int i = 0; i < 3; i++) qRb.U[i] += RFb.U[i] + UFb[i];
    reconstruct_vjp(LFb, gm.lw[0], gm.qw[0], qLb, qRb, gLb);
}

}  // namespace fvm
