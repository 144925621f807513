/* TEST INFRASTRUCTURE ONLY. The reference links adFVM/cpp/scaling.cpp, whose adjoint-viscosity path calls LAPACK's
 * symmetric eigen-solvers (scaling.cpp:66-81). The image has no system LAPACK; scipy's wheel bundles OpenBLAS + LAPACK
 * with `scipy_`-prefixed symbols. These forwarders resolve them at first use from the library named by ADFVM_LAPACK_SO,
 * so the UNMODIFIED reference can compute its adjoint viscosity here (oracle/ref_harness/gen_viscosity.py). */
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
static void* sym(const char* name) {
    static void* h = nullptr;
    if (!h) {
        const char* so = std::getenv("ADFVM_LAPACK_SO");
        if (!so || !(h = dlopen(so, RTLD_NOW | RTLD_GLOBAL))) { std::fprintf(stderr, "lapack_fwd: cannot open ADFVM_LAPACK_SO (%s)\n", so ? so : "unset"); std::abort(); }
    }
    void* p = dlsym(h, name);
    if (!p) { std::fprintf(stderr, "lapack_fwd: %s not found\n", name); std::abort(); }
    return p;
}
extern "C" {
void dsyev_(char* jobz, char* uplo, int* n, double* a, int* lda, double* w, double* work, int* lwork, int* info) {
    typedef void (*fn)(char*, char*, int*, double*, int*, double*, double*, int*, int*, size_t, size_t);
    static fn f = (fn)sym("scipy_dsyev_"); f(jobz, uplo, n, a, lda, w, work, lwork, info, 1, 1);
}
void ssyev_(char* jobz, char* uplo, int* n, float* a, int* lda, float* w, float* work, int* lwork, int* info) {
    typedef void (*fn)(char*, char*, int*, float*, int*, float*, float*, int*, int*, size_t, size_t);
    static fn f = (fn)sym("scipy_ssyev_"); f(jobz, uplo, n, a, lda, w, work, lwork, info, 1, 1);
}
void dsygv_(int* itype, char* jobz, char* uplo, int* n, double* a, int* lda, double* b, int* ldb, double* w, double* work, int* lwork, int* info) {
    typedef void (*fn)(int*, char*, char*, int*, double*, int*, double*, int*, double*, double*, int*, int*, size_t, size_t);
    static fn f = (fn)sym("scipy_dsygv_"); f(itype, jobz, uplo, n, a, lda, b, ldb, w, work, lwork, info, 1, 1);
}
void ssygv_(int* itype, char* jobz, char* uplo, int* n, float* a, int* lda, float* b, int* ldb, float* w, float* work, int* lwork, int* info) {
    typedef void (*fn)(int*, char*, char*, int*, float*, int*, float*, int*, float*, float*, int*, int*, size_t, size_t);
    static fn f = (fn)sym("scipy_ssygv_"); f(itype, jobz, uplo, n, a, lda, b, ldb, w, work, lwork, info, 1, 1);
}
}
