/* The reference always links adFVM/cpp/scaling.cpp (adjoint artificial viscosity, out of scope,
 * SURVEY §2 row 24), which declares LAPACK symmetric eigen-solvers. No LAPACK in this image; the hot
 * path never calls them, so they abort if reached. */
#include <cstdio>
#include <cstdlib>
#define STUB(name) extern "C" void name(...) { std::fprintf(stderr, "lapack stub: " #name " called\n"); std::abort(); }
STUB(dsyev_) STUB(dsygv_) STUB(ssyev_) STUB(ssygv_)
