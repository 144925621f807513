// accuracy probe for the approximate fp64/fp32 reciprocal and rsqrt used in fvm_math.h (run under gpurun)
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double rcp_it(double x, int it) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    for (int i = 0; i < it; i++) { double e = fma(-x, y, 1.0); y = fma(y, e, y); } return y; }
__device__ double rsq_it(double x, int it) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    for (int i = 0; i < it; i++) { double e = fma(-x * y, y, 1.0); y = fma(0.5 * y, e, y); } return y; }
__global__ void k(double* out) {
    double mr[4] = {0, 0, 0, 0}, ms[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < 200000; i += blockDim.x) {
        double x = exp((i / 200000.0 - 0.5) * 60.0) * (1.0 + 1e-3 * (i % 97));
        for (int it = 0; it < 4; it++) {
            double e = fabs(rcp_it(x, it) * x - 1.0); if (e > mr[it]) mr[it] = e;
            double s = rsq_it(x, it); double e2 = fabs(s * s * x - 1.0); if (e2 > ms[it]) ms[it] = e2;
        }
    }
    for (int it = 0; it < 4; it++) { out[(threadIdx.x * 8) + it] = mr[it]; out[threadIdx.x * 8 + 4 + it] = ms[it]; }
}
__global__ void kf(float* out) {
    float mr = 0, ms = 0, md = 0;
    for (int i = threadIdx.x; i < 200000; i += blockDim.x) {
        float x = expf((i / 200000.0f - 0.5f) * 40.0f) * (1.0f + 1e-3f * (i % 97));
        float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        float s; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
        float e = fabsf((float)((double)y * (double)x - 1.0)); if (e > mr) mr = e;
        float e2 = fabsf((float)((double)s * (double)s * (double)x - 1.0)); if (e2 > ms) ms = e2;
        float q = __fdividef(3.0f, x); float e3 = fabsf((float)((double)q * (double)x / 3.0 - 1.0)); if (e3 > md) md = e3;
    }
    out[threadIdx.x * 3] = mr; out[threadIdx.x * 3 + 1] = ms; out[threadIdx.x * 3 + 2] = md;
}
int main() {
    double* d; cudaMalloc(&d, 256 * 8 * 8); k<<<1, 256>>>(d); double h[256 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int it = 0; it < 4; it++) { double a = 0, b = 0; for (int t = 0; t < 256; t++) { a = fmax(a, h[t * 8 + it]); b = fmax(b, h[t * 8 + 4 + it]); }
        printf("fp64 newton iters %d: rcp max rel err %.3e   rsqrt(y*y*x-1) %.3e\n", it, a, b); }
    float* f; cudaMalloc(&f, 256 * 3 * 4); kf<<<1, 256>>>(f); float hf[768]; cudaMemcpy(hf, f, sizeof(hf), cudaMemcpyDeviceToHost);
    float a = 0, b = 0, c = 0; for (int t = 0; t < 256; t++) { a = fmaxf(a, hf[t * 3]); b = fmaxf(b, hf[t * 3 + 1]); c = fmaxf(c, hf[t * 3 + 2]); }
    printf("fp32 approx: rcp %.3e rsqrt(y*y*x-1) %.3e fdividef %.3e\n", a, b, c);
    return 0;
}
