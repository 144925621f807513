// CUDA executor: wraps the per-element bodies of fvm_bodies.h in sm_100a kernels on one stream.
// One thread per cell / face; 128-thread CTAs (the flux bodies are register-heavy fp64 code: 128 threads keep
// >= 2 CTAs resident per SM within the 64K-register file), grids sized by the element count so
// the 148 SMs stay covered by several waves; reductions are two-pass fixed-order trees (deterministic).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include <map>

namespace fvm {

#define FVM_CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #x); } while (0)

enum { kBlock = 128, kRedBlock = 256, kRedMaxBlocks = 148 * 8 };

// resident CTAs per SM a body asks for (register cap = 64K / (128 * n)); bodies without kMinBlocks get 1
template <class B, class = void> struct MinBlocksOf { static constexpr int value = 1; };
template <class B> struct MinBlocksOf<B, decltype((void)B::kMinBlocks)> { static constexpr int value = B::kMinBlocks; };

template <class Body> __global__ void __launch_bounds__(kBlock, MinBlocksOf<Body>::value) k_run(const Body b, int first, int n) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) b(first + i);
}
template <class Body> __global__ void __launch_bounds__(kBlock) k_run_discard(const Body b, int n) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) (void)b(i);
}

template <class Body> __global__ void __launch_bounds__(Body::kThreads, Body::kMinBlocks) k_tile(const Body b, int first) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    b.device_tile(first + blockIdx.x, tile_smem);
}
template <typename R> struct BufferBody {      // reduce an array already in device memory
    static constexpr const char* kName = "reduce_buffer";
    const R* p;
    __device__ R operator()(int i) const { return p[i]; }
};

struct OpSum { template <typename R> __device__ static R id() { return R(0); } template <typename R> __device__ static R op(R a, R b) { return a + b; } };
struct OpMax { template <typename R> __device__ static R id() { return R(-1e30); } template <typename R> __device__ static R op(R a, R b) { return a > b ? a : b; } };

// pass 1: block b owns elements b*kRedBlock + t + k*gridDim*kRedBlock (fixed assignment -> fixed order)
template <typename R, class Op, class Body> __global__ void __launch_bounds__(kRedBlock) k_reduce1(const Body b, int n, R* partial) {
    __shared__ R sm[kRedBlock];
    R acc = Op::template id<R>();
    for (long i = (long)blockIdx.x * kRedBlock + threadIdx.x; i < n; i += (long)gridDim.x * kRedBlock)
        acc = Op::op(acc, b((int)i));
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kRedBlock / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] = Op::op(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
template <typename R, class Op> __global__ void __launch_bounds__(kRedBlock) k_reduce2(const R* partial, int nb, R* out) {
    __shared__ R sm[kRedBlock];
    R acc = Op::template id<R>();
    for (int i = threadIdx.x; i < nb; i += kRedBlock) acc = Op::op(acc, partial[i]);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kRedBlock / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] = Op::op(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

// five sums at once (the conjugate-gradient dot products of fvm_viscosity.h): same fixed-order two-pass tree, V = V5<R>
template <class V, class Body> __global__ void __launch_bounds__(kRedBlock) k_reduce1_v5(const Body b, int n, V* partial) {
    __shared__ V sm[kRedBlock];
    V acc;
    for (int k = 0; k < 5; k++) acc.v[k] = 0;
    for (long i = (long)blockIdx.x * kRedBlock + threadIdx.x; i < n; i += (long)gridDim.x * kRedBlock) {
        const V x = b((int)i);
        for (int k = 0; k < 5; k++) acc.v[k] += x.v[k];
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kRedBlock / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) for (int k = 0; k < 5; k++) sm[threadIdx.x].v[k] += sm[threadIdx.x + s].v[k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
template <class V, typename R> __global__ void __launch_bounds__(kRedBlock) k_reduce2_v5(const V* partial, int nb, R* out) {
    __shared__ V sm[kRedBlock];
    V acc;
    for (int k = 0; k < 5; k++) acc.v[k] = 0;
    for (int i = threadIdx.x; i < nb; i += kRedBlock) for (int k = 0; k < 5; k++) acc.v[k] += partial[i].v[k];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kRedBlock / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) for (int k = 0; k < 5; k++) sm[threadIdx.x].v[k] += sm[threadIdx.x + s].v[k];
        __syncthreads();
    }
    if (threadIdx.x == 0) for (int k = 0; k < 5; k++) out[k] = sm[0].v[k];
}

struct CudaExec {
    cudaStream_t stream = 0;
    int device = 0;
    void* partial = nullptr;     // reduction scratch (kRedMaxBlocks doubles)
    // optional per-kernel timing (bench.py roofline leg): CUDA events around every launch on the launching stream
    struct Rec { const char* name; cudaEvent_t a, b; };
    struct Timing { bool on = false; std::vector<Rec> recs; std::map<std::string, std::pair<double, long>> acc; };
    Timing* timing = nullptr;    // shared by copies of the executor
    void tic(const char* name) {
        if (!timing || !timing->on) return;
        Rec r; r.name = name; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, stream); timing->recs.push_back(r);
    }
    void toc() { if (timing && timing->on) cudaEventRecord(timing->recs.back().b, stream); }
    void timing_enable(bool on) { timing_collect(); if (timing) { timing->on = on; if (on) timing->acc.clear(); } }
    std::string timing_report() {
        timing_collect();
        std::string out;
        if (timing) for (auto& kv : timing->acc)
            out += kv.first + " " + std::to_string(kv.second.second) + " " + std::to_string(kv.second.first) + "\n";
        return out;
    }
    void timing_collect() {
        if (!timing) return;
        cudaStreamSynchronize(stream);
        for (Rec& r : timing->recs) {
            float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
            auto& e = timing->acc[r.name]; e.first += ms; e.second += 1;
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        }
        timing->recs.clear();
    }

    // ---- CUDA graphs of whole steps: small meshes (the shipped cases: 500 - 50 000 cells) are bound by the launch
    // latency of the ~20 (primal) / ~45 (adjoint) kernels of a step. A step with a given key (time step, buffer
    // rotation, ...) runs eagerly the first time, is captured the second time and replayed from then on.
    struct GraphEntry { cudaGraphExec_t exec = nullptr; long launches = 0; int seen = 0; };
    struct Graphs { bool enabled = true; bool capturing = false; std::map<std::vector<unsigned long long>, GraphEntry> cache; };
    Graphs* graphs = nullptr;    // shared by copies of the executor
    // (the legacy default stream cannot be captured: graphs need a context created on an explicit stream)
    bool graph_usable() const { return stream != nullptr && graphs && graphs->enabled && !graphs->capturing && !(timing && timing->on); }
    bool graph_launch(const std::vector<unsigned long long>& key, long& launches) {
        auto it = graphs->cache.find(key);
        if (it == graphs->cache.end() || !it->second.exec) return false;
        FVM_CUDA_CHECK(cudaGraphLaunch(it->second.exec, stream));
        launches = it->second.launches;
        return true;
    }
    bool graph_first_time(const std::vector<unsigned long long>& key) {
        if (graphs->cache.size() > 64) { for (auto& kv : graphs->cache) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); graphs->cache.clear(); }
        return graphs->cache[key].seen++ == 0;
    }
    bool graph_begin() {
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); graphs->enabled = false; return false; }
        graphs->capturing = true;
        return true;
    }
    // ends the capture, instantiates and launches the graph once; false (graphs disabled) if anything failed: the caller
    // then runs the step eagerly
    bool graph_end_launch(const std::vector<unsigned long long>& key, long launches) {
        graphs->capturing = false;
        cudaGraph_t g = nullptr;
        cudaGraphExec_t e = nullptr;
        if (cudaStreamEndCapture(stream, &g) != cudaSuccess || !g || cudaGraphInstantiate(&e, g, 0) != cudaSuccess) {
            cudaGetLastError(); if (g) cudaGraphDestroy(g); graphs->enabled = false; return false;
        }
        cudaGraphDestroy(g);
        GraphEntry& ge = graphs->cache[key]; ge.exec = e; ge.launches = launches;
        FVM_CUDA_CHECK(cudaGraphLaunch(e, stream));
        return true;
    }

    void init(int dev, void* strm) {
        device = dev;
        FVM_CUDA_CHECK(cudaSetDevice(dev));
        stream = (cudaStream_t)strm;
        FVM_CUDA_CHECK(cudaMalloc(&partial, 5 * kRedMaxBlocks * sizeof(double)));    // (five-component reductions: reduce_sum5)
        timing = new Timing();
        graphs = new Graphs();
        configured = new std::map<const void*, size_t>();
        if (const char* e = getenv("ADFVM_NO_GRAPH")) graphs->enabled = !(e[0] && e[0] != '0');
    }
    std::map<const void*, size_t>* configured = nullptr;     // kernels whose shared-memory attribute has been raised
    // releases what init() and the lazily created side stream / graphs hold; called once by the owner (Solver)
    void destroy() {
        if (!timing) return;
        cudaStreamSynchronize(stream);
        timing_collect();
        for (auto& kv : graphs->cache) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (side) { cudaStreamSynchronize(side); cudaStreamDestroy(side); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); side = nullptr; }
        cudaFree(partial); partial = nullptr;
        delete timing; delete graphs; delete configured;
        timing = nullptr; graphs = nullptr; configured = nullptr;
    }
    void* stream_handle() const { return (void*)stream; }
    void* alloc(size_t bytes) { void* p = nullptr; FVM_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1)); return p; }
    void free(void* p) { cudaFree(p); }
    void zero(void* p, size_t bytes) { if (bytes) FVM_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, stream)); }
    void upload(void* dst, const void* src, size_t bytes) { if (bytes) FVM_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream)); }
    void download(void* dst, const void* src, size_t bytes) { if (bytes) FVM_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream)); }
    void copy(void* dst, const void* src, size_t bytes) { if (bytes) FVM_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream)); }
    void sync() { FVM_CUDA_CHECK(cudaStreamSynchronize(stream)); }

    template <class Body> void run(int n, const Body& b) { run_range(0, n, b); }
    // elements [first, first+n)
    template <class Body> void run_range(int first, int n, const Body& b) {
        tic(Body::kName);
        k_run<Body><<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(b, first, n);
        toc();
        FVM_CUDA_CHECK(cudaPeekAtLastError());
    }
    // ---- side stream for the halo exchange (high priority, so that its small kernels and the NCCL copies are not
    // queued behind the grid of the compute kernel they overlap with)
    cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void ensure_side() {
        if (side) return;
        int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
        FVM_CUDA_CHECK(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));
        FVM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        FVM_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    // fork: work issued between side_begin() and side_end() runs on the side stream, after everything issued on the
    // main stream so far; join(): the main stream waits for the side stream's work
    void side_begin() { ensure_side(); FVM_CUDA_CHECK(cudaEventRecord(ev_fork, stream)); FVM_CUDA_CHECK(cudaStreamWaitEvent(side, ev_fork, 0)); std::swap(stream, side); }
    void side_end() { FVM_CUDA_CHECK(cudaEventRecord(ev_join, stream)); std::swap(stream, side); }
    void join() { FVM_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_join, 0)); }
    template <class Body> void run_discard(int n, const Body& b) {
        tic(Body::kName);
        k_run_discard<Body><<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(b, n);
        toc();
        FVM_CUDA_CHECK(cudaPeekAtLastError());
    }
    // one CTA of kBlock threads per tile, dynamic shared memory sized by the body
    template <class Body> void run_tiles(int nTiles, const Body& b) { run_tiles_range(0, nTiles, b); }
    template <class Body> void run_tiles_range(int first, int nTiles, const Body& b) {
        if (nTiles <= 0) return;
        const size_t smem = Body::smem_bytes();
        if (smem > 48 * 1024) {          // opt in to large dynamic shared memory once per kernel and executor (the attribute is per device)
            size_t& done = (*configured)[(const void*)k_tile<Body>];
            if (smem > done) {
                FVM_CUDA_CHECK(cudaFuncSetAttribute(k_tile<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                done = smem;
            }
        }
        tic(Body::kName);
        k_tile<Body><<<nTiles, Body::kThreads, smem, stream>>>(b, first);
        toc();
        FVM_CUDA_CHECK(cudaPeekAtLastError());
    }
    template <typename R> void reduce_max_buffer(const R* in, int n, R* out) { reduce<R, OpMax, BufferBody<R>>(n, BufferBody<R>{in}, out); }
    template <typename R, class Op, class Body> void reduce(int n, const Body& b, R* out) {
        int nb = (n + kRedBlock - 1) / kRedBlock;
        if (nb > kRedMaxBlocks) nb = kRedMaxBlocks;
        if (nb < 1) nb = 1;
        tic(Body::kName);
        k_reduce1<R, Op, Body><<<nb, kRedBlock, 0, stream>>>(b, n, (R*)partial);
        toc();
        k_reduce2<R, Op><<<1, kRedBlock, 0, stream>>>((const R*)partial, nb, out);
        FVM_CUDA_CHECK(cudaPeekAtLastError());
    }
    template <class Body, typename R> void reduce_sum(int n, const Body& b, R* out) { reduce<R, OpSum, Body>(n, b, out); }
    // out[0..4] = sum over i of b(i).v[0..4]; the body may update element i as a side effect (every i is visited exactly once)
    template <class Body, typename R> void reduce_sum5(int n, const Body& b, R* out) {
        typedef decltype(b(0)) V;
        int nb = (n + kRedBlock - 1) / kRedBlock;
        if (nb > kRedMaxBlocks) nb = kRedMaxBlocks;
        if (nb < 1) nb = 1;
        tic(Body::kName);
        k_reduce1_v5<V, Body><<<nb, kRedBlock, 0, stream>>>(b, n, (V*)partial);
        toc();
        k_reduce2_v5<V, R><<<1, kRedBlock, 0, stream>>>((const V*)partial, nb, out);
        FVM_CUDA_CHECK(cudaPeekAtLastError());
    }
    template <class Body, typename R> void reduce_max(int n, const Body& b, R* out) { reduce<R, OpMax, Body>(n, b, out); }
};

}  // namespace fvm
