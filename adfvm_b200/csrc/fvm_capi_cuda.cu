// Product library: C ABI (include/adfvm_b200.h) over the CUDA executor. Built for sm_100a only.
// NCCL is reached through dlopen so a single-GPU run needs no NCCL at all and a multi-GPU run shares the
// libnccl.so.2 already loaded by torch.distributed (same NVLink/NVSwitch transport, one copy in the process).
#include <dlfcn.h>
#include <algorithm>
#include <cstdint>
#include "fvm_solver.h"
#include "fvm_cuda_exec.cuh"

namespace {

// ---- minimal NCCL surface (ABI-stable subset of nccl.h 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat = 7, ncclDouble = 8 };
enum { ncclSum = 0, ncclMax = 2 };
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    void load() {
        if (h) return;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw std::runtime_error(std::string("cannot load libnccl: ") + dlerror());
#define L(name) *(void**)(&name) = dlsym(h, "nccl" #name); if (!name) throw std::runtime_error("libnccl lacks nccl" #name)
        L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(Send); L(Recv); L(GroupStart); L(GroupEnd); L(AllReduce); L(GetErrorString);
#undef L
    }
    void check(ncclResult_t r, const char* what) {
        if (r != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + GetErrorString(r));
    }
};
NcclApi g_nccl;

template <typename R> struct NcclHalo : fvm::HaloComm<R> {
    ncclComm_t comm = nullptr; int rank, nranks; cudaStream_t stream; double* dscal = nullptr;
    NcclHalo(fvm::CudaExec& ex, const void* id, int rank_, int nranks_) : rank(rank_), nranks(nranks_), stream(ex.stream) {
        g_nccl.load();
        ncclUniqueId uid; std::memcpy(&uid, id, sizeof(uid));
        g_nccl.check(g_nccl.CommInitRank(&comm, nranks, uid, rank), "ncclCommInitRank");
        FVM_CUDA_CHECK(cudaMalloc(&dscal, 2 * sizeof(double)));
    }
    ~NcclHalo() override { if (comm) g_nccl.CommDestroy(comm); cudaFree(dscal); }
    // one grouped send+recv per processor patch; both ranks of a pair post their shared patches in tag order
    void exchange(const R* send, R* recv, int ncomp, const std::vector<fvm::PatchHost>& remote, void* strm) override {
        std::vector<const fvm::PatchHost*> order;
        for (const auto& p : remote) order.push_back(&p);
        std::stable_sort(order.begin(), order.end(), [](const fvm::PatchHost* a, const fvm::PatchHost* b) {
            return a->peer != b->peer ? a->peer < b->peer : a->tag < b->tag; });
        int first = remote.empty() ? 0 : remote[0].startFace;
        for (const auto& p : remote) first = std::min(first, p.startFace);
        const int dt = sizeof(R) == 8 ? ncclDouble : ncclFloat;
        g_nccl.check(g_nccl.GroupStart(), "ncclGroupStart");
        for (const fvm::PatchHost* p : order) {
            if (p->nFaces == 0) continue;
            const size_t off = (size_t)(p->startFace - first) * ncomp, cnt = (size_t)p->nFaces * ncomp;
            g_nccl.check(g_nccl.Send(send + off, cnt, dt, p->peer, comm, (cudaStream_t)strm), "ncclSend");
            g_nccl.check(g_nccl.Recv(recv + off, cnt, dt, p->peer, comm, (cudaStream_t)strm), "ncclRecv");
        }
        g_nccl.check(g_nccl.GroupEnd(), "ncclGroupEnd");
    }
    double allreduce(double v, int op) {
        FVM_CUDA_CHECK(cudaMemcpyAsync(dscal, &v, sizeof(double), cudaMemcpyHostToDevice, stream));
        g_nccl.check(g_nccl.AllReduce(dscal, dscal + 1, 1, ncclDouble, op, comm, stream), "ncclAllReduce");
        double r; FVM_CUDA_CHECK(cudaMemcpyAsync(&r, dscal + 1, sizeof(double), cudaMemcpyDeviceToHost, stream));
        FVM_CUDA_CHECK(cudaStreamSynchronize(stream));
        return r;
    }
    void allreduce_sum_device(R* buf, int n, void* strm) override {
        g_nccl.check(g_nccl.AllReduce(buf, buf, (size_t)n, sizeof(R) == 8 ? ncclDouble : ncclFloat, ncclSum, comm, (cudaStream_t)strm), "ncclAllReduce");
    }
    double allreduce_sum(double v) override { return allreduce(v, ncclSum); }
    double allreduce_max(double v) override { return allreduce(v, ncclMax); }
};

}  // namespace

typedef fvm::CudaExec ExecT;
static const int kIsCuda = 1;
static void exec_init(ExecT& ex, int device, void* stream) { ex.init(device, stream); }
template <typename R> fvm::HaloComm<R>* make_comm(ExecT& ex, const void* id, int rank, int nranks) {
    return new NcclHalo<R>(ex, id, rank, nranks);
}
static void* host_alloc_pinned(size_t bytes) { void* p = nullptr; FVM_CUDA_CHECK(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault)); return p; }
static void host_free_pinned(void* p) { if (p) cudaFreeHost(p); }
static int comm_unique_id(void* id128) {
    g_nccl.load();
    ncclUniqueId uid;
    if (g_nccl.GetUniqueId(&uid) != 0) return 1;
    std::memcpy(id128, &uid, sizeof(uid));
    return 0;
}

#include "fvm_capi.inc"
