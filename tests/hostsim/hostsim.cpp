// TEST INFRASTRUCTURE ONLY — CPU simulator of the device code.
// Compiles the very same per-element bodies (adfvm_b200/csrc/fvm_bodies.h, fvm_math.h) and the same step
// orchestration (fvm_solver.h) and C ABI (fvm_capi.inc) with a plain-loop executor, so that the arithmetic, the
// indexing, the reverse sweep and the multi-rank halo logic can be checked against the oracle and the golden
// fixtures in the GPU-less build container (`-m "not gpu"` tests). It is never loaded by the adfvm_b200
// package: the product library is adfvm_b200/csrc/libadfvm_b200.so (CUDA only) and fails loudly without it.
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <map>
#include <mutex>
#include <condition_variable>
#include <unistd.h>
#include <algorithm>
#include <cstdio>
#include "../../adfvm_b200/csrc/fvm_solver.h"

namespace fvm {
struct HostExec {
    void* stream_handle() const { return nullptr; }
    void* alloc(size_t b) { return std::malloc(b ? b : 1); }
    void free(void* p) { std::free(p); }
    void zero(void* p, size_t b) { std::memset(p, 0, b); }
    void upload(void* d, const void* s, size_t b) { std::memcpy(d, s, b); }
    void download(void* d, const void* s, size_t b) { std::memcpy(d, s, b); }
    void copy(void* d, const void* s, size_t b) { std::memcpy(d, s, b); }
    void sync() {}
    void destroy() {}
    void timing_enable(bool) {}
    std::string timing_report() { return ""; }
    template <class B> void run(int n, const B& b) { run_range(0, n, b); }
    template <class B> void run_range(int first, int n, const B& b) {
        #pragma omp parallel for schedule(static)
        for (int i = first; i < first + n; i++) b(i);
    }
    bool graph_usable() const { return false; }
    bool graph_launch(const std::vector<unsigned long long>&, long&) { return false; }
    bool graph_first_time(const std::vector<unsigned long long>&) { return true; }
    bool graph_begin() { return false; }
    bool graph_end_launch(const std::vector<unsigned long long>&, long) { return false; }
    void side_begin() {}
    void side_end() {}
    void join() {}
    template <class B> void run_discard(int n, const B& b) {
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) (void)b(i);
    }
    template <class B> void run_tiles(int nTiles, const B& b) { run_tiles_range(0, nTiles, b); }
    template <class B> void run_tiles_range(int first, int nTiles, const B& b) {
        #pragma omp parallel for schedule(static)
        for (int t = first; t < first + nTiles; t++) b.host_tile(t);
    }
    template <typename R> void reduce_max_buffer(const R* in, int n, R* out) { R a = (R)-1e30; for (int i = 0; i < n; i++) if (in[i] > a) a = in[i]; *out = a; }
    template <class B, typename R> void reduce_sum(int n, const B& b, R* out) { R a = 0; for (int i = 0; i < n; i++) a += b(i); *out = a; }
    template <class B, typename R> void reduce_sum5(int n, const B& b, R* out) {
        R a[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < n; i++) { const auto x = b(i); for (int k = 0; k < 5; k++) a[k] += x.v[k]; }
        for (int k = 0; k < 5; k++) out[k] = a[k];
    }
    template <class B, typename R> void reduce_max(int n, const B& b, R* out) { R a = (R)-1e30; for (int i = 0; i < n; i++) { R v = b(i); if (v > a) a = v; } *out = a; }
};
}  // namespace fvm

namespace {
// In-process multi-rank halo for tests: ranks are threads of one process that meet at a rendezvous keyed by the
// 128-byte id. Each exchange publishes the send buffer, waits for all ranks, copies the peer's blocks, waits again.
struct Rendezvous {
    std::mutex mu; std::condition_variable cv; int n = 0, arrived = 0, gen = 0;
    std::vector<const void*> send; std::vector<std::vector<fvm::PatchHost>> remote; std::vector<double> scal;
    void barrier() {
        std::unique_lock<std::mutex> l(mu);
        int g = gen;
        if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};
std::mutex g_mu; std::map<std::string, Rendezvous*> g_rv;

template <typename R> struct ThreadHalo : fvm::HaloComm<R> {
    Rendezvous* rv; int rank, nranks;
    ThreadHalo(const void* id, int rank_, int nranks_) : rank(rank_), nranks(nranks_) {
        std::lock_guard<std::mutex> l(g_mu);
        std::string key((const char*)id, 128);
        if (!g_rv.count(key)) { auto* r = new Rendezvous(); r->n = nranks; r->send.resize(nranks); r->remote.resize(nranks); r->scal.resize(nranks); g_rv[key] = r; }
        rv = g_rv[key];
    }
    void exchange(const R* send, R* recv, int ncomp, const std::vector<fvm::PatchHost>& remote, void*) override {
        rv->send[rank] = send; rv->remote[rank] = remote;
        rv->barrier();
        int first = remote.empty() ? 0 : remote[0].startFace;
        for (const auto& p : remote) first = std::min(first, p.startFace);
        for (const auto& p : remote) {
            // matching block on the peer: its patch towards me with the same tag
            const auto& pr = rv->remote[p.peer];
            int pfirst = pr.empty() ? 0 : pr[0].startFace;
            for (const auto& q : pr) pfirst = std::min(pfirst, q.startFace);
            bool found = false;
            for (const auto& q : pr) if (q.peer == rank && q.tag == p.tag) {
                if (q.nFaces != p.nFaces) throw std::runtime_error("processor patch size mismatch between ranks");
                const R* src = (const R*)rv->send[p.peer] + (size_t)(q.startFace - pfirst) * ncomp;
                std::memcpy(recv + (size_t)(p.startFace - first) * ncomp, src, (size_t)p.nFaces * ncomp * sizeof(R));
                found = true; break;
            }
            if (!found) throw std::runtime_error("no matching processor patch on the peer rank");
        }
        rv->barrier();
    }
    double allreduce(double v, bool mx) {
        rv->scal[rank] = v; rv->barrier();
        double r = rv->scal[0];
        for (int i = 1; i < nranks; i++) r = mx ? std::max(r, rv->scal[i]) : r + rv->scal[i];
        rv->barrier();
        return r;
    }
    double allreduce_sum(double v) override { return allreduce(v, false); }
    double allreduce_max(double v) override { return allreduce(v, true); }
    void allreduce_sum_device(R* buf, int n, void*) override { for (int i = 0; i < n; i++) buf[i] = (R)allreduce((double)buf[i], false); }
};
}  // namespace

typedef fvm::HostExec ExecT;
static const int kIsCuda = 0;
static void exec_init(ExecT&, int, void*) {}
template <typename R> fvm::HaloComm<R>* make_comm(ExecT&, const void* id, int rank, int nranks) { return new ThreadHalo<R>(id, rank, nranks); }
static void* host_alloc_pinned(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
static void host_free_pinned(void* p) { std::free(p); }
static int comm_unique_id(void* id128) {
    static int counter = 0;
    std::memset(id128, 0, 128);
    std::snprintf((char*)id128, 128, "hostsim-%d-%d", (int)getpid(), counter++);
    return 0;
}
#include <unistd.h>
#include <algorithm>
#include "../../adfvm_b200/csrc/fvm_capi.inc"


// ---- test-only: halo through host callbacks, so that a world_size-2 `gloo` test can drive the rank-local step
// logic (patch order, tags, reverse halo) from two real processes. Not part of the product ABI.
namespace {
typedef void (*exchange_fn)(const void* send, void* recv, int ncomp, int npatches, const int* peer, const int* tag,
                            const long* offset, const long* count, int scalar_bytes);
typedef double (*allreduce_fn)(double v, int is_max);
template <typename R> struct CallbackHalo : fvm::HaloComm<R> {
    exchange_fn ex; allreduce_fn ar;
    CallbackHalo(exchange_fn e, allreduce_fn a) : ex(e), ar(a) {}
    void exchange(const R* send, R* recv, int ncomp, const std::vector<fvm::PatchHost>& remote, void*) override {
        int first = remote.empty() ? 0 : remote[0].startFace;
        for (const auto& p : remote) first = std::min(first, p.startFace);
        std::vector<int> peer, tag; std::vector<long> off, cnt;
        for (const auto& p : remote) { peer.push_back(p.peer); tag.push_back(p.tag); off.push_back((long)(p.startFace - first) * ncomp); cnt.push_back((long)p.nFaces * ncomp); }
        ex(send, recv, ncomp, (int)remote.size(), peer.data(), tag.data(), off.data(), cnt.data(), (int)sizeof(R));
    }
    double allreduce_sum(double v) override { return ar(v, 0); }
    double allreduce_max(double v) override { return ar(v, 1); }
    void allreduce_sum_device(R* buf, int n, void*) override { for (int i = 0; i < n; i++) buf[i] = (R)ar((double)buf[i], 0); }
};
}  // namespace
extern "C" int adfvm_hostsim_comm_callback(adfvm_ctx* c, exchange_fn e, allreduce_fn a) {
    if (!c) return 1;
    if (c->bytes == 8) { delete c->cd; c->cd = new CallbackHalo<double>(e, a); c->sd->comm = c->cd; }
    else { delete c->cf; c->cf = new CallbackHalo<float>(e, a); c->sf->comm = c->cf; }
    return 0;
}
