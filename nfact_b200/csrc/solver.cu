// Host-side runtime of libnfact_b200.so: handles, resident matrices, the NMF loops (multiplicative
// update and coordinate descent with sklearn's stopping rules), dual regression, ingest, and the C ABI.
//
// Reference call sites this replaces:
//   NFACT/decomp/decomposition/sso/nmfsso.py:32-59          nmf_decomp -> sklearn NMF.fit_transform
//   NFACT/dual_reg/dual_regression/dual_regression_methods.py:92-167   nmf_dual_regression
//   NFACT/base/matrix_handling.py:116-158, NFACT/decomp/decomposition/matrix_handling.py:73-100   ingest
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nfact_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace nfb {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

int sm_count() {
    static int value = 0;
    if (value == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&value, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || value <= 0) value = 148;
    }
    return value;
}

std::atomic<int64_t> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Device buffer.  alloc(): cudaMalloc / cudaFree -- long-lived solver state and everything that is shared through
// CUDA IPC (pool memory cannot be).  alloc_pooled(): stream-ordered cudaMallocAsync / cudaFreeAsync for matrices and
// per-call temporaries: cudaFree synchronises the whole device, so a call that ends on one stream would wait for the
// next subject's upload running on another (the prefetching dual-regression loop), and a pooled free does not.
// The pool keeps its default release threshold (unused memory goes back to the driver at the next synchronisation).
// Exact-size recycling of the large per-subject buffers of the nfact_dr loop (operand planes, cell mask, triplet
// staging: 32 GB per C3 subject), one cache per handle = per stream.  The stream-ordered pool is the wrong tool for
// these: with two subjects in flight on two streams its best-fit reuse hands stream B a block stream A freed and
// makes B wait for A's pending work (the ingest ran behind the regression instead of beside it), or it re-maps
// tens of GB (0.2 - 0.7 s stalls in cudaMallocAsync).  A buffer returned here was last used on the handle's own
// stream and is handed out again on that stream only: no dependency, no mapping.
struct BigCache {
    std::mutex mu;
    std::vector<std::pair<void*, size_t>> idle;      // oldest first
    size_t idle_bytes = 0;
    size_t cap_bytes = 0;                             // 0: not yet known (40 % of the device on first use)
    void* get(size_t bytes) {
        std::lock_guard<std::mutex> lock(mu);
        for (size_t i = 0; i < idle.size(); ++i)
            if (idle[i].second == bytes) {
                void* ptr = idle[i].first;
                idle.erase(idle.begin() + i);
                idle_bytes -= bytes;
                return ptr;
            }
        return nullptr;
    }
    // keeps the buffer for the next subject; buffers of sizes nobody asks for any more (subjects of another shape)
    // are evicted oldest first once the cache holds more than its share of the device
    void put(void* ptr, size_t bytes) {
        std::vector<void*> evict;
        {
            std::lock_guard<std::mutex> lock(mu);
            if (cap_bytes == 0) {
                size_t free_b = 0, total_b = 0;
                cap_bytes = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? total_b / 10 * 4 : (size_t(64) << 30);
            }
            idle.emplace_back(ptr, bytes);
            idle_bytes += bytes;
            while (!idle.empty() && (idle_bytes > cap_bytes || idle.size() > 16)) {
                evict.push_back(idle.front().first);
                idle_bytes -= idle.front().second;
                idle.erase(idle.begin());
            }
        }
        for (void* e : evict) cudaFree(e);
    }
    void clear() {
        std::lock_guard<std::mutex> lock(mu);
        for (auto& e : idle) cudaFree(e.first);
        idle.clear();
        idle_bytes = 0;
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool pooled = false;
    cudaStream_t pool_stream = nullptr;
    BigCache* cache = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p && cache) {
            cache->put(p, n * sizeof(T));
            p = nullptr;
            cache = nullptr;
        }
        if (p) {
            if (pooled && cudaFreeAsync(p, pool_stream) != cudaSuccess) {
                cudaGetLastError();      // e.g. the stream is gone: fall back to the synchronous free
                cudaFree(p);
            } else if (!pooled) {
                cudaFree(p);
            }
        }
        p = nullptr;
        n = 0;
        pooled = false;
    }
    int alloc(size_t count, bool zero, cudaStream_t stream) {
        release();
        if (count == 0) count = 1;
        NFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
        n = count;
        if (zero) NFB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), stream));
        return 0;
    }
    // from the handle's exact-size cache (a miss is a plain cudaMalloc); goes back to the cache on release
    int alloc_cached(BigCache* c, size_t count, cudaStream_t stream) {
        release();
        if (count == 0) count = 1;
        p = static_cast<T*>(c->get(count * sizeof(T)));
        if (p == nullptr && cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            c->clear();          // out of memory with idle buffers of other sizes around: give them back and retry
            p = nullptr;
            NFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
        }
        n = count;
        cache = c;
        return 0;
    }
    int alloc_pooled(size_t count, bool zero, cudaStream_t stream) {
        release();
        if (count == 0) count = 1;
        static const bool use_pool = !(getenv("NFB_POOL") && atoi(getenv("NFB_POOL")) == 0);
        if (!use_pool) return alloc(count, zero, stream);
        NFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), stream));
        n = count;
        pooled = true;
        pool_stream = stream;
        if (zero) NFB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), stream));
        return 0;
    }
};

// An event whose last record is on a side stream that reads / writes buffers of the enclosing scope: on EVERY exit path
// (early NFB_TRY / NFB_CUDA returns included) the side stream is drained before those buffers are released, and the
// event is destroyed.  Declare it after the buffers it protects.
struct SideStreamEvent {
    cudaStream_t side = nullptr;
    cudaEvent_t ev = nullptr;
    explicit SideStreamEvent(cudaStream_t s) : side(s) {}
    SideStreamEvent(const SideStreamEvent&) = delete;
    SideStreamEvent& operator=(const SideStreamEvent&) = delete;
    int create() { return cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess ? 0 : 1; }
    void finish() {
        if (ev) {
            cudaStreamSynchronize(side);
            cudaEventDestroy(ev);
            ev = nullptr;
        }
    }
    ~SideStreamEvent() { finish(); }
};

typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclReduceScatterFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclAllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
constexpr int kNcclFloat = 7, kNcclDouble = 8, kNcclSum = 0;

}  // namespace nfb

using namespace nfb;

struct nfb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int cta_group = 2;
    int64_t launches_last = 0;
    // optional per-kernel timing of the two X-streaming GEMMs (CUDA events on the launch stream)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
    // NFB_PHASE_TIMING=1: CUDA events around every phase of an iteration (debug aid, printed by nfb_nmf_iterate)
    bool phase_timing = false;
    std::vector<std::pair<const char*, std::pair<cudaEvent_t, cudaEvent_t>>> phase_events;
    // multi-GPU (row-sharded) -------------------------------------------------
    void* comm = nullptr;
    NcclAllReduceFn allreduce = nullptr;
    NcclReduceScatterFn reduce_scatter = nullptr;
    NcclAllGatherFn all_gather = nullptr;
    int rank = 0, world = 1;
    void* nccl_lib = nullptr;
    bool own_comm = false;
    // scratch -----------------------------------------------------------------
    DevBuf<double> d_scal;        // small fp64 scalars
    DevBuf<unsigned int> d_bits;  // absmax bits
    DevBuf<int> d_int;
    double* h_scal = nullptr;     // pinned mirror
    cudaStream_t copy_stream = nullptr;   // D2H of finished results while the compute stream carries on
    BigCache big;                         // per-subject buffers of the dual-regression loop (see BigCache)
    // second compute stream: the small k x k "denominator" GEMMs run beside the X-streaming GEMM they do not
    // depend on (their CTAs fill the SMs the big launch's last, partial wave leaves idle); NFB_DEN_OVERLAP=0: off
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    bool den_overlap = true;
    bool gram_overlap = true;    // the Gram products as well (NFB_DEN_OVERLAP=1: denominators only)
};

struct nfb_matrix {
    nfb_handle* h = nullptr;
    int64_t m = 0, n = 0, ldp = 0;
    DevBuf<__half> hi, lo;
    DevBuf<float> inv_scale;
    float inv_scale_host = 1.0f;
    double sumsq = 0.0;  // local shard
    double sum = 0.0;
    int has_negative = 0;
    GemmOperand operand() const { return GemmOperand{hi.p, lo.p, m, ldp}; }
};

namespace nfb {

// A factor matrix stored component-major: fp32 master [kp][ld] plus fp16 planes with per-row scales.
struct Factor {
    int64_t k = 0, kp = 0, dim = 0, ld = 0;
    DevBuf<float> f;
    DevBuf<__half> hi, lo;
    DevBuf<float> inv_scale;
    DevBuf<unsigned int> rowmax;
    int init(int64_t k_, int64_t dim_, bool with_master, cudaStream_t stream, int64_t ld_override = 0,
             bool pooled = false) {
        k = k_;
        kp = round_up64(k_, 16);
        dim = dim_;
        ld = ld_override > 0 ? ld_override : round_up64(std::max<int64_t>(dim_, 1), 8);
        auto get = [&](auto& buf, size_t count) {
            return pooled ? buf.alloc_pooled(count, true, stream) : buf.alloc(count, true, stream);
        };
        if (with_master) NFB_TRY(get(f, kp * ld));
        NFB_TRY(get(hi, kp * ld));
        NFB_TRY(get(lo, kp * ld));
        NFB_TRY(get(inv_scale, kp));
        NFB_TRY(get(rowmax, kp));
        return 0;
    }
    GemmOperand operand() const { return GemmOperand{hi.p, lo.p, kp, ld}; }
};

static int refresh_planes(Factor& fac, const float* src, const float* col_mul, bool recompute_max,
                          cudaStream_t stream) {
    if (recompute_max) {
        NFB_CUDA(cudaMemsetAsync(fac.rowmax.p, 0, sizeof(unsigned int) * fac.kp, stream));
        NFB_TRY(rowabsmax_f32(src, fac.k, fac.dim, fac.ld, col_mul, fac.rowmax.p, stream));
        note_launch(1);
    }
    NFB_TRY(split_factor(src, fac.k, fac.dim, fac.ld, col_mul, fac.rowmax.p, fac.hi.p, fac.lo.p, fac.ld,
                         fac.inv_scale.p, stream));
    note_launch(1);
    return 0;
}

}  // namespace nfb

struct nfb_nmf {
    nfb_handle* h = nullptr;
    const nfb_matrix* x = nullptr;
    int k = 0;
    int64_t kp = 0;
    nfb_nmf_params p{};
    Factor wt, hh;        // W^T [k, m] and H [k, n]
    Factor gw, gh;        // Gram matrices W'W and HH' (fp32 [kp][kp]); planes are only used as scratch
    Factor gd;            // planes of a Gram matrix with the K-side scales folded in (denominator GEMM)
    DevBuf<float> part_w, part_h, part_den, part_gram, num_h;
    DevBuf<float> sk_ws;              // stream-K workspace of the two X-streaming GEMMs (main stream only)
    DevBuf<unsigned int> sk_cnt;
    GemmStreamK sk{};
    bool use_sk = false;
    DevBuf<float> num_sl, pack;   // sharded runs: slice-major numerator / gather buffer [P][k][ns], packed slice
    DevBuf<float> snap_w[2], snap_h[2];   // nfb_nmf_run (CD): factors after the last two iterations (see there)
    cudaEvent_t viol_ready[2] = {nullptr, nullptr};
    int64_t slice_cols = 0;       // ns: H columns per rank (multiple of 8)
    // NVLink peer-memory path (p2p.cu): arena + H master of every rank mapped through CUDA IPC
    bool p2p = false;
    // H columns are sliced over the ranks (p2p or NCCL reduce-scatter / all-gather) rather than replicated.
    // Decided once from rank-independent state: every rank must issue the same collectives even when its own
    // slice is empty or happens to cover all of H (n <= slice_cols).
    bool h_sliced = false;
    DevBuf<char> arena;
    P2pPeers peers{};
    std::vector<void*> ipc_opened;
    int epoch = 0;
    // byte offsets inside the arena (p2p.cu): gram_in[P][kp*kp], rowmax_in[P][kp], num_in[P][k][ns]
    int64_t arena_gram_off = 0, arena_rowmax_off = 0, arena_num_off = 0;
    float* arena_num() const { return reinterpret_cast<float*>(arena.p + arena_num_off); }
    float* arena_gram() const {       // this rank's own slot of gram_in
        return reinterpret_cast<float*>(arena.p + arena_gram_off) + static_cast<int64_t>(h->rank) * kp * kp;
    }
    PeerMirror h_mirror(int64_t col0) const {   // the H columns from col0 on in every OTHER rank's H master
        PeerMirror mir{};
        if (!p2p) return mir;
        for (int r = 0; r < h->world; ++r)
            if (r != h->rank) mir.f[mir.n++] = static_cast<float*>(peers.hmaster[r]) + col0;
        return mir;
    }
    int s_w = 1, s_h = 1, s_gw = 1, s_gh = 1;
    double sumsq_global = 0.0;
    bool p1_valid = false;  // part_w holds X H' for the current H
    bool den_w_valid = false;    // part_den holds (HH') W' for the current W planes and Gram (launched beside X H')
    bool den_h_done = false;     // gemm_wtx already launched the H denominator on the aux stream
    bool gh_valid = false;       // gh holds HH' of the current H planes (with the overlap it is computed beside X H')
    bool aux_pending = false;    // work on the aux stream that the main stream has not joined yet
    double last_violation = 0.0;
};

namespace nfb {

static int allreduce_f32(nfb_handle* h, float* buf, size_t count) {
    if (h->world <= 1) return 0;
    NFB_CHECK(h->allreduce != nullptr && h->comm != nullptr, "multi-GPU run without a communicator");
    int rc = h->allreduce(buf, buf, count, kNcclFloat, kNcclSum, h->comm, h->stream);
    NFB_CHECK(rc == 0, "ncclAllReduce(float) failed with code " + std::to_string(rc));
    return 0;
}
static int allreduce_f64(nfb_handle* h, double* buf, size_t count) {
    if (h->world <= 1) return 0;
    NFB_CHECK(h->allreduce != nullptr && h->comm != nullptr, "multi-GPU run without a communicator");
    int rc = h->allreduce(buf, buf, count, kNcclDouble, kNcclSum, h->comm, h->stream);
    NFB_CHECK(rc == 0, "ncclAllReduce(double) failed with code " + std::to_string(rc));
    return 0;
}

// G = F F' (fp32 [k][kp]) through the tensor-core path; partial sums over `splits` K ranges.
// to_arena: leave the local Gram in the peer-visible arena; the pull kernel of the H half-step sums all ranks'
static int compute_gram(nfb_nmf* s, Factor& fac, Factor& gram, int splits, bool allreduce, bool to_arena = false,
                        cudaStream_t stream = nullptr) {
    nfb_handle* h = s->h;
    if (stream == nullptr) stream = h->stream;
    NFB_TRY(gemm_split(fac.operand(), false, fac.operand(), fac.k, fac.dim, static_cast<int>(fac.kp),
                       static_cast<int>(fac.k), s->part_gram.p, gram.ld, splits, fac.inv_scale.p, fac.inv_scale.p,
                       1.0f, h->cta_group, stream));
    NFB_TRY(reduce_partials(s->part_gram.p, splits, fac.k * gram.ld, fac.k, fac.k, gram.ld, 0.0f,
                            to_arena ? s->arena_gram() : gram.f.p, gram.ld, stream));
    note_launch(2);
    if (allreduce && !to_arena) NFB_TRY(allreduce_f32(h, gram.f.p, static_cast<size_t>(gram.kp * gram.ld)));
    return 0;
}

// part_w = X H'   (per split) for the current H planes
struct PhaseTimer {
    nfb_handle* h;
    const char* name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    PhaseTimer(nfb_handle* handle, const char* phase) : h(handle), name(phase) {
        if (h->phase_timing && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess)
            cudaEventRecord(e0, h->stream);
    }
    ~PhaseTimer() {
        if (e0 && e1) {
            cudaEventRecord(e1, h->stream);
            h->phase_events.push_back({name, {e0, e1}});
        }
    }
};
#define NFB_PHASE(handle, name) PhaseTimer phase_timer_##__LINE__(handle, name)

static void report_phases(nfb_handle* h, int iters) {
    if (!h->phase_timing || h->phase_events.empty()) return;
    cudaStreamSynchronize(h->stream);
    std::vector<std::pair<std::string, double>> totals;
    for (auto& ev : h->phase_events) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev.second.first, ev.second.second);
        cudaEventDestroy(ev.second.first);
        cudaEventDestroy(ev.second.second);
        bool found = false;
        for (auto& t : totals)
            if (t.first == ev.first) { t.second += ms; found = true; }
        if (!found) totals.push_back({ev.first, ms});
    }
    h->phase_events.clear();
    double sum = 0.0;
    for (auto& t : totals) sum += t.second;
    fprintf(stderr, "nfact_b200 phases over %d iteration(s): ", iters);
    for (auto& t : totals) fprintf(stderr, "%s=%.3f ms  ", t.first.c_str(), t.second / iters);
    fprintf(stderr, "| sum=%.3f ms/iter\n", sum / iters);
}

struct GemmTimer {
    nfb_handle* h;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    explicit GemmTimer(nfb_handle* handle) : h(handle) {
        if (h->profiling && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess)
            cudaEventRecord(e0, h->stream);
    }
    ~GemmTimer() {
        if (e0 && e1) {
            cudaEventRecord(e1, h->stream);
            h->gemm_events.emplace_back(e0, e1);
        }
    }
};

static int gemm_xht(nfb_nmf* s) {
    nfb_handle* h = s->h;
    GemmTimer timer(h);
    NFB_TRY(gemm_split(s->x->operand(), false, s->hh.operand(), s->x->m, s->x->n, static_cast<int>(s->kp), s->k,
                       s->part_w.p, s->wt.ld, s->s_w, s->hh.inv_scale.p, nullptr, s->x->inv_scale_host,
                       h->cta_group, h->stream));
    note_launch(1);
    s->p1_valid = true;
    return 0;
}

// part_h = W'X   (per split) for the current W planes.  Single GPU: the consumer sums the split slabs.
// Row-sharded runs: the slabs are summed into a slice-major buffer [P][k][ns] and reduce-scattered, so
// that rank r receives only the numerator of the H columns it will update (its slice).
struct HNumerator {
    const float* num;   // [splits][k][ld_part]
    int splits;
    int64_t stride;
    int64_t ld_part;
    int64_t col0;       // first H column this rank updates
    int64_t width;      // number of H columns this rank updates
};

static int aux_fork(nfb_handle* h);
static int gemm_den(nfb_nmf* s, Factor& fac, Factor& gram, int64_t col0, int64_t width, int64_t ld_out,
                    cudaStream_t stream = nullptr);

static int gemm_wtx(nfb_nmf* s, HNumerator* out) {
    nfb_handle* h = s->h;
    const int64_t n = s->x->n;
    if (s->p2p) {
        // Peer-memory path: the k x k Gram goes to every peer first (small), then the unsplit GEMM whose epilogue
        // stores each finished tile into the inbox of the rank that owns those H columns -- the reduce-scatter
        // happens inside the GEMM, tile by tile, over NVLink.  Then: post, wait for everybody, sum the Grams.
        const int64_t ns = s->slice_cols;
        const int gram_elems = static_cast<int>(s->kp * s->kp);
        // W'W of this shard, its push to the peers' slots and its flag need only the planes of W: with the aux stream
        // they run beside the GEMM as well, like the Gram sum and the denominator behind them
        const bool side = h->den_overlap;
        cudaStream_t gst = side ? h->aux_stream : h->stream;
        if (side) NFB_TRY(aux_fork(h));
        NFB_TRY(compute_gram(s, s->wt, s->gw, s->s_gw, false, true, gst));
        NFB_TRY(p2p_push_gram(s->peers, h->world, h->rank, s->arena_gram_off, gram_elems, gst));
        ++s->epoch;
        NFB_TRY(p2p_post(s->peers, h->world, h->rank, kP2pGramFlag, s->epoch, gst));
        if (side) s->aux_pending = true;
        GemmScatter sc{};
        for (int r = 0; r < h->world; ++r)
            sc.base[r] = reinterpret_cast<float*>(static_cast<char*>(s->peers.arena[r]) + s->arena_num_off);
        sc.slice = ns;
        sc.rows = s->k;
        sc.src = h->rank;
        {
        GemmTimer timer(h);
        NFB_TRY(gemm_split(s->x->operand(), true, s->wt.operand(), n, s->x->m, static_cast<int>(s->kp), s->k,
                           s->arena_num(), ns, 1, s->wt.inv_scale.p, nullptr, s->x->inv_scale_host, h->cta_group,
                           h->stream, &sc));
        }
        const int64_t c0 = std::min<int64_t>(n, h->rank * ns);
        const int64_t width = std::min<int64_t>(ns, n - c0);
        // the P inbox slabs are the split-K partials of this rank's columns: the consumer sums them in rank order
        *out = HNumerator{s->arena_num(), h->world, static_cast<int64_t>(s->k) * ns, ns, c0, width};
        // the Gram sum needs only the peers' Gram flags (raised before their GEMMs): with the aux stream it and the
        // denominator GEMM run beside this rank's W'X instead of after it
        NFB_TRY(p2p_wait_sum_gram(s->peers, h->world, h->rank, kP2pGramFlag, s->epoch, s->arena_gram_off, gram_elems,
                                  s->gw.f.p, gst));
        if (side) {
            NFB_TRY(gemm_den(s, s->hh, s->gw, c0, width, ns, h->aux_stream));
            s->aux_pending = true;
            s->den_h_done = true;
        }
        NFB_PHASE(h, "rs_wait");
        NFB_TRY(p2p_post_wait(s->peers, h->world, h->rank, kP2pRsFlag, s->epoch, h->stream));
        note_launch(5);
        return 0;
    }
    {
    GemmTimer timer(h);
    NFB_TRY(gemm_split(s->x->operand(), true, s->wt.operand(), s->x->n, s->x->m, static_cast<int>(s->kp), s->k,
                       s->part_h.p, s->hh.ld, s->s_h, s->wt.inv_scale.p, nullptr, s->x->inv_scale_host, h->cta_group,
                       h->stream, nullptr, s->use_sk ? &s->sk : nullptr));
    }
    note_launch(1);
    if (h->world <= 1) {
        *out = HNumerator{s->part_h.p, s->s_h, static_cast<int64_t>(s->k) * s->hh.ld, s->hh.ld, 0, n};
        return 0;
    }
    if (h->reduce_scatter != nullptr && h->all_gather != nullptr) {
        const int64_t ns = s->slice_cols;
        NFB_TRY(reduce_partials_sliced(s->part_h.p, s->s_h, static_cast<int64_t>(s->k) * s->hh.ld, s->k, n, s->hh.ld,
                                       s->num_sl.p, ns, h->stream));
        note_launch(1);
        int rc = h->reduce_scatter(s->num_sl.p, s->num_h.p, static_cast<size_t>(s->k) * ns, kNcclFloat, kNcclSum,
                                   h->comm, h->stream);
        NFB_CHECK(rc == 0, "ncclReduceScatter failed with code " + std::to_string(rc));
        const int64_t c0 = std::min<int64_t>(n, h->rank * ns);
        *out = HNumerator{s->num_h.p, 1, 0, ns, c0, std::min<int64_t>(ns, n - c0)};
        return 0;
    }
    // communicator without reduce-scatter / all-gather: replicate the H update on every rank
    NFB_TRY(reduce_partials(s->part_h.p, s->s_h, static_cast<int64_t>(s->k) * s->hh.ld, s->k, n, s->hh.ld, 0.0f,
                            s->num_h.p, s->hh.ld, h->stream));
    note_launch(1);
    NFB_TRY(allreduce_f32(h, s->num_h.p, static_cast<size_t>(s->k) * s->hh.ld));
    *out = HNumerator{s->num_h.p, 1, 0, s->hh.ld, 0, n};
    return 0;
}

// after a sliced H update: every rank contributes its columns, everyone ends up with the full H
static int gather_h_slices(nfb_nmf* s, const HNumerator& hn) {
    nfb_handle* h = s->h;
    if (!s->h_sliced) return 0;
    const int64_t ns = s->slice_cols, n = s->x->n;
    if (s->p2p) {
        // the slice itself travelled inside the sweep / update kernel (PeerMirror); what is left is this rank's two
        // violation partial sums (CD) and the per-component maxima of its slice, the flag, and the wait
        double* scal = s->p.solver == 0 ? h->d_scal.p + 2 : nullptr;
        NFB_PHASE(h, "ag_wait");
        NFB_TRY(p2p_post_gather(s->peers, h->world, h->rank, s->epoch, scal, kP2pScalOffset, scal, s->hh.rowmax.p,
                                s->arena_rowmax_off, static_cast<int>(s->kp), h->stream));
        note_launch(1);
        return 0;
    }
    if (hn.width > 0)
        NFB_CUDA(cudaMemcpy2DAsync(s->pack.p, ns * sizeof(float), s->hh.f.p + hn.col0, s->hh.ld * sizeof(float),
                                   hn.width * sizeof(float), s->k, cudaMemcpyDeviceToDevice, h->stream));
    int rc = h->all_gather(s->pack.p, s->num_sl.p, static_cast<size_t>(s->k) * ns, kNcclFloat, h->comm, h->stream);
    NFB_CHECK(rc == 0, "ncclAllGather failed with code " + std::to_string(rc));
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) continue;
        const int64_t c0 = std::min<int64_t>(n, r * ns), width = std::min<int64_t>(ns, n - c0);
        if (width <= 0) continue;
        NFB_CUDA(cudaMemcpy2DAsync(s->hh.f.p + c0, s->hh.ld * sizeof(float),
                                   s->num_sl.p + static_cast<int64_t>(r) * s->k * ns, ns * sizeof(float),
                                   width * sizeof(float), s->k, cudaMemcpyDeviceToDevice, h->stream));
    }
    return 0;
}

// G F for the columns [col0, col0 + width) of F:  part_den[t][j] = sum_r G[t][r] F[r][col0 + j]
// (denominator of the multiplicative update; seed of the coordinate-descent gradient)
static int gemm_den(nfb_nmf* s, Factor& fac, Factor& gram, int64_t col0, int64_t width, int64_t ld_out,
                    cudaStream_t stream) {
    nfb_handle* h = s->h;
    if (stream == nullptr) stream = h->stream;
    // fold the per-component scales of F's planes (the K side of this product) into G's columns
    NFB_TRY(refresh_planes(s->gd, gram.f.p, fac.inv_scale.p, true, stream));
    if (width <= 0) return 0;
    GemmOperand a{fac.hi.p + col0, fac.lo.p + col0, fac.k, fac.ld};
    NFB_TRY(gemm_split(a, true, s->gd.operand(), width, fac.k, static_cast<int>(s->kp), s->k, s->part_den.p, ld_out,
                       1, s->gd.inv_scale.p, nullptr, 1.0f, h->cta_group, stream));
    note_launch(1);
    return 0;
}

// The denominator GEMM of a half-step needs the factor's planes and the OTHER factor's Gram -- not the X-streaming
// product that precedes the update.  fork: everything enqueued on the main stream so far is the aux stream's
// starting point; the caller then launches the big GEMM on the main stream FIRST (its persistent CTAs take the
// SMs) and the denominator work on the aux stream, whose GEMM CTAs start as the big launch's CTAs retire.
static int aux_fork(nfb_handle* h) {
    NFB_CUDA(cudaEventRecord(h->aux_fork, h->stream));
    NFB_CUDA(cudaStreamWaitEvent(h->aux_stream, h->aux_fork, 0));
    return 0;
}
static int aux_join(nfb_nmf* s) {
    if (!s->aux_pending) return 0;
    nfb_handle* h = s->h;
    NFB_CUDA(cudaEventRecord(h->aux_join, h->aux_stream));
    NFB_CUDA(cudaStreamWaitEvent(h->stream, h->aux_join, 0));
    s->aux_pending = false;
    return 0;
}

// X H' for the current H on the main stream and, beside it, the denominator of the W half-step
static int gemm_xht_with_den(nfb_nmf* s) {
    nfb_handle* h = s->h;
    if (!h->den_overlap) {
        if (!s->gh_valid) NFB_TRY(compute_gram(s, s->hh, s->gh, s->s_gh, false));
        s->gh_valid = true;
        return gemm_xht(s);
    }
    NFB_TRY(aux_fork(h));
    NFB_TRY(gemm_xht(s));
    if (!s->gh_valid) NFB_TRY(compute_gram(s, s->hh, s->gh, s->s_gh, false, false, h->aux_stream));
    s->gh_valid = true;
    NFB_TRY(gemm_den(s, s->wt, s->gh, 0, s->x->m, s->wt.ld, h->aux_stream));
    s->den_w_valid = true;
    s->aux_pending = true;
    return 0;
}

// refresh planes + Gram of H after its update (slices gathered first in sharded runs)
static int finish_h(nfb_nmf* s, const HNumerator& hn) {
    nfb_handle* h = s->h;
    const bool sliced = s->h_sliced;
    NFB_TRY(gather_h_slices(s, hn));
    NFB_TRY(refresh_planes(s->hh, s->hh.f.p, nullptr, sliced && !s->p2p, h->stream));   // p2p: maxima came with the gather
    // HH' is needed by the next W half-step (and the error): with the overlap it is computed beside the next X H'
    if (h->gram_overlap) s->gh_valid = false;
    else { NFB_TRY(compute_gram(s, s->hh, s->gh, s->s_gh, false)); s->gh_valid = true; }
    return 0;
}

static int mu_iteration(nfb_nmf* s) {
    nfb_handle* h = s->h;
    const nfb_nmf_params& p = s->p;
    if (!s->p1_valid) { NFB_PHASE(h, "gemm_xht"); NFB_TRY(gemm_xht_with_den(s)); }
    // ---- W half-step ----
    if (!s->den_w_valid) { NFB_PHASE(h, "den_w"); NFB_TRY(gemm_den(s, s->wt, s->gh, 0, s->x->m, s->wt.ld)); }
    NFB_TRY(aux_join(s));
    s->den_w_valid = false;
    NFB_CUDA(cudaMemsetAsync(s->wt.rowmax.p, 0, sizeof(unsigned int) * s->kp, h->stream));
    { NFB_PHASE(h, "update_w");
    NFB_TRY(mu_update(s->wt.f.p, s->k, s->x->m, s->wt.ld, s->part_w.p, s->s_w, static_cast<int64_t>(s->k) * s->wt.ld,
                      s->part_den.p, 1, 0, s->wt.ld, static_cast<float>(p.l1_reg_W), static_cast<float>(p.l2_reg_W),
                      s->wt.rowmax.p, h->stream));
    }
    note_launch(1);
    s->p1_valid = false;
    { NFB_PHASE(h, "split_w"); NFB_TRY(refresh_planes(s->wt, s->wt.f.p, nullptr, false, h->stream)); }
    const bool side = h->world <= 1 && h->den_overlap;
    // peer-memory path: W'W is computed, pushed and summed inside gemm_wtx (beside the GEMM)
    if (!s->p2p && (!side || !h->gram_overlap)) { NFB_PHASE(h, "gram_w"); NFB_TRY(compute_gram(s, s->wt, s->gw, s->s_gw, true)); }
    // ---- H half-step ----
    HNumerator hn;
    if (side) {
        // single GPU: W'W and the H denominator run beside the X-streaming GEMM (see gemm_xht_with_den)
        NFB_TRY(aux_fork(h));
        { NFB_PHASE(h, "gemm_wtx"); NFB_TRY(gemm_wtx(s, &hn)); }
        if (h->gram_overlap) NFB_TRY(compute_gram(s, s->wt, s->gw, s->s_gw, false, false, h->aux_stream));
        NFB_TRY(gemm_den(s, s->hh, s->gw, hn.col0, hn.width, hn.ld_part, h->aux_stream));
        s->aux_pending = true;
        NFB_TRY(aux_join(s));
    } else {
        { NFB_PHASE(h, "gemm_wtx"); NFB_TRY(gemm_wtx(s, &hn)); }
        if (s->den_h_done) {           // peer-memory path: launched beside the GEMM (gemm_wtx)
            NFB_TRY(aux_join(s));
            s->den_h_done = false;
        } else {
            NFB_PHASE(h, "den_h");
            NFB_TRY(gemm_den(s, s->hh, s->gw, hn.col0, hn.width, hn.ld_part));
        }
    }
    NFB_CUDA(cudaMemsetAsync(s->hh.rowmax.p, 0, sizeof(unsigned int) * s->kp, h->stream));
    if (hn.width > 0) {
        NFB_PHASE(h, "update_h");
        const PeerMirror mir = s->h_mirror(hn.col0);
        NFB_TRY(mu_update(s->hh.f.p + hn.col0, s->k, hn.width, s->hh.ld, hn.num, hn.splits, hn.stride, s->part_den.p,
                          1, 0, hn.ld_part, static_cast<float>(p.l1_reg_H), static_cast<float>(p.l2_reg_H),
                          s->hh.rowmax.p, h->stream, &mir));
        note_launch(1);
    }
    { NFB_PHASE(h, "finish_h"); NFB_TRY(finish_h(s, hn)); }
    return 0;
}

// d_scal layout: [0] cross, [1] gram dot, [2] violation W, [3] violation H
// violation_out: synchronous read-back of this iteration's violation; host_slot: asynchronous one into two
// pinned doubles (the caller waits on an event it records afterwards)
static int cd_iteration(nfb_nmf* s, double* violation_out, double* host_slot = nullptr) {
    nfb_handle* h = s->h;
    const nfb_nmf_params& p = s->p;
    NFB_CUDA(cudaMemsetAsync(h->d_scal.p + 2, 0, 2 * sizeof(double), h->stream));
    if (!s->p1_valid) { NFB_PHASE(h, "gemm_xht"); NFB_TRY(gemm_xht_with_den(s)); }
    // ---- W half-step ----
    if (!s->den_w_valid) { NFB_PHASE(h, "den_w"); NFB_TRY(gemm_den(s, s->wt, s->gh, 0, s->x->m, s->wt.ld)); }  // seeds the gradient
    NFB_TRY(aux_join(s));
    s->den_w_valid = false;
    NFB_CUDA(cudaMemsetAsync(s->wt.rowmax.p, 0, sizeof(unsigned int) * s->kp, h->stream));
    { NFB_PHASE(h, "update_w");
    NFB_TRY(cd_sweep(s->wt.f.p, s->k, s->x->m, s->wt.ld, s->part_w.p, s->s_w, static_cast<int64_t>(s->k) * s->wt.ld,
                     s->part_den.p, s->wt.ld, s->gh.f.p, static_cast<float>(p.l1_reg_W),
                     static_cast<float>(p.l2_reg_W), h->d_scal.p + 2, s->wt.rowmax.p, h->stream));
    }
    note_launch(1);
    s->p1_valid = false;
    { NFB_PHASE(h, "split_w"); NFB_TRY(refresh_planes(s->wt, s->wt.f.p, nullptr, false, h->stream)); }
    const bool side = h->world <= 1 && h->den_overlap;
    // peer-memory path: W'W is computed, pushed and summed inside gemm_wtx (beside the GEMM)
    if (!s->p2p && (!side || !h->gram_overlap)) { NFB_PHASE(h, "gram_w"); NFB_TRY(compute_gram(s, s->wt, s->gw, s->s_gw, true)); }
    // ---- H half-step ----
    HNumerator hn;
    if (side) {
        // single GPU: W'W and the H denominator run beside the X-streaming GEMM (see gemm_xht_with_den)
        NFB_TRY(aux_fork(h));
        { NFB_PHASE(h, "gemm_wtx"); NFB_TRY(gemm_wtx(s, &hn)); }
        if (h->gram_overlap) NFB_TRY(compute_gram(s, s->wt, s->gw, s->s_gw, false, false, h->aux_stream));
        NFB_TRY(gemm_den(s, s->hh, s->gw, hn.col0, hn.width, hn.ld_part, h->aux_stream));
        s->aux_pending = true;
        NFB_TRY(aux_join(s));
    } else {
        { NFB_PHASE(h, "gemm_wtx"); NFB_TRY(gemm_wtx(s, &hn)); }
        if (s->den_h_done) {           // peer-memory path: launched beside the GEMM (gemm_wtx)
            NFB_TRY(aux_join(s));
            s->den_h_done = false;
        } else {
            NFB_PHASE(h, "den_h");
            NFB_TRY(gemm_den(s, s->hh, s->gw, hn.col0, hn.width, hn.ld_part));
        }
    }
    NFB_CUDA(cudaMemsetAsync(s->hh.rowmax.p, 0, sizeof(unsigned int) * s->kp, h->stream));
    if (hn.width > 0) {
        NFB_PHASE(h, "update_h");
        const PeerMirror mir = s->h_mirror(hn.col0);
        NFB_TRY(cd_sweep(s->hh.f.p + hn.col0, s->k, hn.width, s->hh.ld, hn.num, hn.splits, hn.stride, s->part_den.p,
                         hn.ld_part, s->gw.f.p, static_cast<float>(p.l1_reg_H), static_cast<float>(p.l2_reg_H),
                         h->d_scal.p + 3, s->hh.rowmax.p, h->stream, &mir));
        note_launch(1);
    }
    { NFB_PHASE(h, "finish_h"); NFB_TRY(finish_h(s, hn)); }
    // W rows are sharded; H columns are sharded too when the update was sliced, replicated otherwise
    const bool h_sliced = s->h_sliced;
    if (host_slot != nullptr || violation_out != nullptr) {
        if (!s->p2p) NFB_TRY(allreduce_f64(h, h->d_scal.p + 2, h_sliced ? 2 : 1));   // p2p: summed with the gather
        double* dst = host_slot != nullptr ? host_slot : h->h_scal + 2;
        NFB_CUDA(cudaMemcpyAsync(dst, h->d_scal.p + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (violation_out != nullptr) {
        NFB_CUDA(cudaStreamSynchronize(h->stream));
        *violation_out = h->h_scal[2] + h->h_scal[3];
    }
    return 0;
}

// ||X - WH||_F via  ||X||^2 - 2 <W, X H'> + <W'W, HH'>   (float64 accumulation)
static int nmf_error(nfb_nmf* s, double* err_out) {
    nfb_handle* h = s->h;
    if (!s->p1_valid) NFB_TRY(gemm_xht_with_den(s));
    NFB_TRY(aux_join(s));          // HH' may have been computed on the aux stream
    NFB_CUDA(cudaMemsetAsync(h->d_scal.p, 0, 2 * sizeof(double), h->stream));
    NFB_TRY(dot_partials(s->wt.f.p, s->wt.ld, s->part_w.p, s->s_w, static_cast<int64_t>(s->k) * s->wt.ld, s->wt.ld,
                         s->k, s->x->m, h->d_scal.p, h->stream));
    NFB_TRY(dot_small(s->gw.f.p, s->gh.f.p, s->k, s->gw.ld, h->d_scal.p + 1, h->stream));
    note_launch(2);
    NFB_TRY(allreduce_f64(h, h->d_scal.p, 1));
    NFB_CUDA(cudaMemcpyAsync(h->h_scal, h->d_scal.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    const double sq = s->sumsq_global - 2.0 * h->h_scal[0] + h->h_scal[1];
    *err_out = std::sqrt(std::max(sq, 0.0));
    return 0;
}

static int nmf_prepare_factors(nfb_nmf* s) {
    nfb_handle* h = s->h;
    NFB_TRY(aux_join(s));          // a denominator launched beside the last X H' belongs to the old factors
    s->den_w_valid = false;
    NFB_TRY(refresh_planes(s->wt, s->wt.f.p, nullptr, true, h->stream));
    NFB_TRY(refresh_planes(s->hh, s->hh.f.p, nullptr, true, h->stream));
    NFB_TRY(compute_gram(s, s->wt, s->gw, s->s_gw, true));
    NFB_TRY(compute_gram(s, s->hh, s->gh, s->s_gh, false));
    s->gh_valid = true;
    s->p1_valid = false;
    return 0;
}


// Map every rank's arena and H master into this process (CUDA IPC; the 2 x 64-byte handles are exchanged with
// ncclAllGather so the C ABI needs no help from the host language).  Any failure on any rank -- IPC not
// permitted in the container, no peer access, more than kP2pMaxWorld ranks, NFB_P2P=0 -- leaves p2p off on
// ALL ranks and the NCCL reduce-scatter / all-gather path in charge.
static int p2p_setup(nfb_nmf* s) {
    nfb_handle* h = s->h;
    cudaStream_t st = h->stream;
    const char* env = getenv("NFB_P2P");
    const bool want = h->world <= kP2pMaxWorld && h->all_gather != nullptr && !(env != nullptr && atoi(env) == 0);
    s->arena_gram_off = kP2pHeaderBytes;
    s->arena_rowmax_off = s->arena_gram_off + round_up64(static_cast<int64_t>(h->world) * s->kp * s->kp * 4, 256);
    s->arena_num_off = s->arena_rowmax_off + round_up64(static_cast<int64_t>(h->world) * s->kp * 4, 256);
    const size_t arena_bytes = static_cast<size_t>(s->arena_num_off) +
                               sizeof(float) * static_cast<size_t>(h->world) * s->k * s->slice_cols;
    int failed = want ? 0 : 1;
    cudaIpcMemHandle_t mine[2];
    memset(mine, 0, sizeof(mine));
    if (!failed) {
        if (s->arena.alloc(arena_bytes, true, st) != 0) failed = 1;
    }
    if (!failed) {
        cudaStreamSynchronize(st);   // the zeroed flags must be in place before any peer can post
        if (cudaIpcGetMemHandle(&mine[0], s->arena.p) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine[1], s->hh.f.p) != cudaSuccess) {
            cudaGetLastError();
            failed = 1;
        }
    }
    // exchange handles (every rank takes part, also the ones that already failed)
    DevBuf<char> send, recv;
    NFB_TRY(send.alloc(sizeof(mine), false, st));
    NFB_TRY(recv.alloc(sizeof(mine) * h->world, false, st));
    NFB_CUDA(cudaMemcpyAsync(send.p, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    if (h->all_gather != nullptr) {
        int rc = h->all_gather(send.p, recv.p, sizeof(mine), /*ncclInt8*/ 0, h->comm, st);
        NFB_CHECK(rc == 0, "ncclAllGather(ipc handles) failed with code " + std::to_string(rc));
    }
    std::vector<cudaIpcMemHandle_t> all(2 * h->world);
    NFB_CUDA(cudaMemcpyAsync(all.data(), recv.p, sizeof(mine) * h->world, cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    if (!failed) {
        for (int p = 0; p < h->world && !failed; ++p) {
            if (p == h->rank) {
                s->peers.arena[p] = s->arena.p;
                s->peers.hmaster[p] = s->hh.f.p;
                continue;
            }
            for (int which = 0; which < 2 && !failed; ++which) {
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, all[2 * p + which], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    failed = 1;
                    break;
                }
                s->ipc_opened.push_back(ptr);
                (which == 0 ? s->peers.arena[p] : s->peers.hmaster[p]) = ptr;
            }
        }
    }
    // agree: one failure anywhere turns the path off everywhere
    h->h_scal[14] = failed ? 1.0 : 0.0;
    NFB_CUDA(cudaMemcpyAsync(h->d_scal.p + 14, h->h_scal + 14, sizeof(double), cudaMemcpyHostToDevice, st));
    NFB_TRY(allreduce_f64(h, h->d_scal.p + 14, 1));
    NFB_CUDA(cudaMemcpyAsync(h->h_scal + 14, h->d_scal.p + 14, sizeof(double), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    s->p2p = h->h_scal[14] == 0.0;
    if (!s->p2p) {
        for (void* ptr : s->ipc_opened) cudaIpcCloseMemHandle(ptr);
        s->ipc_opened.clear();
        if (getenv("NFB_P2P_VERBOSE") && !(env != nullptr && atoi(env) == 0))
            fprintf(stderr, "nfact_b200 rank %d: peer-memory path unavailable, using NCCL collectives\n", h->rank);
    }
    return 0;
}

// a wait that timed out (a peer died or fell out of step) poisons the results: tell the caller
static int p2p_check(nfb_nmf* s) {
    if (!s->p2p) return 0;
    int err = 0;
    NFB_CUDA(cudaMemcpyAsync(&err, reinterpret_cast<int*>(s->arena.p) + kP2pError, sizeof(int), cudaMemcpyDeviceToHost,
                             s->h->stream));
    NFB_CUDA(cudaStreamSynchronize(s->h->stream));
    NFB_CHECK(err == 0, "peer-memory exchange timed out waiting for another rank");
    return 0;
}

}  // namespace nfb

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* nfb_last_error(void) { return g_error.c_str(); }
const char* nfb_version(void) { return "nfact_b200 0.1 (sm_100a)"; }

int nfb_create(int device, void* stream, nfb_handle** out) {
    NFB_CHECK(out != nullptr, "out is NULL");
    NFB_CUDA(cudaSetDevice(device));
    int major = 0;
    NFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    NFB_CHECK(major == 10, "nfact_b200 needs a Blackwell (sm_100a) GPU; found compute capability major " +
                               std::to_string(major));
    std::unique_ptr<nfb_handle> h(new nfb_handle());
    h->device = device;
    if (stream != nullptr) {
        h->stream = static_cast<cudaStream_t>(stream);
    } else {
        NFB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    h->cta_group = gemm_default_cta_group();
    h->phase_timing = getenv("NFB_PHASE_TIMING") != nullptr;
    {
        // stream-ordered pool behind DevBuf::alloc_pooled: keep up to half of the device (NFB_POOL_KEEP_GB
        // overrides) cached between calls instead of handing everything back to the driver at every
        // synchronisation -- a per-subject loop then re-uses its plane / staging / solution buffers
        cudaMemPool_t pool = nullptr;
        size_t free_b = 0, total_b = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            // 85 % of the device: the subject loop of nfact_dr keeps two C3-sized subjects (2 x 32 GB of planes and
            // triplets) plus the regression's buffers in flight on two streams; with half the device the pool was
            // trimmed at synchronisations and re-mapping the planes stalled cudaMallocAsync for 0.1 - 0.7 s
            uint64_t keep = total_b / 100 * 85;
            if (getenv("NFB_POOL_KEEP_GB")) keep = static_cast<uint64_t>(atof(getenv("NFB_POOL_KEEP_GB")) * (1ull << 30));
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    NFB_TRY(h->d_scal.alloc(16, true, h->stream));
    NFB_TRY(h->d_bits.alloc(16, true, h->stream));
    NFB_TRY(h->d_int.alloc(16, true, h->stream));
    NFB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_scal), 16 * sizeof(double)));
    NFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    NFB_CUDA(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    NFB_CUDA(cudaEventCreateWithFlags(&h->aux_fork, cudaEventDisableTiming));
    NFB_CUDA(cudaEventCreateWithFlags(&h->aux_join, cudaEventDisableTiming));
    h->den_overlap = !(getenv("NFB_DEN_OVERLAP") && atoi(getenv("NFB_DEN_OVERLAP")) == 0);
    // the Gram GEMMs want 128 CTAs of their own: beside the big launch they cost it as much as they save on one
    // GPU (measured), so only NFB_DEN_OVERLAP=2 moves them as well
    h->gram_overlap = h->den_overlap && getenv("NFB_DEN_OVERLAP") && atoi(getenv("NFB_DEN_OVERLAP")) >= 2;
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    *out = h.release();
    return 0;
}

int nfb_destroy(nfb_handle* h) {
    if (h == nullptr) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->own_comm && h->comm && h->nccl_lib) {
        typedef int (*DestroyFn)(void*);
        DestroyFn fn = reinterpret_cast<DestroyFn>(dlsym(h->nccl_lib, "ncclCommDestroy"));
        if (fn) fn(h->comm);
    }
    h->big.clear();
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
    if (h->aux_fork) cudaEventDestroy(h->aux_fork);
    if (h->aux_join) cudaEventDestroy(h->aux_join);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int nfb_release_cached(nfb_handle* h) {
    NFB_CHECK(h != nullptr, "handle is NULL");
    NFB_CUDA(cudaSetDevice(h->device));
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    h->big.clear();
    // and what the stream-ordered pool keeps cached (its release threshold is most of the device)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    cudaGetLastError();
    return 0;
}

int nfb_synchronize(nfb_handle* h) {
    NFB_CHECK(h != nullptr, "handle is NULL");
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int64_t nfb_launch_count(const nfb_handle* h) { return h ? h->launches_last : 0; }

int nfb_profile_gemms(nfb_handle* h, int enable, double* total_ms_out, int64_t* launches_out) {
    NFB_CHECK(h != nullptr, "handle is NULL");
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    double total = 0.0;
    for (auto& ev : h->gemm_events) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) total += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (total_ms_out) *total_ms_out = total;
    if (launches_out) *launches_out = static_cast<int64_t>(h->gemm_events.size());
    h->gemm_events.clear();
    h->profiling = enable != 0;
    return 0;
}

int nfb_set_comm(nfb_handle* h, void* nccl_comm, void* nccl_allreduce_fn, int rank, int world_size) {
    NFB_CHECK(h != nullptr, "handle is NULL");
    h->comm = nccl_comm;
    h->allreduce = reinterpret_cast<NcclAllReduceFn>(nccl_allreduce_fn);
    h->rank = rank;
    h->world = world_size;
    return 0;
}

struct NcclUniqueIdBlob { char internal[128]; };

int nfb_nccl_unique_id(const char* libnccl_path, void* id_out) {
    void* lib = dlopen(libnccl_path, RTLD_NOW | RTLD_GLOBAL);
    NFB_CHECK(lib != nullptr, std::string("dlopen failed for ") + libnccl_path + ": " + dlerror());
    typedef int (*GetIdFn)(NcclUniqueIdBlob*);
    GetIdFn fn = reinterpret_cast<GetIdFn>(dlsym(lib, "ncclGetUniqueId"));
    NFB_CHECK(fn != nullptr, "ncclGetUniqueId not found");
    int rc = fn(static_cast<NcclUniqueIdBlob*>(id_out));
    NFB_CHECK(rc == 0, "ncclGetUniqueId failed with code " + std::to_string(rc));
    return 0;
}

int nfb_nccl_init_rank(nfb_handle* h, const char* libnccl_path, const void* id, int rank, int world_size) {
    NFB_CHECK(h != nullptr, "handle is NULL");
    NFB_CUDA(cudaSetDevice(h->device));
    void* lib = dlopen(libnccl_path, RTLD_NOW | RTLD_GLOBAL);
    NFB_CHECK(lib != nullptr, std::string("dlopen failed for ") + libnccl_path + ": " + dlerror());
    typedef int (*InitFn)(void**, int, NcclUniqueIdBlob, int);
    InitFn init = reinterpret_cast<InitFn>(dlsym(lib, "ncclCommInitRank"));
    void* ar = dlsym(lib, "ncclAllReduce");
    NFB_CHECK(init != nullptr && ar != nullptr, "NCCL symbols not found");
    NcclUniqueIdBlob blob;
    memcpy(&blob, id, sizeof(blob));
    void* comm = nullptr;
    int rc = init(&comm, world_size, blob, rank);
    NFB_CHECK(rc == 0, "ncclCommInitRank failed with code " + std::to_string(rc));
    h->nccl_lib = lib;
    h->own_comm = true;
    h->reduce_scatter = reinterpret_cast<NcclReduceScatterFn>(dlsym(lib, "ncclReduceScatter"));
    h->all_gather = reinterpret_cast<NcclAllGatherFn>(dlsym(lib, "ncclAllGather"));
    return nfb_set_comm(h, comm, ar, rank, world_size);
}

// ---- resident matrix ---------------------------------------------------------------------------
int nfb_matrix_from_device(nfb_handle* h, const float* x_dev, int64_t m, int64_t n, int64_t ld, nfb_matrix** out) {
    NFB_CHECK(h != nullptr && out != nullptr, "NULL argument");
    NFB_CHECK(m > 0 && n > 0 && ld >= n, "bad matrix shape");
    NFB_CUDA(cudaSetDevice(h->device));
    std::unique_ptr<nfb_matrix> x(new nfb_matrix());
    x->h = h;
    x->m = m;
    x->n = n;
    x->ldp = round_up64(n, 8);
    NFB_TRY(x->hi.alloc_pooled(m * x->ldp, false, h->stream));
    NFB_TRY(x->lo.alloc_pooled(m * x->ldp, false, h->stream));
    NFB_TRY(x->inv_scale.alloc_pooled(1, true, h->stream));
    NFB_CUDA(cudaMemsetAsync(h->d_bits.p, 0, sizeof(unsigned int), h->stream));
    NFB_CUDA(cudaMemsetAsync(h->d_scal.p + 6, 0, 2 * sizeof(double), h->stream));
    NFB_CUDA(cudaMemsetAsync(h->d_int.p + 3, 0, sizeof(int), h->stream));
    NFB_TRY(absmax_f32(x_dev, m, n, ld, h->d_bits.p, h->stream));
    NFB_TRY(split_matrix(x_dev, m, n, ld, h->d_bits.p, x->hi.p, x->lo.p, x->ldp, x->inv_scale.p, h->d_scal.p + 6,
                         h->d_int.p + 3, h->stream));
    NFB_CUDA(cudaMemcpyAsync(h->h_scal + 6, h->d_scal.p + 6, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                             h->stream));
    NFB_CUDA(cudaMemcpyAsync(&x->inv_scale_host, x->inv_scale.p, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaMemcpyAsync(&x->has_negative, h->d_int.p + 3, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    x->sumsq = h->h_scal[6];
    x->sum = h->h_scal[7];
    *out = x.release();
    return 0;
}

int nfb_matrix_from_host(nfb_handle* h, const float* x_host, int64_t m, int64_t n, nfb_matrix** out) {
    NFB_CHECK(h != nullptr && x_host != nullptr && out != nullptr, "NULL argument");
    NFB_CHECK(m > 0 && n > 0, "bad matrix shape");
    NFB_CUDA(cudaSetDevice(h->device));
    // Preferred: one H2D pass into a temporary fp32 copy (max scan + split then run at HBM speed on the
    // device).  When the temporary does not fit, two passes over row chunks through a staging buffer: pass 1
    // finds the global scale, pass 2 splits -- only the planes (4 B/element) ever stay resident.
    std::unique_ptr<nfb_matrix> x(new nfb_matrix());
    x->h = h;
    x->m = m;
    x->n = n;
    x->ldp = round_up64(n, 8);
    NFB_TRY(x->hi.alloc_pooled(m * x->ldp, false, h->stream));
    NFB_TRY(x->lo.alloc_pooled(m * x->ldp, false, h->stream));
    NFB_TRY(x->inv_scale.alloc_pooled(1, true, h->stream));
    NFB_CUDA(cudaMemsetAsync(h->d_bits.p, 0, sizeof(unsigned int), h->stream));
    NFB_CUDA(cudaMemsetAsync(h->d_scal.p + 6, 0, 2 * sizeof(double), h->stream));
    NFB_CUDA(cudaMemsetAsync(h->d_int.p + 3, 0, sizeof(int), h->stream));
    size_t free_b = 0, total_b = 0;
    NFB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t full_bytes = sizeof(float) * static_cast<size_t>(m) * n;
    int rc = 0;
    if (full_bytes + (static_cast<size_t>(2) << 30) < free_b / 2) {
        DevBuf<float> full;
        NFB_TRY(full.alloc_pooled(static_cast<size_t>(m) * n, false, h->stream));
        // 1 GiB pieces keep the driver's pageable staging pipeline busy and let the max scan overlap the copies
        const int64_t piece_rows = std::max<int64_t>(1, std::min<int64_t>(m, (int64_t(1) << 30) / (4 * n)));
        for (int64_t r0 = 0; r0 < m; r0 += piece_rows) {
            const int64_t rows = std::min(piece_rows, m - r0);
            NFB_CUDA(cudaMemcpyAsync(full.p + r0 * n, x_host + r0 * n, sizeof(float) * rows * n,
                                     cudaMemcpyHostToDevice, h->stream));
            NFB_TRY(absmax_f32(full.p + r0 * n, rows, n, n, h->d_bits.p, h->stream));
        }
        NFB_TRY(split_matrix(full.p, m, n, n, h->d_bits.p, x->hi.p, x->lo.p, x->ldp, x->inv_scale.p,
                             h->d_scal.p + 6, h->d_int.p + 3, h->stream));
        NFB_CUDA(cudaStreamSynchronize(h->stream));
    } else {
        const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>(m, (int64_t(1) << 30) / (4 * n)));
        DevBuf<float> stage;
        NFB_TRY(stage.alloc_pooled(chunk_rows * n, false, h->stream));
        for (int pass = 0; pass < 2 && rc == 0; ++pass) {
            for (int64_t r0 = 0; r0 < m && rc == 0; r0 += chunk_rows) {
                const int64_t rows = std::min(chunk_rows, m - r0);
                if (cudaMemcpyAsync(stage.p, x_host + r0 * n, sizeof(float) * rows * n, cudaMemcpyHostToDevice,
                                    h->stream) != cudaSuccess) { set_error("H2D copy of X failed"); rc = 1; break; }
                if (pass == 0) rc = absmax_f32(stage.p, rows, n, n, h->d_bits.p, h->stream);
                else rc = split_matrix(stage.p, rows, n, n, h->d_bits.p, x->hi.p + r0 * x->ldp,
                                       x->lo.p + r0 * x->ldp, x->ldp, x->inv_scale.p, h->d_scal.p + 6,
                                       h->d_int.p + 3, h->stream);
            }
        }
        NFB_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (rc != 0) return rc;
    NFB_CUDA(cudaMemcpyAsync(h->h_scal + 6, h->d_scal.p + 6, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                             h->stream));
    NFB_CUDA(cudaMemcpyAsync(&x->inv_scale_host, x->inv_scale.p, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaMemcpyAsync(&x->has_negative, h->d_int.p + 3, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    x->sumsq = h->h_scal[6];
    x->sum = h->h_scal[7];
    *out = x.release();
    return 0;
}

int nfb_matrix_free(nfb_matrix* x) {
    if (x == nullptr) return 0;
    cudaSetDevice(x->h->device);
    cudaStreamSynchronize(x->h->stream);
    delete x;
    return 0;
}

int nfb_matrix_shape(const nfb_matrix* x, int64_t* m, int64_t* n) {
    NFB_CHECK(x != nullptr, "matrix is NULL");
    if (m) *m = x->m;
    if (n) *n = x->n;
    return 0;
}

int nfb_matrix_sum(const nfb_matrix* x, double* out) {
    NFB_CHECK(x != nullptr && out != nullptr, "NULL argument");
    *out = x->sum;
    return 0;
}

int nfb_matrix_has_negative(const nfb_matrix* x, int* out) {
    NFB_CHECK(x != nullptr && out != nullptr, "NULL argument");
    *out = x->has_negative;
    return 0;
}

int nfb_matrix_sumsq(const nfb_matrix* x, double* out) {
    NFB_CHECK(x != nullptr && out != nullptr, "NULL argument");
    *out = x->sumsq;
    return 0;
}

// ---- NMF -----------------------------------------------------------------------------------------
int nfb_nmf_create(nfb_handle* h, const nfb_matrix* x, int k, const nfb_nmf_params* p, nfb_nmf** out) {
    NFB_CHECK(h != nullptr && x != nullptr && p != nullptr && out != nullptr, "NULL argument");
    NFB_CHECK(k >= 1 && k <= 512, "n_components must be in [1, 512] in this build");
    NFB_CHECK(p->solver == 0 || p->solver == 1, "solver must be 0 (cd) or 1 (mu)");
    NFB_CUDA(cudaSetDevice(h->device));
    std::unique_ptr<nfb_nmf> s(new nfb_nmf());
    s->h = h;
    s->x = x;
    s->k = k;
    s->kp = round_up64(k, 16);
    s->p = *p;
    cudaStream_t st = h->stream;
    // One GPU: the state comes from the stream-ordered pool (a cudaMalloc / cudaFree of these ~1 GB costs 5 - 100 ms
    // of driver time per solve, and NMF-sso creates 21 states); row-sharded runs map the H master into the peers
    // with CUDA IPC, which needs cudaMalloc memory.
    const bool pooled = h->world <= 1;
    auto get = [&](auto& buf, size_t count) {
        return pooled ? buf.alloc_pooled(count, false, st) : buf.alloc(count, false, st);
    };
    NFB_TRY(s->wt.init(k, x->m, true, st, 0, pooled));
    NFB_TRY(s->hh.init(k, x->n, true, st, 0, pooled));
    // the Gram factors share the padded leading dimension kp so the sweep can use float4 rows
    NFB_TRY(s->gw.init(k, k, true, st, s->kp, pooled));
    NFB_TRY(s->gh.init(k, k, true, st, s->kp, pooled));
    NFB_TRY(s->gd.init(k, k, false, st, s->kp, pooled));
    const int cg = h->cta_group;
    // NFB_STREAMK=1: W'X unsplit with a stream-K tail (gemm_split.cu) where its plan pays; default: split-K slabs
    // chosen by the cost model.
    s->use_sk = h->world <= 1 && gemm_streamk_pays(x->n, x->m, cg);
    // X H' keeps its split-K slabs: the cost model fills its waves to 99 % and the unsplit stream-K form measured
    // slower there (5.43 -> 5.57 ms at C3); the stream-K tail is for W'X, whose 430 tiles leave a 0.81 wave
    s->s_w = gemm_pick_splits(x->m, x->n, cg, k);
    // row-sharded: W'X runs unsplit (its tiles go straight to their owners' inboxes on the peer-memory path)
    s->s_h = (h->world > 1 || s->use_sk) ? 1 : gemm_pick_splits(x->n, x->m, cg, k);
    s->s_gw = gemm_pick_splits(k, x->m, cg, k);
    s->s_gh = gemm_pick_splits(k, x->n, cg, k);
    if (s->use_sk) {
        const size_t ws_floats = gemm_streamk_ws_floats(std::min(k, 256));
        const int workers = gemm_streamk_max_workers();
        NFB_TRY(get(s->sk_ws, ws_floats));
        NFB_TRY(pooled ? s->sk_cnt.alloc_pooled(2 * static_cast<size_t>(workers), true, st)
                       : s->sk_cnt.alloc(2 * static_cast<size_t>(workers), true, st));
        s->sk = GemmStreamK{s->sk_ws.p, s->sk_cnt.p, ws_floats, workers, false};
    }
    NFB_TRY(get(s->part_w, static_cast<size_t>(s->s_w) * k * s->wt.ld));
    NFB_TRY(get(s->part_h, static_cast<size_t>(s->s_h) * k * s->hh.ld));
    NFB_TRY(get(s->part_den, static_cast<size_t>(k) * std::max(s->wt.ld, s->hh.ld)));
    NFB_TRY(get(s->part_gram, static_cast<size_t>(std::max(s->s_gw, s->s_gh)) * k * s->kp));
    if (h->world > 1) {
        s->slice_cols = round_up64(ceil_div64(x->n, h->world), 8);
        NFB_TRY(s->num_h.alloc(static_cast<size_t>(k) * std::max<int64_t>(s->hh.ld, s->slice_cols), true, st));
        NFB_TRY(s->num_sl.alloc(static_cast<size_t>(h->world) * k * s->slice_cols, true, st));
        NFB_TRY(s->pack.alloc(static_cast<size_t>(k) * s->slice_cols, true, st));
    }
    if (h->world > 1) NFB_TRY(p2p_setup(s.get()));
    s->h_sliced = h->world > 1 && (s->p2p || (h->reduce_scatter != nullptr && h->all_gather != nullptr));
    // ||X||^2 over all row shards
    h->h_scal[5] = x->sumsq;
    if (h->world > 1) {
        NFB_CUDA(cudaMemcpyAsync(h->d_scal.p + 5, h->h_scal + 5, sizeof(double), cudaMemcpyHostToDevice, st));
        NFB_TRY(allreduce_f64(h, h->d_scal.p + 5, 1));
        NFB_CUDA(cudaMemcpyAsync(h->h_scal + 5, h->d_scal.p + 5, sizeof(double), cudaMemcpyDeviceToHost, st));
        NFB_CUDA(cudaStreamSynchronize(st));
    }
    s->sumsq_global = h->h_scal[5];
    *out = s.release();
    return 0;
}

int nfb_nmf_free(nfb_nmf* s) {
    if (s == nullptr) return 0;
    cudaSetDevice(s->h->device);
    if (s->p2p) {
        // collective: nobody unmaps / frees an arena while a peer's kernels may still touch it
        allreduce_f64(s->h, s->h->d_scal.p + 15, 1);
        cudaStreamSynchronize(s->h->stream);
        for (void* ptr : s->ipc_opened) cudaIpcCloseMemHandle(ptr);
    }
    cudaStreamSynchronize(s->h->stream);
    if (s->h->aux_stream) cudaStreamSynchronize(s->h->aux_stream);
    for (cudaEvent_t ev : s->viol_ready)
        if (ev) cudaEventDestroy(ev);
    delete s;
    return 0;
}

int nfb_nmf_set_factors_dev(nfb_nmf* s, const float* w_dev, const float* h_dev) {
    NFB_CHECK(s != nullptr && w_dev != nullptr && h_dev != nullptr, "NULL argument");
    nfb_handle* h = s->h;
    NFB_CUDA(cudaSetDevice(h->device));
    NFB_TRY(transpose_f32(w_dev, s->x->m, s->k, s->k, s->wt.f.p, s->wt.ld, h->stream));
    NFB_CUDA(cudaMemcpy2DAsync(s->hh.f.p, s->hh.ld * sizeof(float), h_dev, s->x->n * sizeof(float),
                               s->x->n * sizeof(float), s->k, cudaMemcpyDeviceToDevice, h->stream));
    return nmf_prepare_factors(s);
}

int nfb_nmf_set_factors_host(nfb_nmf* s, const float* w_host, const float* h_host) {
    NFB_CHECK(s != nullptr && w_host != nullptr && h_host != nullptr, "NULL argument");
    nfb_handle* h = s->h;
    NFB_CUDA(cudaSetDevice(h->device));
    DevBuf<float> tmp;
    NFB_TRY(tmp.alloc(static_cast<size_t>(s->x->m) * s->k, false, h->stream));
    NFB_CUDA(cudaMemcpyAsync(tmp.p, w_host, sizeof(float) * s->x->m * s->k, cudaMemcpyHostToDevice, h->stream));
    NFB_TRY(transpose_f32(tmp.p, s->x->m, s->k, s->k, s->wt.f.p, s->wt.ld, h->stream));
    NFB_CUDA(cudaMemcpy2DAsync(s->hh.f.p, s->hh.ld * sizeof(float), h_host, s->x->n * sizeof(float),
                               s->x->n * sizeof(float), s->k, cudaMemcpyHostToDevice, h->stream));
    NFB_TRY(nmf_prepare_factors(s));
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int nfb_nmf_get_factors_dev(nfb_nmf* s, float* w_dev, float* h_dev) {
    NFB_CHECK(s != nullptr, "NULL argument");
    nfb_handle* h = s->h;
    NFB_CUDA(cudaSetDevice(h->device));
    if (w_dev) NFB_TRY(transpose_f32(s->wt.f.p, s->k, s->x->m, s->wt.ld, w_dev, s->k, h->stream));
    if (h_dev)
        NFB_CUDA(cudaMemcpy2DAsync(h_dev, s->x->n * sizeof(float), s->hh.f.p, s->hh.ld * sizeof(float),
                                   s->x->n * sizeof(float), s->k, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

int nfb_nmf_get_factors_host(nfb_nmf* s, float* w_host, float* h_host) {
    NFB_CHECK(s != nullptr, "NULL argument");
    nfb_handle* h = s->h;
    NFB_CUDA(cudaSetDevice(h->device));
    DevBuf<float> tmp;
    if (w_host) {
        NFB_TRY(tmp.alloc(static_cast<size_t>(s->x->m) * s->k, false, h->stream));
        NFB_TRY(transpose_f32(s->wt.f.p, s->k, s->x->m, s->wt.ld, tmp.p, s->k, h->stream));
        NFB_CUDA(cudaMemcpyAsync(w_host, tmp.p, sizeof(float) * s->x->m * s->k, cudaMemcpyDeviceToHost, h->stream));
    }
    if (h_host)
        NFB_CUDA(cudaMemcpy2DAsync(h_host, s->x->n * sizeof(float), s->hh.f.p, s->hh.ld * sizeof(float),
                                   s->x->n * sizeof(float), s->k, cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int nfb_nmf_iterate(nfb_nmf* s, int n_iters, double* violation_out) {
    NFB_CHECK(s != nullptr, "NULL argument");
    NFB_CUDA(cudaSetDevice(s->h->device));
    const int64_t before = g_launches.load();
    for (int it = 0; it < n_iters; ++it) {
        if (s->p.solver == 1) {
            NFB_TRY(mu_iteration(s));
            { NFB_PHASE(s->h, "gemm_xht"); NFB_TRY(gemm_xht_with_den(s)); }  // numerator of the next W half-step
        } else {
            NFB_TRY(cd_iteration(s, (it == n_iters - 1) ? violation_out : nullptr));
        }
    }
    s->h->launches_last = g_launches.load() - before;
    report_phases(s->h, n_iters);
    NFB_TRY(p2p_check(s));
    return 0;
}

int nfb_nmf_error(nfb_nmf* s, double* err_out) {
    NFB_CHECK(s != nullptr && err_out != nullptr, "NULL argument");
    NFB_CUDA(cudaSetDevice(s->h->device));
    return nmf_error(s, err_out);
}

int nfb_nmf_exchange_mode(const nfb_nmf* s, int* out) {
    NFB_CHECK(s != nullptr && out != nullptr, "NULL argument");
    *out = s->h->world <= 1 ? 0 : (s->p2p ? 2 : 1);
    return 0;
}

int nfb_nmf_run(nfb_nmf* s, int* n_iter_out, double* recon_err_out) {
    NFB_CHECK(s != nullptr, "NULL argument");
    NFB_CUDA(cudaSetDevice(s->h->device));
    const nfb_nmf_params& p = s->p;
    const int64_t before = g_launches.load();
    int n_iter = 0;
    if (p.solver == 1) {
        // sklearn/decomposition/_nmf.py:823-879
        double error_at_init = 0.0;
        NFB_TRY(nmf_error(s, &error_at_init));
        double previous_error = error_at_init;
        for (n_iter = 1; n_iter <= p.max_iter; ++n_iter) {
            NFB_TRY(mu_iteration(s));
            NFB_TRY(gemm_xht_with_den(s));
            if (p.tol > 0 && n_iter % 10 == 0) {
                double error = 0.0;
                NFB_TRY(p2p_check(s));
                NFB_TRY(nmf_error(s, &error));
                if (p.verbose) printf("Epoch %02d, error: %f\n", n_iter, error);
                if ((previous_error - error) / error_at_init < p.tol) break;
                previous_error = error;
            }
        }
        if (n_iter > p.max_iter) n_iter = p.max_iter;
    } else {
        // sklearn/decomposition/_nmf.py:489-516: stop after the first iteration whose violation ratio is <= tol.
        // Reading the violation back costs a stream synchronisation; doing that after every iteration leaves the
        // GPU idle while the host enqueues the next ~25 launches (0.7 ms of an 11 ms iteration at C3, a third of
        // one at 8 GPUs).  So the loop runs one iteration ahead: iteration i+1 is enqueued before the violation
        // of iteration i is looked at, and the factors after each of the last two iterations are kept
        // (two device copies of W', H per iteration, 0.4 % of its time).  When iteration i turns out to be the
        // last one, the speculative iteration is discarded by restoring that snapshot -- the result is exactly
        // what the one-at-a-time loop returns.
        nfb_handle* h = s->h;
        const size_t w_elems = static_cast<size_t>(s->wt.kp) * s->wt.ld, h_elems = static_cast<size_t>(s->hh.kp) * s->hh.ld;
        for (int b = 0; b < 2; ++b) {
            if (s->snap_w[b].p == nullptr) NFB_TRY(s->snap_w[b].alloc_pooled(w_elems, false, h->stream));
            if (s->snap_h[b].p == nullptr) NFB_TRY(s->snap_h[b].alloc_pooled(h_elems, false, h->stream));
            if (s->viol_ready[b] == nullptr) NFB_CUDA(cudaEventCreateWithFlags(&s->viol_ready[b], cudaEventDisableTiming));
        }
        double violation_init = 0.0;
        int stop_at = 0;
        // returns true when iteration `it` (already executed) satisfies sklearn's stopping rule
        auto decide = [&](int it) -> int {
            const int b = it & 1;
            NFB_CUDA(cudaEventSynchronize(s->viol_ready[b]));
            const double violation = h->h_scal[8 + 2 * b] + h->h_scal[9 + 2 * b];
            if (s->p2p && violation != violation) NFB_TRY(p2p_check(s));   // the exchange poisons it on a time-out
            s->last_violation = violation;
            if (it == 1) violation_init = violation;
            if (violation_init == 0) { stop_at = it; return 0; }
            if (p.verbose) printf("violation: %g\n", violation / violation_init);
            if (violation / violation_init <= p.tol) stop_at = it;
            return 0;
        };
        for (n_iter = 1; n_iter <= p.max_iter; ++n_iter) {
            const int b = n_iter & 1;
            NFB_TRY(cd_iteration(s, nullptr, h->h_scal + 8 + 2 * b));
            NFB_CUDA(cudaEventRecord(s->viol_ready[b], h->stream));
            NFB_CUDA(cudaMemcpyAsync(s->snap_w[b].p, s->wt.f.p, sizeof(float) * w_elems, cudaMemcpyDeviceToDevice, h->stream));
            NFB_CUDA(cudaMemcpyAsync(s->snap_h[b].p, s->hh.f.p, sizeof(float) * h_elems, cudaMemcpyDeviceToDevice, h->stream));
            if (n_iter >= 2) {
                NFB_TRY(decide(n_iter - 1));
                if (stop_at != 0) break;
            }
        }
        if (stop_at == 0) {
            // every enqueued iteration but the last has been examined
            n_iter = std::min(n_iter, p.max_iter);
            if (n_iter >= 1) NFB_TRY(decide(n_iter));
            stop_at = 0;   // the last iteration is the result either way
        } else if (stop_at < n_iter) {
            // iteration stop_at + 1 ran speculatively: put the factors of iteration stop_at back
            const int b = stop_at & 1;
            NFB_CUDA(cudaMemcpyAsync(s->wt.f.p, s->snap_w[b].p, sizeof(float) * w_elems, cudaMemcpyDeviceToDevice, h->stream));
            NFB_CUDA(cudaMemcpyAsync(s->hh.f.p, s->snap_h[b].p, sizeof(float) * h_elems, cudaMemcpyDeviceToDevice, h->stream));
            NFB_TRY(nmf_prepare_factors(s));
            n_iter = stop_at;
        }
    }
    if (n_iter_out) *n_iter_out = n_iter;
    if (recon_err_out) NFB_TRY(nmf_error(s, recon_err_out));
    s->h->launches_last = g_launches.load() - before;
    NFB_TRY(p2p_check(s));
    return 0;
}

int nfb_nmf_fit_host(nfb_handle* h, const nfb_matrix* x, int k, const nfb_nmf_params* p, float* w_host,
                     float* h_host, int* n_iter_out, double* recon_err_out) {
    nfb_nmf* s = nullptr;
    NFB_TRY(nfb_nmf_create(h, x, k, p, &s));
    int rc = nfb_nmf_set_factors_host(s, w_host, h_host);
    if (rc == 0) rc = nfb_nmf_run(s, n_iter_out, recon_err_out);
    if (rc == 0) rc = nfb_nmf_get_factors_host(s, w_host, h_host);
    nfb_nmf_free(s);
    return rc;
}

// ---- dual regression -------------------------------------------------------------------------------
// ---- pinned host memory pool ----------------------------------------------------------------------------
// Result arrays handed to the caller (W, H, the float64 DR outputs) are hundreds of MB; a pageable
// destination makes the D2H copy run at a few GB/s.  The host layer allocates them here instead: page-locked
// blocks, recycled by exact size (cudaHostAlloc itself costs ~0.1 s per GB).
namespace {
std::mutex g_pin_mutex;
std::multimap<size_t, void*> g_pin_free;           // size -> idle block
std::unordered_map<void*, size_t> g_pin_live;      // block -> size
size_t g_pin_idle_bytes = 0;
constexpr size_t kPinIdleCap = size_t(8) << 30;    // idle blocks kept for reuse
}  // namespace

int nfb_host_alloc(int64_t bytes, void** out) {
    NFB_CHECK(out != nullptr && bytes >= 0, "bad nfb_host_alloc arguments");
    const size_t size = std::max<size_t>(static_cast<size_t>(bytes), 64);
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        auto range = g_pin_free.equal_range(size);
        if (range.first != range.second) {
            auto it = std::prev(range.second);         // most recently freed block of this size (its pages are warm)
            *out = it->second;
            g_pin_idle_bytes -= size;
            g_pin_free.erase(it);
            g_pin_live[*out] = size;
            return 0;
        }
    }
    void* p = nullptr;
    NFB_CUDA(cudaHostAlloc(&p, size, cudaHostAllocPortable));
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    g_pin_live[p] = size;
    *out = p;
    return 0;
}

int nfb_host_free(void* p) {
    if (p == nullptr) return 0;
    size_t size = 0;
    bool release = false;
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        auto it = g_pin_live.find(p);
        NFB_CHECK(it != g_pin_live.end(), "nfb_host_free: pointer was not allocated by nfb_host_alloc");
        size = it->second;
        g_pin_live.erase(it);
        if (g_pin_idle_bytes + size <= kPinIdleCap) {
            g_pin_free.emplace(size, p);
            g_pin_idle_bytes += size;
        } else {
            release = true;
        }
    }
    if (release) cudaFreeHost(p);
    return 0;
}

// The group grey components prepared for stage 1 of the dual regression: component-major float32 master + fp16
// planes (the operand of G'X) and the float64 Gram matrix G'G.  Built once per call by the *_host entry points, or
// once per subject loop by nfb_dr_components_create (the group maps are the same for every subject).
struct nfb_dr_components {
    nfb_handle* h = nullptr;
    int64_t m = 0;
    int k = 0;
    int nonfinite = 0;
    Factor gt;
    DevBuf<double> gram64;
};

static int dr_components_build(nfb_handle* h, const void* g_host, int g_is_f32, int64_t m, int k,
                               nfb_dr_components* c) {
    cudaStream_t st = h->stream;
    c->h = h;
    c->m = m;
    c->k = k;
    NFB_TRY(c->gt.init(k, m, true, st, 0, true));
    NFB_TRY(c->gram64.alloc_pooled(static_cast<size_t>(k) * k, false, st));
    DevBuf<double> g64;
    NFB_TRY(g64.alloc_pooled(static_cast<size_t>(m) * k, false, st));
    NFB_CUDA(cudaMemsetAsync(h->d_int.p + 2, 0, sizeof(int), st));
    // group grey components: [m, k] float64 or float32 -> component-major float32 master + planes
    // (scipy's nnls casts to float64; the product G'X runs on the float32 value either way)
    { NFB_PHASE(h, "g_upload");
    if (g_is_f32) {
        float* g32 = reinterpret_cast<float*>(g64.p);
        NFB_CUDA(cudaMemcpyAsync(g32, g_host, sizeof(float) * m * k, cudaMemcpyHostToDevice, st));
        NFB_TRY(transpose_f32(g32, m, k, k, c->gt.f.p, c->gt.ld, st, h->d_int.p + 2));
    } else {
        NFB_CUDA(cudaMemcpyAsync(g64.p, g_host, sizeof(double) * m * k, cudaMemcpyHostToDevice, st));
        NFB_TRY(transpose_f64_to_f32(g64.p, m, k, k, c->gt.f.p, c->gt.ld, st, h->d_int.p + 2));
    }
    NFB_TRY(refresh_planes(c->gt, c->gt.f.p, nullptr, true, st));
    }
    { NFB_PHASE(h, "gram1"); NFB_TRY(gram_f64(c->gt.f.p, k, m, c->gt.ld, c->gram64.p, st)); }
    NFB_CUDA(cudaMemcpyAsync(&c->nonfinite, h->d_int.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));      // g64 (and the pageable host source) are done with
    return 0;
}

static int dr_solve(nfb_handle* h, const nfb_matrix* x, nfb_dr_components& comps, double zscore,
                    const int32_t* seed_id_host, int n_seeds, double* gs_out, double* hs_out, double* gs_norm_out,
                    double* hs_norm_out, int* n_unconverged_out);

int nfb_dr_components_create(nfb_handle* h, const void* g_host, int g_is_f32, int64_t m, int k,
                             nfb_dr_components** out, int* g_nonfinite_out) {
    NFB_CHECK(h != nullptr && g_host != nullptr && out != nullptr, "NULL argument");
    NFB_CHECK(m >= 1 && k >= 1 && k <= 512, "n_components must be in [1, 512] in this build");
    NFB_CUDA(cudaSetDevice(h->device));
    std::unique_ptr<nfb_dr_components> c(new nfb_dr_components());
    NFB_TRY(dr_components_build(h, g_host, g_is_f32, m, k, c.get()));
    if (g_nonfinite_out) *g_nonfinite_out = c->nonfinite;
    *out = c.release();
    return 0;
}

int nfb_dr_components_free(nfb_dr_components* c) {
    delete c;
    return 0;
}

int nfb_dual_regression_resident_host(nfb_handle* h, const nfb_matrix* x, nfb_dr_components* comps, double zscore,
                                      const int32_t* seed_id_host, int n_seeds, double* gs_out, double* hs_out,
                                      double* gs_norm_out, double* hs_norm_out, int* n_unconverged_out) {
    NFB_CHECK(h != nullptr && x != nullptr && comps != nullptr, "NULL argument");
    NFB_CHECK(comps->h == h, "the components were prepared on another handle");
    NFB_CHECK(comps->m == x->m, "Incompatible dimensions: the components have " + std::to_string(comps->m) +
                                    " rows, the matrix " + std::to_string(x->m));
    NFB_CUDA(cudaSetDevice(h->device));
    return dr_solve(h, x, *comps, zscore, seed_id_host, n_seeds, gs_out, hs_out, gs_norm_out, hs_norm_out,
                    n_unconverged_out);
}

int nfb_dual_regression_host(nfb_handle* h, const nfb_matrix* x, int k, const void* g_host, int g_is_f32,
                             double* gs_out, double* hs_out, int* n_unconverged_out, int* g_nonfinite_out) {
    return nfb_dual_regression_post_host(h, x, k, g_host, g_is_f32, 0.0, nullptr, 0, gs_out, hs_out, nullptr, nullptr,
                                         n_unconverged_out, g_nonfinite_out);
}

int nfb_dual_regression_post_host(nfb_handle* h, const nfb_matrix* x, int k, const void* g_host, int g_is_f32,
                                  double zscore, const int32_t* seed_id_host, int n_seeds, double* gs_out,
                                  double* hs_out, double* gs_norm_out, double* hs_norm_out, int* n_unconverged_out,
                                  int* g_nonfinite_out) {
    NFB_CHECK(h != nullptr && x != nullptr && g_host != nullptr, "NULL argument");
    NFB_CHECK(k >= 1 && k <= 512, "n_components must be in [1, 512] in this build");
    NFB_CUDA(cudaSetDevice(h->device));
    nfb_dr_components comps;
    NFB_TRY(dr_components_build(h, g_host, g_is_f32, x->m, k, &comps));
    if (g_nonfinite_out) *g_nonfinite_out = comps.nonfinite;
    return dr_solve(h, x, comps, zscore, seed_id_host, n_seeds, gs_out, hs_out, gs_norm_out, hs_norm_out,
                    n_unconverged_out);
}

static int dr_solve(nfb_handle* h, const nfb_matrix* x, nfb_dr_components& comps, double zscore,
                    const int32_t* seed_id_host, int n_seeds, double* gs_out, double* hs_out, double* gs_norm_out,
                    double* hs_norm_out, int* n_unconverged_out) {
    NFB_CHECK(zscore == 0.0 || (seed_id_host != nullptr && n_seeds >= 1 && n_seeds <= 4096),
              "thresholding needs the seed index of every row and 1..4096 seeds");
    cudaStream_t st = h->stream;
    const int k = comps.k;
    const int64_t m = x->m, n = x->n, kp = round_up64(k, 16);
    const int cg = h->cta_group;
    Factor& gt = comps.gt;
    Factor hs;
    NFB_TRY(hs.init(k, n, true, st, 0, true));
    DevBuf<double> g64, gram64, sol_h, sol_g;
    // g64 receives the transposed stage-2 solution [m, k] at the end; gram64 is stage 2's Gram matrix
    NFB_TRY(g64.alloc_pooled(static_cast<size_t>(m) * k, false, st));
    NFB_TRY(gram64.alloc_pooled(static_cast<size_t>(k) * k, false, st));
    NFB_TRY(sol_h.alloc_pooled(static_cast<size_t>(k) * hs.ld, true, st));
    NFB_TRY(sol_g.alloc_pooled(static_cast<size_t>(k) * gt.ld, true, st));
    int s1 = gemm_pick_splits(n, m, cg, k);
    const int s2 = gemm_pick_splits(m, n, cg, k);
    // stage 1's G'X is a W'X-type launch: unsplit with the stream-K tail (see nfb_nmf_create)
    DevBuf<float> sk_ws;
    DevBuf<unsigned int> sk_cnt;
    GemmStreamK sk{};
    const bool use_sk = gemm_streamk_pays(n, m, cg);
    if (use_sk) {
        const size_t ws_floats = gemm_streamk_ws_floats(std::min(k, 256));
        NFB_TRY(sk_ws.alloc_pooled(ws_floats, false, st));
        NFB_TRY(sk_cnt.alloc_pooled(2 * static_cast<size_t>(gemm_streamk_max_workers()), true, st));
        sk = GemmStreamK{sk_ws.p, sk_cnt.p, ws_floats, gemm_streamk_max_workers(), false};
        s1 = 1;
    }
    DevBuf<float> part;
    NFB_TRY(part.alloc_pooled(std::max(static_cast<size_t>(s1) * k * hs.ld, static_cast<size_t>(s2) * k * gt.ld), false, st));
    NFB_CUDA(cudaMemsetAsync(h->d_int.p, 0, 2 * sizeof(int), st));
    NFB_CUDA(cudaMemsetAsync(h->d_scal.p + 8, 0, 2 * sizeof(double), st));
    // ---- stage 1: white matter.  Hs[:, j] = nnls(G, X[:, j])  (dual_regression_methods.py:137-149)
    { NFB_PHASE(h, "gemm_gtx");
    NFB_TRY(gemm_split(x->operand(), true, gt.operand(), n, m, static_cast<int>(kp), k, part.p, hs.ld, s1,
                       gt.inv_scale.p, nullptr, x->inv_scale_host, cg, st, nullptr, use_sk ? &sk : nullptr));
    }
    { NFB_PHASE(h, "nnls1");
    NFB_TRY(nnls_gram_batched(comps.gram64.p, k, part.p, s1, static_cast<int64_t>(k) * hs.ld, hs.ld, n, sol_h.p, hs.ld,
                              400, h->d_int.p, st, reinterpret_cast<unsigned long long*>(h->d_scal.p + 8),
                              reinterpret_cast<unsigned long long*>(h->d_scal.p + 10)));
    }
    unsigned long long flagged_cols[2] = {0, 0};   // NFB_DR_STATS: columns the greedy NNLS kernel left to the clean-up pass
    if (getenv("NFB_DR_STATS")) cudaMemcpy(&flagged_cols[0], h->d_scal.p + 11, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    // ---- stage 2 operand: the (un-thresholded) white components as float32 planes ----------------------
    { NFB_PHASE(h, "planes2");
    NFB_CUDA(cudaMemsetAsync(hs.rowmax.p, 0, sizeof(unsigned int) * kp, st));
    NFB_TRY(f64_to_f32_rowmax(sol_h.p, k, n, hs.ld, hs.f.p, hs.ld, hs.rowmax.p, st));
    NFB_TRY(refresh_planes(hs, hs.f.p, nullptr, false, st));
    }
    // post-processing of the white components (thresholding_components / normalise_components,
    // NFACT/base/matrix_handling.py:241-277, 45-68) while they are still resident; sol_h is not read again
    const bool post = zscore != 0.0;
    DevBuf<double> stats, norm;
    DevBuf<int> seed_dev;
    if (post || hs_norm_out || gs_norm_out) {
        NFB_TRY(stats.alloc_pooled(static_cast<size_t>(4) * k * std::max(1, n_seeds), false, st));
        if (hs_norm_out || gs_norm_out) NFB_TRY(norm.alloc_pooled(static_cast<size_t>(k) * std::max(hs.ld, gt.ld), false, st));
    }
    { NFB_PHASE(h, "post_h");
    if (post) NFB_TRY(comp_threshold<double>(sol_h.p, k, n, hs.ld, nullptr, 1, zscore, stats.p, st));
    if (hs_norm_out) NFB_TRY(comp_standardize<double>(sol_h.p, k, n, hs.ld, norm.p, hs.ld, stats.p, st));
    }
    // the white components can travel while stage 2 runs
    SideStreamEvent hs_guard(h->copy_stream);
    cudaEvent_t& hs_ready = hs_guard.ev;
    if (hs_out || hs_norm_out) {
        NFB_CHECK(hs_guard.create() == 0, "cudaEventCreate failed");
        NFB_CUDA(cudaEventRecord(hs_ready, st));
        NFB_CUDA(cudaStreamWaitEvent(h->copy_stream, hs_ready, 0));
        if (hs_out)
            NFB_CUDA(cudaMemcpy2DAsync(hs_out, n * sizeof(double), sol_h.p, hs.ld * sizeof(double), n * sizeof(double),
                                       k, cudaMemcpyDeviceToHost, h->copy_stream));
        if (hs_norm_out) {
            NFB_CUDA(cudaMemcpy2DAsync(hs_norm_out, n * sizeof(double), norm.p, hs.ld * sizeof(double),
                                       n * sizeof(double), k, cudaMemcpyDeviceToHost, h->copy_stream));
            // the scratch is reused for the grey components: stage 2's post-processing waits for this copy
            NFB_CUDA(cudaEventRecord(hs_ready, h->copy_stream));
        }
    }
    // ---- stage 2: grey matter.  Gs[i, :] = nnls(Hs, X[i, :])    (dual_regression_methods.py:151-163)
    { NFB_PHASE(h, "gram2"); NFB_TRY(gram_f64(hs.f.p, k, n, hs.ld, gram64.p, st)); }
    { NFB_PHASE(h, "gemm_xhst");
    NFB_TRY(gemm_split(x->operand(), false, hs.operand(), m, n, static_cast<int>(kp), k, part.p, gt.ld, s2,
                       hs.inv_scale.p, nullptr, x->inv_scale_host, cg, st));
    }
    { NFB_PHASE(h, "nnls2");
    NFB_TRY(nnls_gram_batched(gram64.p, k, part.p, s2, static_cast<int64_t>(k) * gt.ld, gt.ld, m, sol_g.p, gt.ld,
                              400, h->d_int.p + 1, st, reinterpret_cast<unsigned long long*>(h->d_scal.p + 9),
                              reinterpret_cast<unsigned long long*>(h->d_scal.p + 10)));
    }
    // ---- results -------------------------------------------------------------------------------------
    { NFB_PHASE(h, "post_g");
    if (post) {
        // threshold_grey_components: statistics per (component, seed); rows of no listed seed stay as they are
        NFB_TRY(seed_dev.alloc_pooled(static_cast<size_t>(m), false, st));
        NFB_CUDA(cudaMemcpyAsync(seed_dev.p, seed_id_host, sizeof(int) * m, cudaMemcpyHostToDevice, st));
        NFB_TRY(comp_threshold<double>(sol_g.p, k, m, gt.ld, seed_dev.p, n_seeds, zscore, stats.p, st));
    }
    }
    { NFB_PHASE(h, "d2h");
    if (gs_out) {
        NFB_TRY(transpose_f64(sol_g.p, k, m, gt.ld, g64.p, k, st));
        NFB_CUDA(cudaMemcpyAsync(gs_out, g64.p, sizeof(double) * m * k, cudaMemcpyDeviceToHost, st));
    }
    if (gs_norm_out) {
        if (hs_norm_out) NFB_CUDA(cudaStreamWaitEvent(st, hs_ready, 0));
        NFB_TRY(comp_standardize<double>(sol_g.p, k, m, gt.ld, norm.p, gt.ld, stats.p, st));
        NFB_TRY(transpose_f64(norm.p, k, m, gt.ld, g64.p, k, st));   // ordered after the gs_out copy on the stream
        NFB_CUDA(cudaMemcpyAsync(gs_norm_out, g64.p, sizeof(double) * m * k, cudaMemcpyDeviceToHost, st));
    }
    }
    int status[2] = {0, 0};
    NFB_CUDA(cudaMemcpyAsync(status, h->d_int.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    hs_guard.finish();
    NFB_CUDA(cudaGetLastError());
    if (n_unconverged_out) *n_unconverged_out = status[0] + status[1];
    report_phases(h, 1);
    if (getenv("NFB_DR_STATS")) {
        unsigned long long steps[2] = {0, 0};
        cudaMemcpy(steps, h->d_scal.p + 8, sizeof(steps), cudaMemcpyDeviceToHost);
        cudaMemcpy(&flagged_cols[1], h->d_scal.p + 11, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        fprintf(stderr, "nfact_b200 DR: stage1 %.1f coordinate steps/column (n=%lld, %llu left to the clean-up pass), "
                "stage2 %.1f steps/row (m=%lld, %llu left to the clean-up pass), k=%d\n",
                static_cast<double>(steps[0]) / n, static_cast<long long>(n), flagged_cols[0],
                static_cast<double>(steps[1]) / m, static_cast<long long>(m), flagged_cols[1], k);
    }
    return 0;
}

// ---- post-processing of component maps given as host arrays ------------------------------------------
}  // extern "C"

template <typename T>
static int component_postops(nfb_handle* h, T* data, int comp_major, int64_t k, int64_t len, const int32_t* group_id,
                             int n_groups, int do_threshold, double zscore, T* norm_out) {
    cudaStream_t st = h->stream;
    const int64_t ld = round_up64(len, 8);
    const int groups = group_id ? n_groups : 1;
    DevBuf<T> raw, cm, norm;
    DevBuf<double> stats;
    DevBuf<int> gid;
    NFB_TRY(cm.alloc_pooled(static_cast<size_t>(k) * ld, false, st));
    NFB_TRY(stats.alloc_pooled(static_cast<size_t>(4) * k * groups, false, st));
    auto transpose = [&](const T* in, int64_t rows, int64_t cols, int64_t ld_in, T* out, int64_t ld_out) -> int {
        if constexpr (sizeof(T) == 8) return transpose_f64(in, rows, cols, ld_in, out, ld_out, st);
        else return transpose_f32(in, rows, cols, ld_in, out, ld_out, st);
    };
    if (comp_major) {
        NFB_CUDA(cudaMemcpy2DAsync(cm.p, ld * sizeof(T), data, len * sizeof(T), len * sizeof(T), k,
                                   cudaMemcpyHostToDevice, st));
    } else {   // [len, k] on the host (grey components): component-major on the device
        NFB_TRY(raw.alloc_pooled(static_cast<size_t>(k) * len, false, st));
        NFB_CUDA(cudaMemcpyAsync(raw.p, data, sizeof(T) * k * len, cudaMemcpyHostToDevice, st));
        NFB_TRY(transpose(raw.p, len, k, k, cm.p, ld));
    }
    if (do_threshold) {
        if (group_id) {
            NFB_TRY(gid.alloc_pooled(static_cast<size_t>(len), false, st));
            NFB_CUDA(cudaMemcpyAsync(gid.p, group_id, sizeof(int) * len, cudaMemcpyHostToDevice, st));
        }
        NFB_TRY(comp_threshold<T>(cm.p, k, len, ld, group_id ? gid.p : nullptr, groups, zscore, stats.p, st));
        if (comp_major) {
            NFB_CUDA(cudaMemcpy2DAsync(data, len * sizeof(T), cm.p, ld * sizeof(T), len * sizeof(T), k,
                                       cudaMemcpyDeviceToHost, st));
        } else {
            NFB_TRY(transpose(cm.p, k, len, ld, raw.p, k));
            NFB_CUDA(cudaMemcpyAsync(data, raw.p, sizeof(T) * k * len, cudaMemcpyDeviceToHost, st));
        }
    }
    if (norm_out) {
        NFB_TRY(norm.alloc_pooled(static_cast<size_t>(k) * ld, false, st));
        NFB_TRY(comp_standardize<T>(cm.p, k, len, ld, norm.p, ld, stats.p, st));
        if (comp_major) {
            NFB_CUDA(cudaMemcpy2DAsync(norm_out, len * sizeof(T), norm.p, ld * sizeof(T), len * sizeof(T), k,
                                       cudaMemcpyDeviceToHost, st));
        } else {
            NFB_TRY(transpose(norm.p, k, len, ld, raw.p, k));
            NFB_CUDA(cudaMemcpyAsync(norm_out, raw.p, sizeof(T) * k * len, cudaMemcpyDeviceToHost, st));
        }
    }
    NFB_CUDA(cudaStreamSynchronize(st));
    note_launch(4);
    return 0;
}

template <typename T>
static int component_loadings(nfb_handle* h, const T* subject, const T* group, int comp_major, int64_t k, int64_t len,
                              double* out_host) {
    cudaStream_t st = h->stream;
    const int64_t ld = round_up64(len, 8);
    DevBuf<T> raw, cm[2];
    DevBuf<double> out;
    NFB_TRY(out.alloc_pooled(static_cast<size_t>(k), false, st));
    if (!comp_major) NFB_TRY(raw.alloc_pooled(static_cast<size_t>(k) * len, false, st));
    const T* src[2] = {subject, group};
    for (int a = 0; a < 2; ++a) {
        NFB_TRY(cm[a].alloc_pooled(static_cast<size_t>(k) * ld, false, st));
        if (comp_major) {
            NFB_CUDA(cudaMemcpy2DAsync(cm[a].p, ld * sizeof(T), src[a], len * sizeof(T), len * sizeof(T), k,
                                       cudaMemcpyHostToDevice, st));
        } else {
            NFB_CUDA(cudaMemcpyAsync(raw.p, src[a], sizeof(T) * k * len, cudaMemcpyHostToDevice, st));
            if constexpr (sizeof(T) == 8) NFB_TRY(transpose_f64(raw.p, len, k, k, cm[a].p, ld, st));
            else NFB_TRY(transpose_f32(raw.p, len, k, k, cm[a].p, ld, st));
        }
    }
    NFB_TRY(comp_loadings<T>(cm[0].p, cm[1].p, k, len, ld, out.p, st));
    NFB_CUDA(cudaMemcpyAsync(out_host, out.p, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    note_launch(1);
    return 0;
}

template <typename T>
static int component_wta(nfb_handle* h, const T* data, int comp_major, int64_t k, int64_t len, double z_thr,
                         int64_t* wta_out) {
    cudaStream_t st = h->stream;
    const int64_t ld = round_up64(len, 8);
    DevBuf<T> raw, cm;
    DevBuf<double> stats;
    DevBuf<long long> out;
    NFB_TRY(cm.alloc_pooled(static_cast<size_t>(k) * ld, false, st));
    NFB_TRY(stats.alloc_pooled(static_cast<size_t>(4) * k, false, st));
    NFB_TRY(out.alloc_pooled(static_cast<size_t>(len), false, st));
    if (comp_major) {
        NFB_CUDA(cudaMemcpy2DAsync(cm.p, ld * sizeof(T), data, len * sizeof(T), len * sizeof(T), k,
                                   cudaMemcpyHostToDevice, st));
    } else {   // [len, k] on the host (grey components): component-major on the device
        NFB_TRY(raw.alloc_pooled(static_cast<size_t>(k) * len, false, st));
        NFB_CUDA(cudaMemcpyAsync(raw.p, data, sizeof(T) * k * len, cudaMemcpyHostToDevice, st));
        if constexpr (sizeof(T) == 8) NFB_TRY(transpose_f64(raw.p, len, k, k, cm.p, ld, st));
        else NFB_TRY(transpose_f32(raw.p, len, k, k, cm.p, ld, st));
    }
    NFB_TRY(comp_wta<T>(cm.p, k, len, ld, comp_major ? 0 : 1, z_thr, stats.p, out.p, st));
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 output");
    NFB_CUDA(cudaMemcpyAsync(wta_out, out.p, sizeof(int64_t) * len, cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    note_launch(comp_major ? 1 : 3);
    return 0;
}

extern "C" {

int nfb_component_postops_host(nfb_handle* h, void* data, int is_f64, int comp_major, int64_t k, int64_t len,
                               const int32_t* group_id, int n_groups, int do_threshold, double zscore,
                               void* norm_out) {
    NFB_CHECK(h != nullptr && (data != nullptr || k * len == 0) && k >= 0 && len >= 0, "bad nfb_component_postops_host arguments");
    NFB_CHECK(group_id == nullptr || (n_groups >= 1 && n_groups <= 4096), "n_groups must be in [1, 4096]");
    NFB_CUDA(cudaSetDevice(h->device));
    if (k * len == 0) return 0;
    if (is_f64)
        return component_postops<double>(h, static_cast<double*>(data), comp_major, k, len, group_id, n_groups,
                                         do_threshold, zscore, static_cast<double*>(norm_out));
    return component_postops<float>(h, static_cast<float*>(data), comp_major, k, len, group_id, n_groups, do_threshold,
                                    zscore, static_cast<float*>(norm_out));
}

int nfb_component_wta_host(nfb_handle* h, const void* data, int is_f64, int comp_major, int64_t k, int64_t len,
                           double z_thr, int64_t* wta_out) {
    NFB_CHECK(h != nullptr && data != nullptr && wta_out != nullptr, "NULL argument");
    NFB_CHECK(k >= 1 && len >= 1, "bad nfb_component_wta_host shape");
    NFB_CUDA(cudaSetDevice(h->device));
    if (is_f64)
        return component_wta<double>(h, static_cast<const double*>(data), comp_major, k, len, z_thr, wta_out);
    return component_wta<float>(h, static_cast<const float*>(data), comp_major, k, len, z_thr, wta_out);
}

// NFB_HOST_TIMING=1: wall-clock marks of the host-visible phases of the ingest (each mark synchronises the stream)
struct HostMarks {
    cudaStream_t st;
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string line;
    explicit HostMarks(cudaStream_t stream) : st(stream), on(getenv("NFB_HOST_TIMING") != nullptr) {
        if (on) t0 = std::chrono::steady_clock::now();
    }
    void mark(const char* name) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof(buf), "%s=%.1f ms  ", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
        line += buf;
        t0 = t1;
    }
    ~HostMarks() {
        if (on && !line.empty()) fprintf(stderr, "nfact_b200 ingest (host clock): %s\n", line.c_str());
    }
};

int nfb_component_loadings_host(nfb_handle* h, const void* subject, const void* group, int is_f64, int comp_major,
                                int64_t k, int64_t len, double* loadings_out) {
    NFB_CHECK(h != nullptr && subject != nullptr && group != nullptr && loadings_out != nullptr, "NULL argument");
    NFB_CHECK(k >= 1 && len >= 1, "bad nfb_component_loadings_host shape");
    NFB_CUDA(cudaSetDevice(h->device));
    if (is_f64)
        return component_loadings<double>(h, static_cast<const double*>(subject), static_cast<const double*>(group),
                                          comp_major, k, len, loadings_out);
    return component_loadings<float>(h, static_cast<const float*>(subject), static_cast<const float*>(group),
                                     comp_major, k, len, loadings_out);
}

// ---- ingest ----------------------------------------------------------------------------------------
int nfb_ingest_coo(nfb_handle* h, int n_subjects, const int32_t* const* rows_host, const int32_t* const* cols_host,
                   const double* const* vals_host, const int64_t* nnz, const double* scale, int64_t nrows,
                   int64_t ncols, int group_mode, float* x_out_host, float* x_out_dev) {
    NFB_CHECK(h != nullptr && n_subjects >= 1 && nrows > 0 && ncols > 0, "bad ingest arguments");
    NFB_CHECK(group_mode || n_subjects == 1, "single-subject mode takes exactly one subject");
    NFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    int64_t max_nnz = 0;
    for (int s = 0; s < n_subjects; ++s) max_nnz = std::max(max_nnz, nnz[s]);
    size_t free_b = 0, total_b = 0;
    NFB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double budget = 0.6 * static_cast<double>(free_b) - 16.0 * static_cast<double>(max_nnz);
    // group mode: scratch + accumulator (fp64) per cell; single subject: X = float32(scratch * scale) needs no
    // accumulator (0 + d * scale == d * scale), so a C3-sized subject is ingested in one row block
    const bool single = !group_mode;
    const double per_row = ((single ? 8.0 : 16.0) + (x_out_dev == nullptr ? 4.0 : 0.0)) * static_cast<double>(ncols);
    NFB_CHECK(budget > per_row, "not enough device memory for ingest");
    const int64_t blk_rows = std::max<int64_t>(1, std::min<int64_t>(nrows, static_cast<int64_t>(budget / per_row)));
    HostMarks marks(st);
    DevBuf<double> tmp, acc, vals;
    DevBuf<int> rows, cols;
    DevBuf<float> xblk;
    NFB_TRY(tmp.alloc_pooled(static_cast<size_t>(blk_rows) * ncols, true, st));
    if (!single) NFB_TRY(acc.alloc_pooled(static_cast<size_t>(blk_rows) * ncols, false, st));
    if (x_out_dev == nullptr) NFB_TRY(xblk.alloc_pooled(static_cast<size_t>(blk_rows) * ncols, false, st));
    NFB_TRY(rows.alloc_pooled(max_nnz, false, st));
    NFB_TRY(cols.alloc_pooled(max_nnz, false, st));
    NFB_TRY(vals.alloc_pooled(max_nnz, false, st));
    NFB_CUDA(cudaMemsetAsync(h->d_int.p + 2, 0, sizeof(int), st));
    marks.mark("alloc+zero");
    const double recip = 1.0 / static_cast<double>(n_subjects);
    auto upload = [&](int s) -> int {
        NFB_CUDA(cudaMemcpyAsync(rows.p, rows_host[s], sizeof(int) * nnz[s], cudaMemcpyHostToDevice, st));
        NFB_CUDA(cudaMemcpyAsync(cols.p, cols_host[s], sizeof(int) * nnz[s], cudaMemcpyHostToDevice, st));
        NFB_CUDA(cudaMemcpyAsync(vals.p, vals_host[s], sizeof(double) * nnz[s], cudaMemcpyHostToDevice, st));
        return 0;
    };
    if (n_subjects == 1) NFB_TRY(upload(0));   // one subject: its triplets are uploaded once for all row blocks
    marks.mark("upload");
    for (int64_t r0 = 0; r0 < nrows; r0 += blk_rows) {
        const int64_t nr = std::min(blk_rows, nrows - r0);
        if (!single) NFB_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * nr * ncols, st));
        if (r0 > 0 && single) NFB_CUDA(cudaMemsetAsync(tmp.p, 0, sizeof(double) * nr * ncols, st));
        for (int s = 0; s < n_subjects; ++s) {
            // the staging buffers are reused by the next subject: pageable H2D copies are synchronous
            // with respect to the host, the kernels are ordered on the stream.
            if (n_subjects > 1) NFB_TRY(upload(s));
            NFB_TRY(ingest_scatter(rows.p, cols.p, vals.p, nnz[s], r0, nr, nrows, ncols, scale[s], tmp.p,
                                   single ? nullptr : acc.p, h->d_int.p + 2, st));
        }
        float* dst = x_out_dev ? x_out_dev + r0 * ncols : xblk.p;
        if (single) NFB_TRY(ingest_finalize(tmp.p, nr * ncols, scale[0], 1, dst, st));
        else NFB_TRY(ingest_finalize(acc.p, nr * ncols, recip, group_mode ? 1 : 0, dst, st));
        if (x_out_host)
            NFB_CUDA(cudaMemcpyAsync(x_out_host + r0 * ncols, dst, sizeof(float) * nr * ncols, cudaMemcpyDeviceToHost,
                                     st));
    }
    int oob = 0;
    NFB_CUDA(cudaMemcpyAsync(&oob, h->d_int.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    marks.mark("scatter+finalize");
    NFB_CHECK(oob == 0, "row/column index exceeds matrix dimensions");
    tmp.release(); acc.release(); vals.release(); rows.release(); cols.release(); xblk.release();
    marks.mark("free");
    return 0;
}

// One subject whose triplets are all distinct: straight into the operand planes (ingest.cu), no dense float64
// scratch and no dense float32 matrix.  *duplicates = true (and no matrix) when a coordinate occurs twice.
// rows_host == nullptr: the triplets are already on the device in rows_dev / cols_dev / vals_dev (the device .dot
// parser); otherwise they are uploaded from the host arrays.
static int ingest_single_to_planes(nfb_handle* h, const int32_t* rows_host, const int32_t* cols_host,
                                   const double* vals_host, const int* rows_dev, const int* cols_dev,
                                   const double* vals_dev, int64_t nnz, double scale, int64_t nrows, int64_t ncols,
                                   nfb_matrix** out, bool* duplicates) {
    cudaStream_t st = h->stream;
    HostMarks marks(st);
    std::unique_ptr<nfb_matrix> x(new nfb_matrix());
    x->h = h;
    x->m = nrows;
    x->n = ncols;
    x->ldp = round_up64(ncols, 8);
    const size_t cells = static_cast<size_t>(nrows) * x->ldp;
    const bool upload = rows_host != nullptr;
    DevBuf<unsigned int> mask;
    DevBuf<int> rows, cols;
    DevBuf<double> vals;
    NFB_TRY(x->hi.alloc_cached(&h->big, cells, st));
    NFB_TRY(x->lo.alloc_cached(&h->big, cells, st));
    NFB_TRY(x->inv_scale.alloc_pooled(1, true, st));
    NFB_TRY(mask.alloc_cached(&h->big, (cells + 31) / 32, st));
    SideStreamEvent up_guard(h->copy_stream);
    cudaEvent_t& uploaded = up_guard.ev;
    if (upload) {
        // staging sized in 64 Mi-triplet steps so that subjects with slightly different counts share the cache
        const size_t nnz_cap = static_cast<size_t>(round_up64(std::max<int64_t>(nnz, 1), int64_t(1) << 26));
        NFB_TRY(rows.alloc_cached(&h->big, nnz_cap, st));
        NFB_TRY(cols.alloc_cached(&h->big, nnz_cap, st));
        NFB_TRY(vals.alloc_cached(&h->big, nnz_cap, st));
        marks.mark("alloc");
        // the triplets travel on the copy stream while the planes and the mask are zeroed on the main one
        // (the staging buffers are stream-ordered allocations of the main stream: the copy stream may only touch
        // them after the point of allocation)
        NFB_CHECK(up_guard.create() == 0, "cudaEventCreate failed");
        NFB_CUDA(cudaEventRecord(uploaded, st));
        NFB_CUDA(cudaStreamWaitEvent(h->copy_stream, uploaded, 0));
        NFB_CUDA(cudaMemcpyAsync(rows.p, rows_host, sizeof(int) * nnz, cudaMemcpyHostToDevice, h->copy_stream));
        NFB_CUDA(cudaMemcpyAsync(cols.p, cols_host, sizeof(int) * nnz, cudaMemcpyHostToDevice, h->copy_stream));
        NFB_CUDA(cudaMemcpyAsync(vals.p, vals_host, sizeof(double) * nnz, cudaMemcpyHostToDevice, h->copy_stream));
        NFB_CUDA(cudaEventRecord(uploaded, h->copy_stream));
        rows_dev = rows.p;
        cols_dev = cols.p;
        vals_dev = vals.p;
    }
    NFB_CUDA(cudaMemsetAsync(x->hi.p, 0, cells * sizeof(__half), st));
    NFB_CUDA(cudaMemsetAsync(x->lo.p, 0, cells * sizeof(__half), st));
    NFB_CUDA(cudaMemsetAsync(mask.p, 0, ((cells + 31) / 32) * sizeof(unsigned int), st));
    NFB_CUDA(cudaMemsetAsync(h->d_bits.p, 0, sizeof(unsigned int), st));
    NFB_CUDA(cudaMemsetAsync(h->d_scal.p + 6, 0, 2 * sizeof(double), st));
    NFB_CUDA(cudaMemsetAsync(h->d_int.p + 4, 0, 3 * sizeof(int), st));
    if (upload) {
        NFB_CUDA(cudaStreamWaitEvent(st, uploaded, 0));
        marks.mark("upload");
    }
    NFB_TRY(ingest_coo_to_planes(rows_dev, cols_dev, vals_dev, nnz, scale, nrows, ncols, h->d_bits.p, x->hi.p, x->lo.p,
                                 x->ldp, mask.p, x->inv_scale.p, h->d_scal.p + 6, h->d_int.p + 4, st));
    int flags[3] = {0, 0, 0};
    NFB_CUDA(cudaMemcpyAsync(flags, h->d_int.p + 4, sizeof(flags), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaMemcpyAsync(h->h_scal + 6, h->d_scal.p + 6, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaMemcpyAsync(&x->inv_scale_host, x->inv_scale.p, sizeof(float), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    up_guard.finish();
    marks.mark("scatter");
    note_launch(2);
    NFB_CHECK(flags[0] == 0, "row/column index exceeds matrix dimensions");
    *duplicates = flags[1] != 0;
    if (*duplicates) return 0;             // x (and its planes) is released; the caller takes the exact path
    x->has_negative = flags[2];
    x->sumsq = h->h_scal[6];
    x->sum = h->h_scal[7];
    rows.release(); cols.release(); vals.release(); mask.release();
    marks.mark("free");
    *out = x.release();
    return 0;
}

// fdt_matrix2.dot text -> COO triplets in device memory (dot_parse_gpu.cu)
struct DotCoo {
    DevBuf<int> rows, cols;
    DevBuf<double> vals;
    int64_t nnz = 0, nrows = 0, ncols = 0;
    int fallback = 0;      // 1: a number outside the device parser's exact fast path
};

static int dot_text_to_device_coo(nfb_handle* h, const char* text_host, int64_t len, DotCoo* coo) {
    cudaStream_t st = h->stream;
    HostMarks marks(st);
    NFB_CHECK(len > 0, "file is empty");
    const int64_t nseg = dot_gpu_segments(len), nblocks = dot_gpu_blocks(len);
    DevBuf<char> txt;
    DevBuf<unsigned char> seg_lines;
    DevBuf<unsigned int> block_lines;
    DevBuf<long long> block_first, total;
    NFB_TRY(txt.alloc_pooled(static_cast<size_t>(len), false, st));
    NFB_TRY(seg_lines.alloc_pooled(static_cast<size_t>(nseg), false, st));
    NFB_TRY(block_lines.alloc_pooled(static_cast<size_t>(nblocks), false, st));
    NFB_TRY(block_first.alloc_pooled(static_cast<size_t>(nblocks), false, st));
    NFB_TRY(total.alloc_pooled(1, true, st));
    NFB_CUDA(cudaMemcpyAsync(txt.p, text_host, static_cast<size_t>(len), cudaMemcpyHostToDevice, st));
    marks.mark("text upload");
    NFB_TRY(dot_gpu_count(txt.p, len, seg_lines.p, block_lines.p, block_first.p, total.p, st));
    long long n_lines = 0;
    NFB_CUDA(cudaMemcpyAsync(&n_lines, total.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    NFB_CHECK(n_lines > 0, "file is empty");
    coo->nnz = n_lines - 1;
    DevBuf<int> meta;
    DevBuf<double> shape;
    NFB_TRY(coo->rows.alloc_pooled(static_cast<size_t>(coo->nnz), false, st));
    NFB_TRY(coo->cols.alloc_pooled(static_cast<size_t>(coo->nnz), false, st));
    NFB_TRY(coo->vals.alloc_pooled(static_cast<size_t>(coo->nnz), false, st));
    NFB_TRY(shape.alloc_pooled(3, true, st));
    NFB_TRY(meta.alloc_pooled(8, true, st));             // flags[3] + range[4]
    NFB_TRY(dot_gpu_parse(txt.p, len, seg_lines.p, block_first.p, n_lines, coo->rows.p, coo->cols.p, coo->vals.p,
                          shape.p, meta.p, meta.p + 3, st));
    int meta_h[7] = {0, 0, 0, 0, 0, 0, 0};
    double shape_h[3] = {0.0, 0.0, 0.0};
    NFB_CUDA(cudaMemcpyAsync(meta_h, meta.p, sizeof(meta_h), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaMemcpyAsync(shape_h, shape.p, sizeof(shape_h), cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    marks.mark("parse");
    note_launch(4);
    NFB_CHECK(meta_h[0] == 0, "could not convert string to float: malformed line in fdt_matrix2.dot (expected 3 columns)");
    if (meta_h[1] != 0) {                                // a number outside the exact fast path: host parser
        coo->fallback = 1;
        return 0;
    }
    coo->nrows = static_cast<int64_t>(shape_h[0]);
    coo->ncols = static_cast<int64_t>(shape_h[1]);
    if (coo->nnz > 0) {                                  // the messages scipy's coo / csc constructor raises
        NFB_CHECK(meta_h[2] == 0, "row / column index is not a 32-bit integer");
        NFB_CHECK(meta_h[3] >= 0, "negative row index found");
        NFB_CHECK(meta_h[5] >= 0, "negative column index found");
        NFB_CHECK(meta_h[4] < coo->nrows, "row index exceeds matrix dimensions");
        NFB_CHECK(meta_h[6] < coo->ncols, "column index exceeds matrix dimensions");
    }
    return 0;
}

// ... and on to the resident matrix, entirely on the device (+ ingest.cu)
int nfb_ingest_dot_matrix(nfb_handle* h, const char* text_host, int64_t len, double scale, nfb_matrix** out,
                          int64_t* shape_out, int* fallback_out) {
    NFB_CHECK(h != nullptr && out != nullptr && shape_out != nullptr && fallback_out != nullptr && len >= 0 &&
                  (text_host != nullptr || len == 0), "bad nfb_ingest_dot_matrix arguments");
    NFB_CUDA(cudaSetDevice(h->device));
    *out = nullptr;
    *fallback_out = 0;
    DotCoo coo;
    NFB_TRY(dot_text_to_device_coo(h, text_host, len, &coo));
    if (coo.fallback) {
        *fallback_out = coo.fallback;
        return 0;
    }
    shape_out[0] = coo.nrows;
    shape_out[1] = coo.ncols;
    NFB_CHECK(coo.nrows > 0 && coo.ncols > 0, "bad ingest arguments");
    bool duplicates = false;
    NFB_TRY(ingest_single_to_planes(h, nullptr, nullptr, nullptr, coo.rows.p, coo.cols.p, coo.vals.p, coo.nnz, scale,
                                    coo.nrows, coo.ncols, out, &duplicates));
    if (duplicates) *fallback_out = 2;                   // duplicate coordinates: exact scatter-add path (host triplets)
    return 0;
}

// ... or back to the host as page-locked arrays (the group average adds subjects in float64 from host triplets)
int nfb_dot_parse_coo_gpu(nfb_handle* h, const char* text_host, int64_t len, int32_t** rows0_out, int32_t** cols0_out,
                          double** vals_out, int64_t* nnz_out, int64_t* shape_out, int* fallback_out) {
    NFB_CHECK(h != nullptr && rows0_out != nullptr && cols0_out != nullptr && vals_out != nullptr &&
                  nnz_out != nullptr && shape_out != nullptr && fallback_out != nullptr && len >= 0 &&
                  (text_host != nullptr || len == 0), "bad nfb_dot_parse_coo_gpu arguments");
    NFB_CUDA(cudaSetDevice(h->device));
    *rows0_out = *cols0_out = nullptr;
    *vals_out = nullptr;
    *fallback_out = 0;
    DotCoo coo;
    NFB_TRY(dot_text_to_device_coo(h, text_host, len, &coo));
    if (coo.fallback) {
        *fallback_out = coo.fallback;
        return 0;
    }
    void *r = nullptr, *c = nullptr, *v = nullptr;
    NFB_TRY(nfb_host_alloc(static_cast<int64_t>(sizeof(int32_t)) * coo.nnz, &r));
    NFB_TRY(nfb_host_alloc(static_cast<int64_t>(sizeof(int32_t)) * coo.nnz, &c));
    NFB_TRY(nfb_host_alloc(static_cast<int64_t>(sizeof(double)) * coo.nnz, &v));
    cudaStream_t st = h->stream;
    NFB_CUDA(cudaMemcpyAsync(r, coo.rows.p, sizeof(int32_t) * coo.nnz, cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaMemcpyAsync(c, coo.cols.p, sizeof(int32_t) * coo.nnz, cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaMemcpyAsync(v, coo.vals.p, sizeof(double) * coo.nnz, cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    *rows0_out = static_cast<int32_t*>(r);
    *cols0_out = static_cast<int32_t*>(c);
    *vals_out = static_cast<double*>(v);
    *nnz_out = coo.nnz;
    shape_out[0] = coo.nrows;
    shape_out[1] = coo.ncols;
    return 0;
}

int nfb_ingest_coo_matrix(nfb_handle* h, int n_subjects, const int32_t* const* rows_host,
                          const int32_t* const* cols_host, const double* const* vals_host, const int64_t* nnz,
                          const double* scale, int64_t nrows, int64_t ncols, int group_mode, nfb_matrix** out) {
    NFB_CHECK(h != nullptr && out != nullptr, "NULL argument");
    NFB_CHECK(n_subjects >= 1 && nrows > 0 && ncols > 0, "bad ingest arguments");
    NFB_CUDA(cudaSetDevice(h->device));
    static const bool direct_ok = !(getenv("NFB_INGEST_DIRECT") && atoi(getenv("NFB_INGEST_DIRECT")) == 0);
    if (!group_mode && n_subjects == 1 && direct_ok) {
        bool duplicates = false;
        *out = nullptr;
        NFB_TRY(ingest_single_to_planes(h, rows_host[0], cols_host[0], vals_host[0], nullptr, nullptr, nullptr, nnz[0],
                                        scale[0], nrows, ncols, out, &duplicates));
        if (!duplicates) return 0;
    }
    // group average, or duplicate coordinates to be summed in float64: the exact scatter-add path.
    // The float32 matrix exists only on the device and only until its fp16 planes are built.
    DevBuf<float> dense;
    NFB_TRY(dense.alloc_pooled(static_cast<size_t>(nrows) * ncols, false, h->stream));
    NFB_TRY(nfb_ingest_coo(h, n_subjects, rows_host, cols_host, vals_host, nnz, scale, nrows, ncols, group_mode,
                           nullptr, dense.p));
    HostMarks marks(h->stream);
    const int rc = nfb_matrix_from_device(h, dense.p, nrows, ncols, ncols, out);
    marks.mark("planes");
    dense.release();
    marks.mark("free dense");
    return rc;
}

int nfb_matrix_matmul(nfb_handle* h, const nfb_matrix* x, int trans, const float* b_dev, int64_t n_b,
                      float* out_dev, int64_t ld_out) {
    NFB_CHECK(h != nullptr && x != nullptr && b_dev != nullptr && out_dev != nullptr, "NULL argument");
    NFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t m_total = trans ? x->n : x->m;
    const int64_t k_total = trans ? x->m : x->n;
    NFB_CHECK(ld_out >= m_total, "ld_out too small");
    const int cg = h->cta_group;
    const int splits = gemm_pick_splits(m_total, k_total, cg, std::min<int64_t>(n_b, 512));
    for (int64_t t0 = 0; t0 < n_b; t0 += 512) {
        const int64_t nb = std::min<int64_t>(512, n_b - t0);
        Factor b;
        NFB_TRY(b.init(nb, k_total, true, st));
        NFB_CUDA(cudaMemcpy2DAsync(b.f.p, b.ld * sizeof(float), b_dev + t0 * k_total, k_total * sizeof(float),
                                   k_total * sizeof(float), nb, cudaMemcpyDeviceToDevice, st));
        NFB_TRY(refresh_planes(b, b.f.p, nullptr, true, st));
        DevBuf<float> part;
        NFB_TRY(part.alloc(static_cast<size_t>(splits) * nb * ld_out, false, st));
        NFB_TRY(gemm_split(x->operand(), trans != 0, b.operand(), m_total, k_total, static_cast<int>(b.kp),
                           static_cast<int>(nb), part.p, ld_out, splits, b.inv_scale.p, nullptr, x->inv_scale_host,
                           cg, st));
        NFB_TRY(reduce_partials(part.p, splits, nb * ld_out, nb, m_total, ld_out, 0.0f, out_dev + t0 * ld_out,
                                ld_out, st));
        NFB_CUDA(cudaStreamSynchronize(st));  // temporaries are freed at scope exit
    }
    return 0;
}

// ---- NMF-sso similarity ------------------------------------------------------------------------------
// rows_dev [r, n] float32 (contiguous) -> Pearson correlations of the rows on the host
static int row_correlation_dev(nfb_handle* h, const float* rows_dev, int64_t r, int64_t n, float* sim_out_host,
                               int32_t* dead_out_host) {
    cudaStream_t st = h->stream;
    DevBuf<float> z, sim;
    DevBuf<int> dead;
    NFB_TRY(z.alloc(static_cast<size_t>(r) * n, false, st));
    NFB_TRY(sim.alloc(static_cast<size_t>(r) * r, false, st));
    NFB_TRY(dead.alloc(r, true, st));
    NFB_TRY(rows_standardize(rows_dev, r, n, n, z.p, n, dead.p, st));
    // corr = Z Z': Z as the resident "matrix" (K-major read), 512 rows of Z at a time as the factor operand
    nfb_matrix* zmat = nullptr;
    NFB_TRY(nfb_matrix_from_device(h, z.p, r, n, n, &zmat));
    std::unique_ptr<nfb_matrix> guard(zmat);
    NFB_TRY(nfb_matrix_matmul(h, zmat, 0, z.p, r, sim.p, r));
    NFB_CUDA(cudaMemcpyAsync(sim_out_host, sim.p, sizeof(float) * r * r, cudaMemcpyDeviceToHost, st));
    if (dead_out_host)
        NFB_CUDA(cudaMemcpyAsync(dead_out_host, dead.p, sizeof(int) * r, cudaMemcpyDeviceToHost, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int nfb_row_correlation_host(nfb_handle* h, const float* rows_host, int64_t r, int64_t n, float* sim_out_host,
                             int32_t* dead_out_host) {
    NFB_CHECK(h != nullptr && rows_host != nullptr && sim_out_host != nullptr, "NULL argument");
    NFB_CHECK(r >= 1 && n >= 2, "need at least one row and two columns");
    NFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    DevBuf<float> raw;
    NFB_TRY(raw.alloc(static_cast<size_t>(r) * n, false, st));
    NFB_CUDA(cudaMemcpyAsync(raw.p, rows_host, sizeof(float) * r * n, cudaMemcpyHostToDevice, st));
    return row_correlation_dev(h, raw.p, r, n, sim_out_host, dead_out_host);
}

// ---- NMF-sso on the device --------------------------------------------------------------------------------
// The reference's NMFsso (nmfsso.py:62-238, 322-360) runs N random-initialised solves, z-thresholds every run's white
// components, stacks all runs, correlates the stack, clusters it on the host and starts the final solve from the
// centrotype rows.  Here the stacks live in HBM: a run's factors are copied from the solver state into the stacks
// (device to device) and thresholded there, the similarity is computed from the resident stack, the centrotypes
// are gathered on the device straight into the final solver state.  Only the [R, R] similarity matrix and the
// final factors ever travel to the host.
struct nfb_sso {
    nfb_handle* h = nullptr;
    int64_t m = 0, n = 0;
    int k = 0, runs = 0;
    DevBuf<float> white;    // [runs * k][n]   thresholded white components of every run
    DevBuf<float> grey;     // [runs * k][m]   grey components of every run, component-major
    DevBuf<double> stats;
};

int nfb_sso_create(nfb_handle* h, int64_t m, int64_t n, int k, int runs, nfb_sso** out) {
    NFB_CHECK(h != nullptr && out != nullptr, "NULL argument");
    NFB_CHECK(m > 0 && n > 0 && k > 0 && runs > 0, "bad NMF-sso shape");
    NFB_CUDA(cudaSetDevice(h->device));
    std::unique_ptr<nfb_sso> s(new nfb_sso());
    s->h = h;
    s->m = m;
    s->n = n;
    s->k = k;
    s->runs = runs;
    NFB_TRY(s->white.alloc(static_cast<size_t>(runs) * k * n, true, h->stream));
    NFB_TRY(s->grey.alloc(static_cast<size_t>(runs) * k * m, true, h->stream));
    NFB_TRY(s->stats.alloc(static_cast<size_t>(4) * k, false, h->stream));
    *out = s.release();
    return 0;
}

int nfb_sso_free(nfb_sso* s) {
    if (s == nullptr) return 0;
    cudaSetDevice(s->h->device);
    cudaStreamSynchronize(s->h->stream);
    delete s;
    return 0;
}

// |avg N(0,1)| into both factors of a solver state (sklearn's init='random' distribution, _nmf.py:296-307)
int nfb_nmf_init_random(nfb_nmf* s, double avg, uint64_t seed) {
    NFB_CHECK(s != nullptr, "NULL argument");
    nfb_handle* h = s->h;
    NFB_CHECK(h->world <= 1, "nfb_nmf_init_random is for replicas on one GPU (row-sharded runs draw on the host)");
    NFB_CUDA(cudaSetDevice(h->device));
    NFB_TRY(fill_abs_normal(s->hh.f.p, s->k, s->x->n, s->hh.ld, static_cast<float>(avg), seed, h->stream));
    NFB_TRY(fill_abs_normal(s->wt.f.p, s->k, s->x->m, s->wt.ld, static_cast<float>(avg), seed ^ 0x5bf03635f0a5b0c1ULL,
                            h->stream));
    return nmf_prepare_factors(s);
}

// factors of a finished run -> rows [run k, (run + 1) k) of the stacks; the white copy is z-thresholded per
// component in place (thresholding(components, zscore), base/matrix_handling.py:218-238; zscore <= 0: left as is)
int nfb_sso_add_run(nfb_sso* sso, nfb_nmf* s, int run, double zscore) {
    NFB_CHECK(sso != nullptr && s != nullptr, "NULL argument");
    NFB_CHECK(run >= 0 && run < sso->runs, "run index out of range");
    NFB_CHECK(s->k == sso->k && s->x->m == sso->m && s->x->n == sso->n, "solver state does not match the NMF-sso shape");
    nfb_handle* h = sso->h;
    NFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    float* white = sso->white.p + static_cast<int64_t>(run) * sso->k * sso->n;
    float* grey = sso->grey.p + static_cast<int64_t>(run) * sso->k * sso->m;
    NFB_CUDA(cudaMemcpy2DAsync(white, sso->n * sizeof(float), s->hh.f.p, s->hh.ld * sizeof(float),
                               sso->n * sizeof(float), sso->k, cudaMemcpyDeviceToDevice, st));
    NFB_CUDA(cudaMemcpy2DAsync(grey, sso->m * sizeof(float), s->wt.f.p, s->wt.ld * sizeof(float),
                               sso->m * sizeof(float), sso->k, cudaMemcpyDeviceToDevice, st));
    if (zscore > 0.0) {
        NFB_TRY(comp_threshold<float>(white, sso->k, sso->n, sso->n, nullptr, 1, zscore, sso->stats.p, st));
        note_launch(2);
    }
    return 0;
}

int nfb_sso_similarity_host(nfb_sso* sso, float* sim_out_host, int32_t* dead_out_host) {
    NFB_CHECK(sso != nullptr && sim_out_host != nullptr, "NULL argument");
    NFB_CHECK(sso->n >= 2, "need at least two columns");
    NFB_CUDA(cudaSetDevice(sso->h->device));
    return row_correlation_dev(sso->h, sso->white.p, static_cast<int64_t>(sso->runs) * sso->k, sso->n, sim_out_host,
                               dead_out_host);
}

// start of the final solve (nmfsso.py:346-357): W = stacked grey[:, centroids], H = stacked white[centroids, :]
int nfb_sso_seed_final(nfb_sso* sso, nfb_nmf* s, const int32_t* centroids_host, int n_centroids) {
    NFB_CHECK(sso != nullptr && s != nullptr && centroids_host != nullptr, "NULL argument");
    NFB_CHECK(n_centroids == s->k, "the final solver state must have one component per centrotype");
    NFB_CHECK(s->x->m == sso->m && s->x->n == sso->n, "solver state does not match the NMF-sso shape");
    nfb_handle* h = sso->h;
    NFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t rows = static_cast<int64_t>(sso->runs) * sso->k;
    for (int c = 0; c < n_centroids; ++c) {
        const int64_t r = centroids_host[c];
        NFB_CHECK(r >= 0 && r < rows, "centrotype index out of range");
        NFB_CUDA(cudaMemcpyAsync(s->hh.f.p + static_cast<int64_t>(c) * s->hh.ld, sso->white.p + r * sso->n,
                                 sizeof(float) * sso->n, cudaMemcpyDeviceToDevice, st));
        NFB_CUDA(cudaMemcpyAsync(s->wt.f.p + static_cast<int64_t>(c) * s->wt.ld, sso->grey.p + r * sso->m,
                                 sizeof(float) * sso->m, cudaMemcpyDeviceToDevice, st));
    }
    return nmf_prepare_factors(s);
}

// sklearn's checks on a custom start (_nmf.py:1186-1203) for factors that never visit the host:
// flags[0] = 1 when W is all zeros, flags[1] = 1 when H is all zeros (values are non-negative by construction)
int nfb_nmf_factors_all_zero(nfb_nmf* s, int32_t* flags_host) {
    NFB_CHECK(s != nullptr && flags_host != nullptr, "NULL argument");
    nfb_handle* h = s->h;
    NFB_CUDA(cudaSetDevice(h->device));
    std::vector<unsigned int> bits(2 * s->kp);
    NFB_CUDA(cudaMemcpyAsync(bits.data(), s->wt.rowmax.p, sizeof(unsigned int) * s->kp, cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaMemcpyAsync(bits.data() + s->kp, s->hh.rowmax.p, sizeof(unsigned int) * s->kp, cudaMemcpyDeviceToHost, h->stream));
    NFB_CUDA(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < 2; ++f) {
        unsigned int any = 0;
        for (int t = 0; t < s->k; ++t) any |= bits[f * s->kp + t];
        flags_host[f] = any == 0u ? 1 : 0;
    }
    return 0;
}

// ---- test hook --------------------------------------------------------------------------------------
int nfb_test_gemm(nfb_handle* h, const float* a_dev, int64_t a_rows, int64_t a_cols, int a_is_transposed,
                  const float* b_dev, int64_t n_b, float* out_dev, int64_t ld_out, int cta_group, int splits) {
    NFB_CHECK(h != nullptr, "handle is NULL");
    NFB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t m_total = a_is_transposed ? a_cols : a_rows;
    const int64_t k_total = a_is_transposed ? a_rows : a_cols;
    nfb_matrix* a = nullptr;
    NFB_TRY(nfb_matrix_from_device(h, a_dev, a_rows, a_cols, a_cols, &a));
    std::unique_ptr<nfb_matrix> a_guard(a);
    Factor b;
    NFB_TRY(b.init(n_b, k_total, false, st));
    // b_dev is [n_b, k_total] with ld = k_total; Factor planes use ld rounded up to 8
    DevBuf<float> bpad;
    NFB_TRY(bpad.alloc(static_cast<size_t>(b.kp) * b.ld, true, st));
    NFB_CUDA(cudaMemcpy2DAsync(bpad.p, b.ld * sizeof(float), b_dev, k_total * sizeof(float), k_total * sizeof(float),
                               n_b, cudaMemcpyDeviceToDevice, st));
    NFB_TRY(refresh_planes(b, bpad.p, nullptr, true, st));
    if (cta_group == 0) cta_group = h->cta_group;
    // splits == -1: unsplit with the stream-K tail (the launch falls back to whole tiles where that does not pay)
    const bool streamk = splits == -1;
    if (streamk) splits = 1;
    if (splits <= 0) splits = gemm_pick_splits(m_total, k_total, cta_group, n_b);
    DevBuf<float> part, sk_ws;
    DevBuf<unsigned int> sk_cnt;
    GemmStreamK sk{};
    if (streamk) {
        const size_t ws_floats = gemm_streamk_ws_floats(static_cast<int>(std::min<int64_t>(n_b, 256)));
        NFB_TRY(sk_ws.alloc(ws_floats, false, st));
        NFB_TRY(sk_cnt.alloc(2 * static_cast<size_t>(gemm_streamk_max_workers()), true, st));
        sk = GemmStreamK{sk_ws.p, sk_cnt.p, ws_floats, gemm_streamk_max_workers(), true};
    }
    NFB_TRY(part.alloc(static_cast<size_t>(splits) * n_b * ld_out, false, st));
    // twice with the same workspace: the second launch sees the counters the first one re-armed
    for (int rep = 0; rep < (streamk ? 2 : 1); ++rep)
    NFB_TRY(gemm_split(a->operand(), a_is_transposed != 0, b.operand(), m_total, k_total, static_cast<int>(b.kp),
                       static_cast<int>(n_b), part.p, ld_out, splits, b.inv_scale.p, nullptr, a->inv_scale_host,
                       cta_group, st, nullptr, streamk ? &sk : nullptr));
    NFB_TRY(reduce_partials(part.p, splits, n_b * ld_out, n_b, m_total, ld_out, 0.0f, out_dev, ld_out, st));
    NFB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
