// Coordinate-descent (HALS) sweep of one NMF factor -- sklearn/decomposition/_cdnmf_fast.pyx:8-38.
//
// sklearn walks the components t = 0..k-1 and, for every sample i, recomputes
//     grad = -XHt[i,t] + sum_r HHt[t,r] W[i,r]                       (k FMAs per coordinate, k^2 per sample)
//     pg = W[i,t] == 0 ? min(0, grad) : grad ;  violation += |pg|
//     W[i,t] = max(W[i,t] - grad / HHt[t,t], 0)
// Samples are independent; components are Gauss-Seidel.  The gradient g = G F[:,j] - num[:,j] of every sample
// starts from the tensor-core product G.F (the same "denominator" GEMM the multiplicative update uses) and is
// kept current incrementally, so a sweep costs k^2 / 2 FMAs per sample instead of k^2, in the exact
// Gauss-Seidel order of the reference.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

// ---- blocked sweep: thread-per-sample steps + register-tiled rank-16 updates -----------------------------
// A CTA owns TJ samples (columns j of F).  Their gradient vectors live in shared memory as gs[u][c]
// (component-major, conflict-free for a thread-per-sample walk).  Components are processed in blocks of 16:
//   step phase  : thread c runs the 16 coordinate steps of sample c in registers against the 16 x 16 diagonal
//                 block of G (exact Gauss-Seidel order) and leaves its 16 deltas in the - now dead - gradient
//                 rows of the block;
//   apply phase : gs[u][c] += sum_i delta_i[c] G[t0+i][u] for the components still to come, as a register-tiled
//                 FMA GEMM: every thread owns 4 samples x 4 components, keeps its 64 deltas in registers and
//                 reads G as one broadcast 128-bit load per 16 FMAs.
// The FMA order per gradient entry (i = 0..15 within a block, blocks in order) is that of the rank-1
// formulation above, so results are bit-identical to it; the cost is ~k^2/2 FMAs per sample at FMA-pipe rate.
// The next block's G rows, hessians and factor values are prefetched into registers across the apply phase.
constexpr int kBlk = 16;

template <int TJ, int NT>
__global__ void __launch_bounds__(NT, 1)
cd_sweep_tile_kernel(float* __restrict__ f, int k, int64_t cols, int64_t ld, const float* __restrict__ num,
                     int num_splits, int64_t num_stride, const float* __restrict__ den, int64_t ld_part,
                     const float* __restrict__ gram, int ldg, float l1, float l2, double* __restrict__ violation,
                     unsigned int* __restrict__ rowmax_bits, bool vec, const PeerMirror mirror) {
    constexpr int JG = TJ / 4;        // groups of 4 samples
    constexpr int UG = NT / JG;       // groups of 4 components updated concurrently in the apply phase
    static_assert(TJ % 32 == 0 && NT % TJ == 0 && NT % JG == 0, "tile shape");
    extern __shared__ __align__(16) float sm[];
    __shared__ double scratch[32];
    __shared__ float invh[kBlk];
    __shared__ unsigned int max_s[kBlk];
    const int kp4 = (k + 3) & ~3;
    float* gs = sm;                                        // [kp4][TJ] gradient (rows >= k stay zero)
    float* gb = sm + static_cast<size_t>(kp4) * TJ;        // [kBlk][kp4]: gb[i][u] = G[t0 + i][u], 0 for u >= k
    const int tid = threadIdx.x;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * TJ + tid;
    const bool stepper = tid < TJ;
    const bool active = stepper && j < cols;

    // column `tid` (and tid + NT, ...) of the 16 G rows of block tb: independent clamped loads whose values are
    // only looked at when they are committed to shared memory, so the round trip to L2 overlaps whatever runs
    // in between (the gradient seed for block 0, the apply phase afterwards)
    const int u_own = min(tid, k - 1);
    auto fetch_rows = [&](int tb, float (&pre)[kBlk]) {
        if (tb + kBlk <= k) {                              // whole block inside the matrix: one base, 16 strides
            const float* gp = gram + static_cast<int64_t>(tb) * ldg + u_own;
#pragma unroll
            for (int i = 0; i < kBlk; ++i) pre[i] = __ldg(gp + i * ldg);
        } else {
#pragma unroll
            for (int i = 0; i < kBlk; ++i) pre[i] = __ldg(gram + static_cast<int64_t>(min(tb + i, k - 1)) * ldg + u_own);
        }
    };
    auto fetch_factor = [&](int tb, int64_t jc_own, float (&fl)[kBlk]) {   // F[tb + i][own sample], clamped rows
        if (tb + kBlk <= k) {
            const float* fp = f + static_cast<int64_t>(tb) * ld + jc_own;
#pragma unroll
            for (int i = 0; i < kBlk; ++i) fl[i] = fp[i * ld];
        } else {
#pragma unroll
            for (int i = 0; i < kBlk; ++i) fl[i] = f[static_cast<int64_t>(min(tb + i, k - 1)) * ld + jc_own];
        }
    };
    auto fetch_hess = [&](int tb) -> float {               // raw G[t][t] of component tb + tid (clamped)
        const int t = min(tb + (tid & (kBlk - 1)), k - 1);
        return __ldg(gram + static_cast<int64_t>(t) * ldg + t);
    };
    auto commit_rows = [&](int tb, const float (&pre)[kBlk], float hess_raw) {
        if (tid < kp4) {
#pragma unroll
            for (int i = 0; i < kBlk; ++i) gb[i * kp4 + tid] = (tb + i < k && tid < k) ? pre[i] : 0.0f;
        }
        if (kp4 > NT) {
            // columns the register prefetch does not cover (k > NT): 16 independent clamped loads per pass
            for (int base = NT; base < kp4; base += NT) {
                const int u = base + tid, uc = min(u, k - 1);
                float ov[kBlk];
#pragma unroll
                for (int i = 0; i < kBlk; ++i) ov[i] = __ldg(gram + static_cast<int64_t>(min(tb + i, k - 1)) * ldg + uc);
                if (u < kp4) {
#pragma unroll
                    for (int i = 0; i < kBlk; ++i) gb[i * kp4 + u] = (tb + i < k && u < k) ? ov[i] : 0.0f;
                }
            }
        }
        if (tid < kBlk) {
            const float hess = tb + tid < k ? hess_raw + l2 : 0.0f;
            invh[tid] = hess != 0.0f ? 1.0f / hess : 0.0f;
        }
    };

    // ---- block 0 operands in flight, then the gradient seed g = G F[:, j] + l2 F[:, j] - (num[:, j] - l1) ----
    float pre0[kBlk];
    fetch_rows(0, pre0);
    const float hess0 = fetch_hess(0);
    if (tid < kBlk) max_s[tid] = 0u;
    {
        // every thread owns 4 consecutive samples of one component row per pass; all (splits + 2) 128-bit loads
        // of a pass are issued before the first use, and the loads of pass p + 1 are in flight while pass p is
        // summed (two register buffers), so the phase runs at memory bandwidth, not latency
        constexpr int kSeedBatch = 8;
        const int c4 = tid % JG;
        const int64_t jc = static_cast<int64_t>(blockIdx.x) * TJ + 4 * c4;
        const bool full = vec && jc + 3 < cols;
        int u_first = tid / JG;
        if (full && num_splits <= kSeedBatch) {
            struct Row { float4 fv, dv, b[kSeedBatch]; };
            auto issue = [&](int u, Row& r) {
                const int64_t uc = min(u, k - 1);          // padding rows (u >= k) load row k - 1 and store zeros
                r.fv = *reinterpret_cast<const float4*>(f + uc * ld + jc);
                r.dv = __ldcs(reinterpret_cast<const float4*>(den + uc * ld_part + jc));
                const float* np = num + uc * ld_part + jc;
#pragma unroll
                for (int q = 0; q < kSeedBatch; ++q)
                    if (q < num_splits) r.b[q] = __ldcs(reinterpret_cast<const float4*>(np + q * num_stride));
            };
            auto finish = [&](int u, const Row& r) {
                float4 nv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int q = 0; q < kSeedBatch; ++q)
                    if (q < num_splits) { nv.x += r.b[q].x; nv.y += r.b[q].y; nv.z += r.b[q].z; nv.w += r.b[q].w; }
                float4 gv;
                gv.x = fmaf(l2, r.fv.x, r.dv.x) - (nv.x - l1);
                gv.y = fmaf(l2, r.fv.y, r.dv.y) - (nv.y - l1);
                gv.z = fmaf(l2, r.fv.z, r.dv.z) - (nv.z - l1);
                gv.w = fmaf(l2, r.fv.w, r.dv.w) - (nv.w - l1);
                if (u >= k) gv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                *reinterpret_cast<float4*>(gs + u * TJ + 4 * c4) = gv;
            };
            Row ra, rb;
            int u = u_first;
            if (u < kp4) issue(u, ra);
            while (u < kp4) {
                const int u1 = u + UG;
                if (u1 < kp4) issue(u1, rb);
                finish(u, ra);
                if (u1 >= kp4) break;
                u = u1 + UG;
                if (u < kp4) issue(u, ra);
                finish(u1, rb);
            }
            u_first = kp4;                                 // nothing left for the generic loop below
        }
        for (int u = u_first; u < kp4; u += UG) {
            float4 gv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (u < k && full) {
                const float4 fv = *reinterpret_cast<const float4*>(f + static_cast<int64_t>(u) * ld + jc);
                const float4 dv = __ldcs(reinterpret_cast<const float4*>(den + static_cast<int64_t>(u) * ld_part + jc));
                const float* np = num + static_cast<int64_t>(u) * ld_part + jc;
                float4 nv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                for (int s0 = 0; s0 < num_splits; s0 += kSeedBatch) {
                    float4 b[kSeedBatch];
#pragma unroll
                    for (int q = 0; q < kSeedBatch; ++q)
                        b[q] = (s0 + q < num_splits) ? __ldcs(reinterpret_cast<const float4*>(np + (s0 + q) * num_stride))
                                                     : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                    for (int q = 0; q < kSeedBatch; ++q) {
                        nv.x += b[q].x; nv.y += b[q].y; nv.z += b[q].z; nv.w += b[q].w;
                    }
                }
                gv.x = fmaf(l2, fv.x, dv.x) - (nv.x - l1);
                gv.y = fmaf(l2, fv.y, dv.y) - (nv.y - l1);
                gv.z = fmaf(l2, fv.z, dv.z) - (nv.z - l1);
                gv.w = fmaf(l2, fv.w, dv.w) - (nv.w - l1);
            } else if (u < k) {
                float out[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    out[q] = 0.0f;
                    if (jc + q < cols) {
                        const float fv = f[static_cast<int64_t>(u) * ld + jc + q];
                        float nv = 0.0f;
                        for (int sp = 0; sp < num_splits; ++sp)
                            nv += __ldcs(num + sp * num_stride + static_cast<int64_t>(u) * ld_part + jc + q);
                        out[q] = fmaf(l2, fv, __ldcs(den + static_cast<int64_t>(u) * ld_part + jc + q)) - (nv - l1);
                    }
                }
                gv = make_float4(out[0], out[1], out[2], out[3]);
            }
            *reinterpret_cast<float4*>(gs + u * TJ + 4 * c4) = gv;
        }
    }
    // factor values of the block in flight: unconditional loads (clamped indices) whose predicates are applied
    // where the values are used, so that the 16 of them are independent straight-line loads
    const int64_t jc_own = active ? j : 0;
    float fl[kBlk];
    fetch_factor(0, jc_own, fl);
    commit_rows(0, pre0, hess0);
    __syncthreads();

    float viol = 0.0f;
    const int jg = tid % JG, ug = tid / JG;
    // one block of coordinate steps for sample `tid`; kFull: all 16 components exist (no bounds predicates)
    auto step_block = [&](auto full_tag, int t0, int nb) {
        constexpr bool kFull = decltype(full_tag)::value;
        // The 16 steps form one dependent chain per thread (fn -> delta -> next gradient), ~5 FP32 operations per
        // step.  Nothing else may sit on that chain: the deltas and new values stay in registers until the block
        // is done (a shared-memory store inside the loop would pin every later G load behind it, and the warp
        // reduction for the per-component maximum brings a branch), so the broadcast loads of G's diagonal block
        // can be issued ahead of the arithmetic.
        float gl[kBlk], dl[kBlk];
        unsigned int fb[kBlk];
#pragma unroll
        for (int i = 0; i < kBlk; ++i) gl[i] = (kFull || i < nb) ? gs[(t0 + i) * TJ + tid] : 0.0f;
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            dl[i] = 0.0f;
            fb[i] = 0u;
            if (kFull || i < nb) {
                const float ft = active ? fl[i] : 0.0f;
                const float gt = gl[i];
                viol += fabsf((ft == 0.0f) ? fminf(0.0f, gt) : gt);
                const float inv = invh[i];
                const float fn = inv != 0.0f ? fmaxf(fmaf(-gt, inv, ft), 0.0f) : ft;   // a dead component keeps its value
                const float delta = fn - ft;
                dl[i] = delta;
                fb[i] = active ? __float_as_uint(fn) : 0u;   // the value to store; >= 0: unsigned order == float order
                // row i of the diagonal block of G, columns i+1..15, as 128-bit broadcast loads
                const float4* grow4 = reinterpret_cast<const float4*>(gb + i * kp4 + t0);
                float grow[kBlk];
#pragma unroll
                for (int c = (i + 1) / 4; c < kBlk / 4; ++c) {
                    const float4 q = grow4[c];
                    grow[4 * c] = q.x; grow[4 * c + 1] = q.y; grow[4 * c + 2] = q.z; grow[4 * c + 3] = q.w;
                }
#pragma unroll
                for (int i2 = i + 1; i2 < kBlk; ++i2)
                    if (kFull || i2 < nb) gl[i2] = fmaf(delta, grow[i2], gl[i2]);
            }
        }
        // the gradient rows of the block are dead: they now carry the deltas for the apply phase
#pragma unroll
        for (int i = 0; i < kBlk; ++i)
            if (kFull || i < nb) gs[(t0 + i) * TJ + tid] = dl[i];
        // The block's new values leave the chain in one run of stores behind it (a store per step inside the chain
        // cost a 64-bit address, a predicate and -- where no lane of the warp moved -- a branch per coordinate:
        // ~10 of the ~44 instructions of a step).  Unchanged values are rewritten: the whole factor, 4 B per entry
        // and iteration, against the (S + 3) x 4 B the seed phase reads.
        if (active) {
            float* fp = f + static_cast<int64_t>(t0) * ld + jc_own;
#pragma unroll
            for (int i = 0; i < kBlk; ++i) {
                if (kFull || i < nb) *fp = __uint_as_float(fb[i]);
                fp += ld;
            }
        }
        // Row-sharded runs with rows that cannot take the 128-bit path below: every value that moved goes straight
        // into the peers' copies of H from here
        if (mirror.n > 0 && !vec && active) {
            const int64_t off = static_cast<int64_t>(t0) * ld + jc_own;
#pragma unroll
            for (int i = 0; i < kBlk; ++i) {
                if ((kFull || i < nb) && dl[i] != 0.0f) {
                    const float v = __uint_as_float(fb[i]);
#pragma unroll
                    for (int p = 0; p < 7; ++p)
                        if (p < mirror.n) mirror.f[p][off + static_cast<int64_t>(i) * ld] = v;
                }
            }
        }
        // per-component maxima for the plane split: 16 independent warp reductions, one shared atomic per warp
        unsigned int vmax[kBlk];
#pragma unroll
        for (int i = 0; i < kBlk; ++i) vmax[i] = __reduce_max_sync(0xffffffffu, fb[i]);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int i = 0; i < kBlk; ++i)
                if (vmax[i] != 0u) atomicMax(&max_s[i], vmax[i]);
        }
    };
    for (int t0 = 0; t0 < k; t0 += kBlk) {
        const int nb = min(kBlk, k - t0);
        const bool more = t0 + kBlk < k;                   // components after this block exist (then nb == 16)
        // ---- step phase: the block's coordinate steps for sample `tid`, in registers ---------------------
        if (stepper) {
            if (nb == kBlk) step_block(std::true_type{}, t0, nb);
            else step_block(std::false_type{}, t0, nb);
        }
        // ---- prefetch the next block's operands across the apply phase --------------------------------------
        float pre[kBlk];                                   // column `tid` of the next block's G rows
        float hess_next = 0.0f;
        if (more) {
            fetch_rows(t0 + kBlk, pre);
            hess_next = fetch_hess(t0 + kBlk);
            fetch_factor(t0 + kBlk, jc_own, fl);
        }
        __syncthreads();
        if (tid < kBlk) {
            if (t0 + tid < k && max_s[tid] != 0u) atomicMax(&rowmax_bits[t0 + tid], max_s[tid]);
            max_s[tid] = 0u;
        }
        // Row-sharded runs: the block's new values go straight into the peers' copies of H (the all-gather of the
        // slices, fused).  All threads of the CTA share the job, four samples per 128-bit store: the values come
        // back from this CTA's own stores to F (visible after the barrier above), quads whose four deltas are zero
        // -- most zeros stay zero -- are skipped, so unchanged entries cost no NVLink traffic, and the stores of
        // this block travel while the apply phase and the following blocks compute.  (Done by the stepper threads
        // inside their dependent chain this cost 0.09 ms per sweep at 8 GPUs.)
        if (mirror.n > 0 && vec) {
            const int64_t j0 = static_cast<int64_t>(blockIdx.x) * TJ;
            for (int q = tid; q < nb * JG; q += NT) {
                const int i = q / JG, c4 = q - i * JG;
                const int64_t jq = j0 + 4 * c4;
                if (jq >= cols) continue;
                const float4 dq = *reinterpret_cast<const float4*>(gs + (t0 + i) * TJ + 4 * c4);
                if (dq.x == 0.0f && dq.y == 0.0f && dq.z == 0.0f && dq.w == 0.0f) continue;
                const int64_t off = static_cast<int64_t>(t0 + i) * ld + jq;
                if (jq + 3 < cols) {
                    const float4 v = *reinterpret_cast<const float4*>(f + off);
#pragma unroll
                    for (int p = 0; p < 7; ++p)
                        if (p < mirror.n) *reinterpret_cast<float4*>(mirror.f[p] + off) = v;
                } else {
                    for (int e = 0; e < 4 && jq + e < cols; ++e) {
                        const float v = f[off + e];
#pragma unroll
                        for (int p = 0; p < 7; ++p)
                            if (p < mirror.n) mirror.f[p][off + e] = v;
                    }
                }
            }
        }
        if (!more) break;
        // ---- apply phase: gs[u][4 samples] += sum_i delta_i G[t0+i][u], 4 x 4 register tiles ---------------
        {
            float4 d[kBlk];
#pragma unroll
            for (int i = 0; i < kBlk; ++i) d[i] = *reinterpret_cast<const float4*>(gs + (t0 + i) * TJ + 4 * jg);
            for (int ub = t0 + kBlk + 4 * ug; ub < kp4; ub += 4 * UG) {
                float4* grow = reinterpret_cast<float4*>(gs + ub * TJ + 4 * jg);
                float4 a0 = grow[0], a1 = grow[TJ / 4], a2 = grow[2 * (TJ / 4)], a3 = grow[3 * (TJ / 4)];
#pragma unroll
                for (int i = 0; i < kBlk; ++i) {
                    const float4 g = *reinterpret_cast<const float4*>(gb + i * kp4 + ub);
                    a0.x = fmaf(d[i].x, g.x, a0.x); a0.y = fmaf(d[i].y, g.x, a0.y);
                    a0.z = fmaf(d[i].z, g.x, a0.z); a0.w = fmaf(d[i].w, g.x, a0.w);
                    a1.x = fmaf(d[i].x, g.y, a1.x); a1.y = fmaf(d[i].y, g.y, a1.y);
                    a1.z = fmaf(d[i].z, g.y, a1.z); a1.w = fmaf(d[i].w, g.y, a1.w);
                    a2.x = fmaf(d[i].x, g.z, a2.x); a2.y = fmaf(d[i].y, g.z, a2.y);
                    a2.z = fmaf(d[i].z, g.z, a2.z); a2.w = fmaf(d[i].w, g.z, a2.w);
                    a3.x = fmaf(d[i].x, g.w, a3.x); a3.y = fmaf(d[i].y, g.w, a3.y);
                    a3.z = fmaf(d[i].z, g.w, a3.z); a3.w = fmaf(d[i].w, g.w, a3.w);
                }
                grow[0] = a0; grow[TJ / 4] = a1; grow[2 * (TJ / 4)] = a2; grow[3 * (TJ / 4)] = a3;
            }
        }
        __syncthreads();
        // ---- next block's G rows / hessians into shared memory --------------------------------------------
        commit_rows(t0 + kBlk, pre, hess_next);
        __syncthreads();
    }
    const double total = block_sum(active ? static_cast<double>(viol) : 0.0, scratch);
    if (tid == 0 && total != 0.0) atomicAdd(violation, total);
}

}  // namespace

int cd_sweep(float* f, int64_t k, int64_t cols, int64_t ld, const float* num, int num_splits, int64_t num_stride,
             const float* den, int64_t ld_part, const float* gram, float l1, float l2, double* violation,
             unsigned int* rowmax_bits, cudaStream_t stream, const PeerMirror* mirror) {
    if (k * cols == 0) return 0;
    NFB_CHECK(k <= 512, "n_components > 512 is not supported by the coordinate-descent kernel");
    const PeerMirror mir = mirror ? *mirror : PeerMirror{};
    const int ldg = static_cast<int>(round_up64(k, 16));
    // blocked sweep: pick the sample tile that minimises waves x tile (one CTA per SM)
    {
        const int64_t kp4 = round_up64(k, 4);
        // 128-bit seed loads need 16-byte aligned rows in all three operands
        const bool vec = ((ld | ld_part | num_stride) % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(num) |
                           reinterpret_cast<uintptr_t>(den)) % 16 == 0);
        auto smem_tile = [&](int tj) { return static_cast<size_t>(kp4) * (tj + kBlk) * sizeof(float); };
        auto run_tile = [&](auto tj_tag, auto nt_tag) -> int {
            constexpr int TJ = decltype(tj_tag)::value;
            constexpr int NT = decltype(nt_tag)::value;
            auto kernel = cd_sweep_tile_kernel<TJ, NT>;
            const size_t smem = smem_tile(TJ);
            NFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kernel<<<static_cast<unsigned>(ceil_div64(cols, TJ)), NT, smem, stream>>>(
                f, static_cast<int>(k), cols, ld, num, num_splits, num_stride, den, ld_part, gram, ldg, l1, l2,
                violation, rowmax_bits, vec, mir);
            NFB_CUDA(cudaGetLastError());
            return 0;
        };
        static const int tiles[] = {256, 224, 192, 160, 128, 96, 64, 32};
        const size_t smem_cap = 227 * 1024 - 1024;
        int best = 0;
        int64_t best_cost = 0;
        const int forced = getenv("NFB_CD_TILE") ? atoi(getenv("NFB_CD_TILE")) : 0;   // tests: pin the tile size
        for (int tj : tiles) {
            if (smem_tile(tj) > smem_cap) continue;
            if (forced != 0 && tj != forced) continue;
            const int64_t waves = ceil_div64(ceil_div64(cols, tj), sm_count());
            const int64_t cost = waves * (tj + 48);        // + the per-CTA latency-bound part
            if (best == 0 || cost < best_cost) { best = tj; best_cost = cost; }
        }
        switch (best) {
            case 256: return run_tile(std::integral_constant<int, 256>{}, std::integral_constant<int, 256>{});
            case 224: return run_tile(std::integral_constant<int, 224>{}, std::integral_constant<int, 224>{});
            case 192: return run_tile(std::integral_constant<int, 192>{}, std::integral_constant<int, 192>{});
            case 160: return run_tile(std::integral_constant<int, 160>{}, std::integral_constant<int, 160>{});
            case 128: return run_tile(std::integral_constant<int, 128>{}, std::integral_constant<int, 256>{});
            case 96: return run_tile(std::integral_constant<int, 96>{}, std::integral_constant<int, 192>{});
            case 64: return run_tile(std::integral_constant<int, 64>{}, std::integral_constant<int, 256>{});
            case 32: return run_tile(std::integral_constant<int, 32>{}, std::integral_constant<int, 256>{});
            default: break;
        }
    }
    NFB_CHECK(false, "n_components too large for the coordinate-descent kernel");
    return 1;
}

}  // namespace nfb
