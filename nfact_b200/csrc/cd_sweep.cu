// Coordinate-descent (HALS) sweep of one NMF factor -- sklearn/decomposition/_cdnmf_fast.pyx:8-38.
//
// sklearn walks the components t = 0..k-1 and, for every sample i, recomputes
//     grad = -XHt[i,t] + sum_r HHt[t,r] W[i,r]                       (k FMAs per coordinate, k^2 per sample)
//     pg = W[i,t] == 0 ? min(0, grad) : grad ;  violation += |pg|
//     W[i,t] = max(W[i,t] - grad / HHt[t,t], 0)
// Samples are independent; components are Gauss-Seidel.  Here one WARP owns one sample (= one column j of
// the component-major factor F [k, cols]) and keeps the whole gradient vector g = G F[:,j] - num[:,j]
// in registers (k/32 values per lane).  g starts from the tensor-core product G.F (the same "denominator"
// GEMM the multiplicative update uses) and is kept current with a rank-1 update g += delta * G[t,:]
// only when coordinate t actually moves -- a coordinate that stays at zero costs O(1), so the sparse factors
// NFACT's L1 penalty produces are swept in ~k*nnz instead of k^2 operations, with the Gauss-Seidel
// order of the reference kept exactly.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

constexpr int kTileCols = 32;
constexpr int kWarps = 8;

template <int KPL>  // components per lane: k <= 32 * KPL
__global__ void __launch_bounds__(kWarps * 32)
cd_sweep_warp_kernel(float* __restrict__ f, int k, int64_t cols, int64_t ld, const float* __restrict__ num,
                     int num_splits, int64_t num_stride, const float* __restrict__ den, int64_t ld_part,
                     const float* __restrict__ gram, int ldg, float l1, float l2, double* __restrict__ violation,
                     unsigned int* __restrict__ rowmax_bits) {
    extern __shared__ float sm[];
    __shared__ double scratch[32];
    const int kpad = k | 1;  // odd row length -> conflict-free transposed access
    float* g_s = sm;                              // [kTileCols][kpad] gradient
    float* f_s = sm + kTileCols * kpad;           // [kTileCols][kpad] factor values
    unsigned int* max_s = reinterpret_cast<unsigned int*>(sm + 2 * kTileCols * kpad);  // [k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int u = threadIdx.x; u < k; u += blockDim.x) max_s[u] = 0u;
    float viol = 0.0f;
    const int64_t ntiles = (cols + kTileCols - 1) / kTileCols;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t j0 = tile * kTileCols;
        const int64_t j = j0 + lane;
        const bool active = j < cols;
        __syncthreads();
        // ---- phase 1: coalesced loads (row u, 32 consecutive columns), transposed into smem -----------
        for (int u = warp; u < k; u += kWarps) {
            float fv = 0.0f, gv = 0.0f;
            if (active) {
                fv = f[static_cast<int64_t>(u) * ld + j];
                float nv = 0.0f;
                for (int s = 0; s < num_splits; ++s) nv += num[s * num_stride + static_cast<int64_t>(u) * ld_part + j];
                const float dv = den[static_cast<int64_t>(u) * ld_part + j];
                gv = fmaf(l2, fv, dv) - (nv - l1);
            }
            g_s[lane * kpad + u] = gv;
            f_s[lane * kpad + u] = fv;
        }
        __syncthreads();
        // ---- phase 2: one warp per column, components in order -----------------------------------------
        for (int jj = warp; jj < kTileCols; jj += kWarps) {
            if (j0 + jj >= cols) break;
            float g[KPL], fv[KPL];
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int u = lane + 32 * i;
                g[i] = u < k ? g_s[jj * kpad + u] : 0.0f;
                fv[i] = u < k ? f_s[jj * kpad + u] : 0.0f;
            }
            for (int t = 0; t < k; ++t) {
                const int owner = t & 31, slot = t >> 5;
                float gt = 0.0f, ft = 0.0f;
#pragma unroll
                for (int i = 0; i < KPL; ++i)
                    if (i == slot) { gt = g[i]; ft = fv[i]; }
                gt = __shfl_sync(0xffffffffu, gt, owner);
                ft = __shfl_sync(0xffffffffu, ft, owner);
                const float* grow = gram + static_cast<int64_t>(t) * ldg;
                const float hess = __ldg(grow + t) + l2;
                const float pg = (ft == 0.0f) ? fminf(0.0f, gt) : gt;
                viol += fabsf(pg);
                if (hess != 0.0f) {
                    const float fn = fmaxf(ft - gt / hess, 0.0f);
                    const float delta = fn - ft;
                    if (delta != 0.0f) {  // warp-uniform
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int u = lane + 32 * i;
                            // components below t are final for this sweep; skip their chunks
                            if (32 * i + 31 >= t && u < k) g[i] = fmaf(delta, __ldg(grow + u), g[i]);
                            if (i == slot && lane == owner) fv[i] = fn;
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int u = lane + 32 * i;
                if (u < k) f_s[jj * kpad + u] = fv[i];
            }
        }
        __syncthreads();
        // ---- phase 3: coalesced store + per-row maxima -------------------------------------------------
        for (int u = warp; u < k; u += kWarps) {
            const float v = active ? f_s[lane * kpad + u] : 0.0f;
            if (active) f[static_cast<int64_t>(u) * ld + j] = v;
            const float m = warp_max(v);
            if (lane == 0 && m > 0.0f) atomicMax(&max_s[u], __float_as_uint(m));
        }
    }
    __syncthreads();
    for (int u = threadIdx.x; u < k; u += blockDim.x)
        if (max_s[u] != 0u) atomicMax(&rowmax_bits[u], max_s[u]);
    // every lane of a warp accumulated the same violation; count it once per warp
    const double total = block_sum(lane == 0 ? static_cast<double>(viol) : 0.0, scratch);
    if (threadIdx.x == 0 && total != 0.0) atomicAdd(violation, total);
}


// ---- fast path: Gram matrix resident in shared memory (k <= ~216) ----------------------------------
// One CTA per SM, 16 warps, one column per warp at a time.  Per component step the dependent chain is
// shfl -> multiply by the precomputed 1/hess -> clamp -> rank-1 update from smem, ~20 instructions.
constexpr int kFastWarps = 24;   // 6 warps per scheduler hide the per-step dependency chain
constexpr int kFastTile = 24;    // one column per warp

template <int KPL>
__global__ void __launch_bounds__(kFastWarps * 32, 1)
cd_sweep_smem_kernel(float* __restrict__ f, int k, int64_t cols, int64_t ld, const float* __restrict__ num,
                     int num_splits, int64_t num_stride, const float* __restrict__ den, int64_t ld_part,
                     const float* __restrict__ gram, int ldg, float l1, float l2, double* __restrict__ violation,
                     unsigned int* __restrict__ rowmax_bits) {
    extern __shared__ float sm[];
    __shared__ double scratch[32];
    const int kpad = k | 1;
    constexpr int kRow = 32 * KPL;                        // zero-padded Gram row: no bounds predicate
    float* gram_s = sm;                                   // [k][kRow]
    float* invh_s = gram_s + k * kRow;                    // [k]  1 / (G[t][t] + l2), 0 when the hessian is 0
    float* g_s = invh_s + k;                              // [kFastTile][kpad]
    float* f_s = g_s + kFastTile * kpad;                  // [kFastTile][kpad]
    unsigned int* max_s = reinterpret_cast<unsigned int*>(f_s + kFastTile * kpad);  // [k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int idx = threadIdx.x; idx < k * kRow; idx += blockDim.x) {
        const int t = idx / kRow, u = idx - t * kRow;
        gram_s[idx] = u < k ? gram[static_cast<int64_t>(t) * ldg + u] : 0.0f;
    }
    for (int u = threadIdx.x; u < k; u += blockDim.x) {
        const float hess = gram[static_cast<int64_t>(u) * ldg + u] + l2;
        invh_s[u] = hess != 0.0f ? 1.0f / hess : 0.0f;
        max_s[u] = 0u;
    }
    float viol = 0.0f;
    const int64_t ntiles = (cols + kFastTile - 1) / kFastTile;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t j0 = tile * kFastTile;
        __syncthreads();
        // ---- phase 1: consecutive threads read consecutive columns of one component row ------------------
#pragma unroll 4
        for (int idx = threadIdx.x; idx < k * kFastTile; idx += kFastWarps * 32) {
            const int u = idx / kFastTile, c = idx - u * kFastTile;
            const int64_t j = j0 + c;
            float fv = 0.0f, gv = 0.0f;
            if (j < cols) {
                fv = f[static_cast<int64_t>(u) * ld + j];
                float nv = 0.0f;
                for (int s = 0; s < num_splits; ++s)
                    nv += __ldcs(num + s * num_stride + static_cast<int64_t>(u) * ld_part + j);
                const float dv = __ldcs(den + static_cast<int64_t>(u) * ld_part + j);
                gv = fmaf(l2, fv, dv) - (nv - l1);
            }
            g_s[c * kpad + u] = gv;
            f_s[c * kpad + u] = fv;
        }
        __syncthreads();
        // ---- phase 2: one warp = one column ------------------------------------------------------------
        if (j0 + warp < cols) {
            float g[KPL], fv[KPL];
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int u = lane + 32 * i;
                g[i] = u < k ? g_s[warp * kpad + u] : 0.0f;
                fv[i] = u < k ? f_s[warp * kpad + u] : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int t_end = min(32, k - 32 * i);
                for (int o = 0; o < t_end; ++o) {
                    const int t = 32 * i + o;
                    const float gt = __shfl_sync(0xffffffffu, g[i], o);
                    const float ft = __shfl_sync(0xffffffffu, fv[i], o);
                    const float inv = invh_s[t];
                    viol += fabsf((ft == 0.0f) ? fminf(0.0f, gt) : gt);
                    const float fn = fmaxf(fmaf(-gt, inv, ft), 0.0f);
                    const float delta = (inv != 0.0f) ? fn - ft : 0.0f;
                    if (delta != 0.0f) {  // warp-uniform
                        const float* grow = gram_s + t * kRow + lane;
#pragma unroll
                        for (int i2 = i; i2 < KPL; ++i2) g[i2] = fmaf(delta, grow[32 * i2], g[i2]);
                        if (lane == o) fv[i] = fn;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int u = lane + 32 * i;
                if (u < k) f_s[warp * kpad + u] = fv[i];
            }
        }
        __syncthreads();
        // ---- phase 3: store + per-row maxima --------------------------------------------------------------
        for (int idx = threadIdx.x; idx < k * kFastTile; idx += kFastWarps * 32) {
            const int u = idx / kFastTile, c = idx - u * kFastTile;
            const int64_t j = j0 + c;
            if (j < cols) {
                const float v = f_s[c * kpad + u];
                f[static_cast<int64_t>(u) * ld + j] = v;
                const unsigned int bits = __float_as_uint(v);
                if (bits > max_s[u]) atomicMax(&max_s[u], bits);   // v >= 0: float order == uint order
            }
        }
    }
    __syncthreads();
    for (int u = threadIdx.x; u < k; u += blockDim.x)
        if (max_s[u] != 0u) atomicMax(&rowmax_bits[u], max_s[u]);
    const double total = block_sum(lane == 0 ? static_cast<double>(viol) : 0.0, scratch);
    if (threadIdx.x == 0 && total != 0.0) atomicAdd(violation, total);
}


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
                     unsigned int* __restrict__ rowmax_bits, bool vec) {
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

    auto load_g = [&](int t, int u) -> float {             // G[t][u], zero outside the k x k matrix
        return (t < k && u < k) ? __ldg(gram + static_cast<int64_t>(t) * ldg + u) : 0.0f;
    };
    auto load_hess = [&](int t0) -> float {
        return (tid < kBlk && t0 + tid < k) ? __ldg(gram + static_cast<int64_t>(t0 + tid) * ldg + t0 + tid) + l2 : 0.0f;
    };

    // ---- block 0 operands, then the gradient seed g = G F[:, j] + l2 F[:, j] - (num[:, j] - l1) -----------
#pragma unroll 4
    for (int i = 0; i < kBlk; ++i)
        for (int u = tid; u < kp4; u += NT) gb[i * kp4 + u] = load_g(i, u);
    if (tid < kBlk) {
        const float hess = load_hess(0);
        invh[tid] = hess != 0.0f ? 1.0f / hess : 0.0f;
        max_s[tid] = 0u;
    }
    {
        // every thread owns 4 consecutive samples of one component row per pass; all (splits + 2) 128-bit loads
        // of a pass are issued before the first use so the phase runs at memory bandwidth, not latency
        constexpr int kSeedBatch = 8;
        const int c4 = tid % JG;
        const int64_t jc = static_cast<int64_t>(blockIdx.x) * TJ + 4 * c4;
        const bool full = vec && jc + 3 < cols;
        for (int u = tid / JG; u < kp4; u += UG) {
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
    // factor values of the block in flight; loads are unconditional (clamped indices) and zeroed afterwards so
    // that the 16 of them are independent straight-line loads
    const int64_t jc_own = active ? j : 0;
    const int u_own = min(tid, k - 1);
    float fl[kBlk];
#pragma unroll
    for (int i = 0; i < kBlk; ++i) {
        const float v = f[static_cast<int64_t>(min(i, k - 1)) * ld + jc_own];
        fl[i] = (active && i < k) ? v : 0.0f;
    }
    __syncthreads();

    float viol = 0.0f;
    const int jg = tid % JG, ug = tid / JG;
    // one block of coordinate steps for sample `tid`; kFull: all 16 components exist (no bounds predicates)
    auto step_block = [&](auto full_tag, int t0, int nb) {
        constexpr bool kFull = decltype(full_tag)::value;
        float gl[kBlk];
#pragma unroll
        for (int i = 0; i < kBlk; ++i) gl[i] = (kFull || i < nb) ? gs[(t0 + i) * TJ + tid] : 0.0f;
        float* fp = f + static_cast<int64_t>(t0) * ld + jc_own;
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            if (kFull || i < nb) {
                const float ft = fl[i];
                const float gt = gl[i];
                viol += fabsf((ft == 0.0f) ? fminf(0.0f, gt) : gt);
                const float inv = invh[i];
                const float fn = fmaxf(fmaf(-gt, inv, ft), 0.0f);
                const float delta = inv != 0.0f ? fn - ft : 0.0f;
                if (active && delta != 0.0f) *fp = fn;
                fp += ld;
                // values are >= 0, so unsigned order == float order
                const unsigned int vmax = __reduce_max_sync(0xffffffffu, active ? __float_as_uint(ft + delta) : 0u);
                if ((tid & 31) == 0 && vmax != 0u) atomicMax(&max_s[i], vmax);
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
                gs[(t0 + i) * TJ + tid] = delta;           // the gradient row is dead: it now carries the deltas
            }
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
#pragma unroll
            for (int i = 0; i < kBlk; ++i) {
                const int t = t0 + kBlk + i;
                const float v = __ldg(gram + static_cast<int64_t>(min(t, k - 1)) * ldg + u_own);
                pre[i] = (t < k && tid < k) ? v : 0.0f;
            }
            hess_next = load_hess(t0 + kBlk);
#pragma unroll
            for (int i = 0; i < kBlk; ++i) {
                const int t = t0 + kBlk + i;
                const float v = f[static_cast<int64_t>(min(t, k - 1)) * ld + jc_own];
                fl[i] = (active && t < k) ? v : 0.0f;
            }
        }
        __syncthreads();
        if (tid < kBlk) {
            if (t0 + tid < k && max_s[tid] != 0u) atomicMax(&rowmax_bits[t0 + tid], max_s[tid]);
            max_s[tid] = 0u;
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
        if (tid < kp4) {
#pragma unroll
            for (int i = 0; i < kBlk; ++i) gb[i * kp4 + tid] = pre[i];
        }
        if (kp4 > NT) {                                    // columns the register prefetch does not cover
            for (int i = 0; i < kBlk; ++i)
                for (int u = tid + NT; u < kp4; u += NT) gb[i * kp4 + u] = load_g(t0 + kBlk + i, u);
        }
        if (tid < kBlk) invh[tid] = hess_next != 0.0f ? 1.0f / hess_next : 0.0f;
        __syncthreads();
    }
    const double total = block_sum(active ? static_cast<double>(viol) : 0.0, scratch);
    if (tid == 0 && total != 0.0) atomicAdd(violation, total);
}

}  // namespace

int cd_sweep(float* f, int64_t k, int64_t cols, int64_t ld, const float* num, int num_splits, int64_t num_stride,
             const float* den, int64_t ld_part, const float* gram, float l1, float l2, double* violation,
             unsigned int* rowmax_bits, cudaStream_t stream) {
    if (k * cols == 0) return 0;
    NFB_CHECK(k <= 512, "n_components > 512 is not supported by the coordinate-descent kernel");
    const int ldg = static_cast<int>(round_up64(k, 16));
    const int kpad = static_cast<int>(k) | 1;
    // preferred: blocked sweep.  Pick the sample tile that minimises waves x tile (one CTA per SM).
    if (getenv("NFB_CD_KERNEL") == nullptr || atoi(getenv("NFB_CD_KERNEL")) == 0) {
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
                violation, rowmax_bits, vec);
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
    const int kpl_fast = k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : 7;
    const size_t smem_fast = (static_cast<size_t>(k) * 32 * kpl_fast + 2 * k +
                              static_cast<size_t>(2) * kFastTile * kpad) * sizeof(float);
    if (k <= 32 * kpl_fast && smem_fast <= 224 * 1024) {
        const int64_t ntiles_fast = ceil_div64(cols, kFastTile);
        const int grid_fast = static_cast<int>(std::min<int64_t>(ntiles_fast, sm_count()));
#define NFB_CD_FAST(KPL)                                                                                         \
    do {                                                                                                         \
        auto kernel = cd_sweep_smem_kernel<KPL>;                                                                 \
        NFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                      static_cast<int>(smem_fast)));                                             \
        kernel<<<grid_fast, kFastWarps * 32, smem_fast, stream>>>(f, static_cast<int>(k), cols, ld, num,         \
                                                                  num_splits, num_stride, den, ld_part, gram,    \
                                                                  ldg, l1, l2, violation, rowmax_bits);          \
    } while (0)
        if (k <= 32) NFB_CD_FAST(1);
        else if (k <= 64) NFB_CD_FAST(2);
        else if (k <= 128) NFB_CD_FAST(4);
        else NFB_CD_FAST(7);
#undef NFB_CD_FAST
        NFB_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = (static_cast<size_t>(2) * kTileCols * kpad + k) * sizeof(float);
    const int64_t ntiles = ceil_div64(cols, kTileCols);
    NFB_CHECK(smem <= 220 * 1024, "n_components too large for the coordinate-descent kernel");
    const int blocks_per_sm = std::max<int>(1, std::min<int>(4, static_cast<int>((200 * 1024) / (smem + 1024))));
    const int grid = static_cast<int>(std::min<int64_t>(ntiles, static_cast<int64_t>(sm_count()) * blocks_per_sm));
#define NFB_CD_LAUNCH(KPL)                                                                                       \
    do {                                                                                                         \
        auto kernel = cd_sweep_warp_kernel<KPL>;                                                                 \
        NFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
        kernel<<<grid, kWarps * 32, smem, stream>>>(f, static_cast<int>(k), cols, ld, num, num_splits, num_stride, \
                                                    den, ld_part, gram, ldg, l1, l2, violation, rowmax_bits);    \
    } while (0)
    if (k <= 32) NFB_CD_LAUNCH(1);
    else if (k <= 64) NFB_CD_LAUNCH(2);
    else if (k <= 128) NFB_CD_LAUNCH(4);
    else if (k <= 256) NFB_CD_LAUNCH(8);
    else NFB_CD_LAUNCH(16);
#undef NFB_CD_LAUNCH
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
