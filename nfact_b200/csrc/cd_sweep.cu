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


// ---- thread-per-sample blocked sweep ------------------------------------------------------------------
// One THREAD owns one sample (column j of F).  Its gradient vector lives in shared memory as gs[u][tid]
// (conflict-free), components are processed in blocks of 16: the 16 coordinate steps of a block run in
// registers against the 16 x 16 diagonal block of G, then the block's 16 deltas are applied to the rest of
// the gradient, g[u] += sum_i delta_i G[t0+i][u], as 16 FMAs per entry with G broadcast from shared memory.
// Same arithmetic as the rank-1 formulation above (exact Gauss-Seidel order), but the per-coordinate
// bookkeeping is paid per thread instead of per warp: ~k^2/2 FMAs per sample at FMA-pipe throughput.
constexpr int kBlk = 16;

template <int TJ>
__global__ void __launch_bounds__(TJ)
cd_sweep_rows_kernel(float* __restrict__ f, int k, int64_t cols, int64_t ld, const float* __restrict__ num,
                     int num_splits, int64_t num_stride, const float* __restrict__ den, int64_t ld_part,
                     const float* __restrict__ gram, int ldg, float l1, float l2, double* __restrict__ violation,
                     unsigned int* __restrict__ rowmax_bits) {
    extern __shared__ float sm[];
    __shared__ double scratch[32];
    __shared__ float invh[kBlk];
    float* gs = sm;                                   // [k][TJ]
    float* gb = sm + static_cast<size_t>(k) * TJ;     // [k][kBlk]: gb[u][i] = G[t0 + i][u]
    const int tid = threadIdx.x;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * TJ + tid;
    const bool active = j < cols;
    // ---- gradient seed: g = G F[:, j] + l2 F[:, j] - (num[:, j] - l1) ------------------------------------
#pragma unroll 4
    for (int u = 0; u < k; ++u) {
        float gv = 0.0f;
        if (active) {
            const float fv = f[static_cast<int64_t>(u) * ld + j];
            float nv = 0.0f;
            for (int s = 0; s < num_splits; ++s) nv += __ldcs(num + s * num_stride + static_cast<int64_t>(u) * ld_part + j);
            gv = fmaf(l2, fv, __ldcs(den + static_cast<int64_t>(u) * ld_part + j)) - (nv - l1);
        }
        gs[u * TJ + tid] = gv;
    }
    float viol = 0.0f;
    for (int t0 = 0; t0 < k; t0 += kBlk) {
        const int nb = min(kBlk, k - t0);
        __syncthreads();   // previous block's readers of gb / invh are done
        for (int e = tid; e < kBlk * k; e += TJ) {
            const int i = e / k, u = e - i * k;
            gb[u * kBlk + i] = i < nb ? gram[static_cast<int64_t>(t0 + i) * ldg + u] : 0.0f;
        }
        if (tid < kBlk) {
            const float hess = tid < nb ? gram[static_cast<int64_t>(t0 + tid) * ldg + t0 + tid] + l2 : 0.0f;
            invh[tid] = hess != 0.0f ? 1.0f / hess : 0.0f;
        }
        __syncthreads();
        // ---- the 16 coordinate steps of this block, in registers -----------------------------------------
        float gl[kBlk], dl[kBlk], fl[kBlk];
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            gl[i] = i < nb ? gs[(t0 + i) * TJ + tid] : 0.0f;
            fl[i] = (i < nb && active) ? f[static_cast<int64_t>(t0 + i) * ld + j] : 0.0f;   // 16 loads in flight
        }
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            float delta = 0.0f;
            if (i < nb) {
                const float ft = fl[i];
                const float gt = gl[i];
                viol += fabsf((ft == 0.0f) ? fminf(0.0f, gt) : gt);
                const float inv = invh[i];
                const float fn = fmaxf(fmaf(-gt, inv, ft), 0.0f);
                delta = inv != 0.0f ? fn - ft : 0.0f;
                if (active && delta != 0.0f) f[static_cast<int64_t>(t0 + i) * ld + j] = fn;
                const float vmax = warp_max(active ? ft + delta : 0.0f);
                if ((tid & 31) == 0 && vmax > 0.0f) atomicMax(&rowmax_bits[t0 + i], __float_as_uint(vmax));
#pragma unroll
                for (int i2 = i + 1; i2 < kBlk; ++i2) gl[i2] = fmaf(delta, gb[(t0 + i2) * kBlk + i], gl[i2]);
            }
            dl[i] = delta;
        }
        // ---- apply the block's deltas to the components still to come ------------------------------------
#pragma unroll 4
        for (int u = t0 + nb; u < k; ++u) {
            const float4* grow = reinterpret_cast<const float4*>(gb + u * kBlk);
            const float4 a = grow[0], b = grow[1], c = grow[2], d = grow[3];
            float acc = gs[u * TJ + tid];
            acc = fmaf(dl[0], a.x, acc);  acc = fmaf(dl[1], a.y, acc);  acc = fmaf(dl[2], a.z, acc);  acc = fmaf(dl[3], a.w, acc);
            acc = fmaf(dl[4], b.x, acc);  acc = fmaf(dl[5], b.y, acc);  acc = fmaf(dl[6], b.z, acc);  acc = fmaf(dl[7], b.w, acc);
            acc = fmaf(dl[8], c.x, acc);  acc = fmaf(dl[9], c.y, acc);  acc = fmaf(dl[10], c.z, acc); acc = fmaf(dl[11], c.w, acc);
            acc = fmaf(dl[12], d.x, acc); acc = fmaf(dl[13], d.y, acc); acc = fmaf(dl[14], d.z, acc); acc = fmaf(dl[15], d.w, acc);
            gs[u * TJ + tid] = acc;
        }
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
    // preferred: thread-per-sample blocked sweep (largest tile whose gradient block fits in shared memory)
    if (getenv("NFB_CD_KERNEL") == nullptr || atoi(getenv("NFB_CD_KERNEL")) == 0) {
        auto smem_rows = [&](int tj) { return (static_cast<size_t>(k) * tj + static_cast<size_t>(k) * kBlk) * sizeof(float); };
        auto run_rows = [&](auto tj_tag) -> int {
            constexpr int TJ = decltype(tj_tag)::value;
            auto kernel = cd_sweep_rows_kernel<TJ>;
            const size_t smem = smem_rows(TJ);
            NFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kernel<<<static_cast<unsigned>(ceil_div64(cols, TJ)), TJ, smem, stream>>>(
                f, static_cast<int>(k), cols, ld, num, num_splits, num_stride, den, ld_part, gram, ldg, l1, l2,
                violation, rowmax_bits);
            NFB_CUDA(cudaGetLastError());
            return 0;
        };
        if (smem_rows(256) <= 224 * 1024 && cols >= 256 * static_cast<int64_t>(sm_count()))
            return run_rows(std::integral_constant<int, 256>{});
        if (smem_rows(128) <= 110 * 1024) return run_rows(std::integral_constant<int, 128>{});
        if (smem_rows(256) <= 224 * 1024) return run_rows(std::integral_constant<int, 256>{});
        if (smem_rows(128) <= 224 * 1024) return run_rows(std::integral_constant<int, 128>{});
        if (smem_rows(64) <= 224 * 1024) return run_rows(std::integral_constant<int, 64>{});
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
