// Row-sharded H half-step over NVLink peer memory (one process per GPU, CUDA IPC mappings).
//
// Every rank owns an "arena" (one cudaMalloc, mapped into all peers):
//     int    rs_flag[16]      rs_flag[p] = last epoch for which rank p's local numerator is complete
//     int    ag_flag[16]      ag_flag[p] = last epoch for which rank p's H slice (and scalars) have landed HERE
//     int    error            a wait timed out (peer died): results are garbage, the host raises
//     double scal[16][2]      scal[p] = rank p's local (violation_W, violation_H) / error cross term, pushed with the slice
//     float  gram[kp * kp]    this rank's local W'W
//     float  num[k][ld]       this rank's local W'X (the GEMM writes it here directly when it runs unsplit);
//                             rank r needs columns [r ns, (r + 1) ns) of every rank's copy
// and its H master [kp][ld] is mapped into all peers as well.
//
//   reduce-scatter  : rank r PULLS its column slice of num from every arena (7 remote reads + 1 local per element, summed in rank
//                     order -> deterministic) and the k x k Gram from every arena, instead of ncclReduceScatter +
//                     ncclAllReduce; the kernel first spins (on local memory) until every peer has posted its epoch.
//   all-gather      : after the sweep, rank r PUSHES its H columns straight into every peer's H master (no pack /
//                     unpack copies) together with its two fp64 scalars, then posts ag_flag; a one-CTA kernel waits
//                     for all peers and sums the scalars in rank order.
// Epochs only grow, flags are never reset.  Buffer reuse is safe because each phase is fenced by the other one:
// nobody can start epoch e+1's numerator before having passed the all-gather of epoch e, which needs my post,
// which I issue after my pull of epoch e (and symmetrically for the H master).
#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// spin until flags[p] >= epoch for all p < world (threads 0..world-1 of the CTA); ~30 s timeout (ranks may
// legitimately drift by the time one of them spends in a host-side step between two collective phases).
// A wait that times out raises the error flag in EVERY rank's arena, and a raised flag ends every later wait at
// once: a dead or out-of-step peer costs one timeout, not one per remaining phase.
__device__ __forceinline__ void wait_all(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch) {
    int* my = reinterpret_cast<int*>(peers.arena[rank]);
    if (threadIdx.x < world) {
        const int* flag = my + flag_offset + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < epoch) {
            if (ld_acquire_sys(my + kP2pError) != 0) break;
            if (clock64() - t0 > 60000000000LL) {
                for (int p = 0; p < world; ++p) st_release_sys(reinterpret_cast<int*>(peers.arena[p]) + kP2pError, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

__global__ void p2p_post_kernel(P2pPeers peers, int world, int rank, int flag_offset, int epoch) {
    __threadfence_system();   // the producing kernels ran before this one on the same stream
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<int*>(peers.arena[threadIdx.x]) + flag_offset + rank, epoch);
}

// out[t][j] = sum_p arena_p.num[t][c0 + j]   (t < k, j < ns; 0 past the matrix);  gram_out = sum_p arena_p.gram
__global__ void __launch_bounds__(256)
p2p_pull_kernel(P2pPeers peers, int world, int rank, int epoch, int k, int64_t ns, int64_t ld, int64_t c0,
                int64_t num_offset, int64_t gram_offset, int gram_elems, float* __restrict__ out,
                float* __restrict__ gram_out) {
    wait_all(peers, world, rank, kP2pRsFlag, epoch);
    const int64_t w4 = ns / 4;                                        // ns % 8 == 0
    const int64_t valid4 = max(static_cast<int64_t>(0), min(ns, ld - c0)) / 4;   // ld % 8 == 0, c0 % 8 == 0
    const int64_t vecs = static_cast<int64_t>(k) * w4;
    const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t v = tid; v < vecs; v += nthreads) {
        const int64_t t = v / w4, j4 = v - t * w4;
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (j4 < valid4) {
            float4 part[kP2pMaxWorld];
            const int64_t off = num_offset + t * ld + c0 + 4 * j4;
#pragma unroll
            for (int p = 0; p < kP2pMaxWorld; ++p)
                if (p < world) part[p] = ld_peer4(reinterpret_cast<const float*>(peers.arena[p]) + off);
#pragma unroll
            for (int p = 0; p < kP2pMaxWorld; ++p)
                if (p < world) { acc.x += part[p].x; acc.y += part[p].y; acc.z += part[p].z; acc.w += part[p].w; }
        }
        reinterpret_cast<float4*>(out)[v] = acc;
    }
    for (int64_t e = tid; e < gram_elems; e += nthreads) {
        float acc = 0.0f;
        for (int p = 0; p < world; ++p) acc += ld_peer(reinterpret_cast<const float*>(peers.arena[p]) + gram_offset + e);
        gram_out[e] = acc;
    }
}

// H[t][c0 + j] (t < k, j < width) -> the same place in every peer's H master; scal[2] -> slot `rank` of every arena
__global__ void __launch_bounds__(256)
p2p_push_kernel(P2pPeers peers, int world, int rank, int k, int64_t ld, int64_t c0, int64_t width,
                const double* __restrict__ scal_src, int64_t scal_offset) {
    const float* src = reinterpret_cast<const float*>(peers.hmaster[rank]);
    const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t w4 = width / 4;     // c0 % 8 == 0 and ld % 8 == 0: rows start 16-byte aligned
    for (int64_t e = tid; e < static_cast<int64_t>(k) * w4; e += nthreads) {
        const int64_t t = e / w4, j = (e - t * w4) * 4;
        const int64_t off = t * ld + c0 + j;
        const float4 v = *reinterpret_cast<const float4*>(src + off);
#pragma unroll
        for (int p = 0; p < kP2pMaxWorld; ++p)
            if (p < world && p != rank) *reinterpret_cast<float4*>(reinterpret_cast<float*>(peers.hmaster[p]) + off) = v;
    }
    const int64_t tail = width - 4 * w4;
    for (int64_t e = tid; e < static_cast<int64_t>(k) * tail; e += nthreads) {
        const int64_t t = e / tail, j = 4 * w4 + (e - t * tail);
        const int64_t off = t * ld + c0 + j;
        const float v = src[off];
        for (int p = 0; p < world; ++p)
            if (p != rank) reinterpret_cast<float*>(peers.hmaster[p])[off] = v;
    }
    if (tid < 2 * world) {
        const int p = static_cast<int>(tid) / 2, i = static_cast<int>(tid) & 1;
        double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(peers.arena[p]) + scal_offset) + 2 * rank + i;
        *dst = scal_src != nullptr ? scal_src[i] : 0.0;
    }
    __threadfence_system();
}

// wait for every peer's slice, then scal_out[i] = sum_p scal[p][i] in rank order
__global__ void p2p_gather_wait_kernel(P2pPeers peers, int world, int rank, int epoch, int64_t scal_offset,
                                       double* __restrict__ scal_out) {
    wait_all(peers, world, rank, kP2pAgFlag, epoch);
    if (threadIdx.x < 2 && scal_out != nullptr) {
        const volatile double* scal =
            reinterpret_cast<const volatile double*>(reinterpret_cast<char*>(peers.arena[rank]) + scal_offset);
        double acc = 0.0;
        for (int p = 0; p < world; ++p) acc += scal[2 * p + threadIdx.x];
        // a failed exchange poisons the scalars the host reads every iteration (the CD violation), so the
        // stopping-rule check sees it at once instead of after the whole loop
        if (ld_acquire_sys(reinterpret_cast<int*>(peers.arena[rank]) + kP2pError) != 0) acc = __longlong_as_double(0x7ff8000000000000LL);
        scal_out[threadIdx.x] = acc;
    }
}

}  // namespace

int p2p_post(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, cudaStream_t stream) {
    p2p_post_kernel<<<1, 32, 0, stream>>>(peers, world, rank, flag_offset, epoch);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_pull(const P2pPeers& peers, int world, int rank, int epoch, int k, int64_t ns, int64_t ld, int64_t c0,
             int64_t num_offset, int64_t gram_offset, int gram_elems, float* out, float* gram_out,
             cudaStream_t stream) {
    const int64_t vecs = static_cast<int64_t>(k) * ns / 4;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(vecs, 256), 4 * sm_count())));
    p2p_pull_kernel<<<grid, 256, 0, stream>>>(peers, world, rank, epoch, k, ns, ld, c0, num_offset, gram_offset,
                                              gram_elems, out, gram_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_push(const P2pPeers& peers, int world, int rank, int k, int64_t ld, int64_t c0, int64_t width,
             const double* scal_src, int64_t scal_offset, cudaStream_t stream) {
    const int64_t vecs = std::max<int64_t>(1, static_cast<int64_t>(k) * (width / 4 + 1));
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(vecs, 256), 4 * sm_count())));
    p2p_push_kernel<<<grid, 256, 0, stream>>>(peers, world, rank, k, ld, c0, width, scal_src, scal_offset);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_gather_wait(const P2pPeers& peers, int world, int rank, int epoch, int64_t scal_offset, double* scal_out,
                    cudaStream_t stream) {
    p2p_gather_wait_kernel<<<1, 32, 0, stream>>>(peers, world, rank, epoch, scal_offset, scal_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
