// Row-sharded H half-step over NVLink peer memory (one process per GPU, CUDA IPC mappings).
//
// Every rank owns an "arena" (one cudaMalloc, mapped into all peers):
//     int    rs_flag[16]        rs_flag[p] = last epoch for which rank p's numerator tiles and Gram have landed HERE
//     int    ag_flag[16]        ag_flag[p] = last epoch for which rank p's H slice (and scalars) have landed HERE
//     int    gram_flag[16]      gram_flag[p] = last epoch for which rank p's W'W has landed HERE (posted before the
//                               W'X GEMM starts, so that the Gram sum and the denominator GEMM that needs it run on
//                               the aux stream beside the GEMM instead of after it)
//     int    error              a wait timed out (peer died): results are garbage, the host raises
//     double scal[16][2]        scal[p] = rank p's local (violation_W, violation_H), posted with its H slice
//     float  gram_in[P][kp*kp]  gram_in[p] = rank p's local W'W
//     uint   rowmax_in[P][kp]   rowmax_in[p] = per-component maxima of rank p's H slice (bit patterns)
//     float  num_in[P][k][ns]   num_in[p] = rank p's local W'X restricted to MY H columns [rank ns, (rank+1) ns)
// and its H master [kp][ld] is mapped into all peers as well.  Nothing here is a separate "collective" pass:
//
//   reduce-scatter  : the W'X GEMM's epilogue stores every finished tile straight into the inbox (num_in[rank]) of
//                     the rank that owns those H columns (gemm_split.cu, GemmScatter) -- remote stores over NVLink
//                     that overlap the mainloop of the following tiles; the consumer (sweep / multiplicative update)
//                     reads the P slabs as split-K partials, summed in rank order -> deterministic.  The k x k Gram
//                     is pushed to every peer by a small kernel before the GEMM and flagged on its own (gram_flag).
//                     After the GEMM a one-CTA kernel posts rs_flag; the owner waits for all P flags.
//   all-gather      : the sweep / update kernel writes every H value it changes into all peers' H masters as part
//                     of its own write-back (cd_sweep.cu / update.cu, PeerMirror); the post kernel adds this rank's
//                     two float64 scalars and per-component maxima and raises ag_flag; a one-CTA kernel waits for
//                     all peers, sums the scalars in rank order and takes the maxima (the plane split of the
//                     gathered H then needs no pass of its own to find them).
// Epochs only grow, flags are never reset.  Buffer reuse is safe because each phase is fenced by the other one:
// nobody can start epoch e+1's numerator before having passed the all-gather of epoch e, which needs my post,
// which I issue after my sweep of epoch e has read its inbox (and symmetrically for the H master).
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

// payloads (nullable) -> slot `rank` of every peer's arena, then fence + flag[rank] = epoch everywhere
__global__ void __launch_bounds__(256)
p2p_post_kernel(P2pPeers peers, int world, int rank, int flag_offset, int epoch, const double* __restrict__ scal_src,
                int64_t scal_offset, const unsigned int* __restrict__ rowmax_src, int64_t rowmax_offset, int kp) {
    if (scal_src != nullptr && threadIdx.x < 2 * world) {
        const int p = threadIdx.x / 2, i = threadIdx.x & 1;
        double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(peers.arena[p]) + scal_offset) + 2 * rank + i;
        *dst = scal_src[i];
    }
    if (rowmax_src != nullptr) {
        for (int e = threadIdx.x; e < world * kp; e += blockDim.x) {
            const int p = e / kp, t = e - p * kp;
            unsigned int* dst = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(peers.arena[p]) + rowmax_offset);
            dst[static_cast<int64_t>(rank) * kp + t] = rowmax_src[t];
        }
    }
    __threadfence_system();   // the payload above and whatever the producing kernels stored before this one
    __syncthreads();
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<int*>(peers.arena[threadIdx.x]) + flag_offset + rank, epoch);
}

// p2p_post_kernel + p2p_wait_kernel in one launch (the reduce-scatter's flag round trip)
__global__ void __launch_bounds__(256)
p2p_post_wait_kernel(P2pPeers peers, int world, int rank, int flag_offset, int epoch) {
    __threadfence_system();   // whatever the producing kernels stored before this one
    __syncthreads();
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<int*>(peers.arena[threadIdx.x]) + flag_offset + rank, epoch);
    wait_all(peers, world, rank, flag_offset, epoch);
}

// p2p_post_kernel (scalars + row maxima + flag) and p2p_gather_wait_kernel in one launch (the all-gather's round trip)
__global__ void __launch_bounds__(256)
p2p_post_gather_kernel(P2pPeers peers, int world, int rank, int epoch, const double* scal_src,
                       int64_t scal_offset, double* scal_out, unsigned int* rowmax,
                       int64_t rowmax_offset, int kp) {
    if (scal_src != nullptr && threadIdx.x < 2 * world) {
        const int p = threadIdx.x / 2, i = threadIdx.x & 1;
        double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(peers.arena[p]) + scal_offset) + 2 * rank + i;
        *dst = scal_src[i];
    }
    for (int e = threadIdx.x; e < world * kp; e += blockDim.x) {
        const int p = e / kp, t = e - p * kp;
        unsigned int* dst = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(peers.arena[p]) + rowmax_offset);
        dst[static_cast<int64_t>(rank) * kp + t] = rowmax[t];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<int*>(peers.arena[threadIdx.x]) + kP2pAgFlag + rank, epoch);
    wait_all(peers, world, rank, kP2pAgFlag, epoch);
    if (threadIdx.x < 2 && scal_out != nullptr) {
        const volatile double* scal =
            reinterpret_cast<const volatile double*>(reinterpret_cast<char*>(peers.arena[rank]) + scal_offset);
        double acc = 0.0;
        for (int p = 0; p < world; ++p) acc += scal[2 * p + threadIdx.x];
        if (ld_acquire_sys(reinterpret_cast<int*>(peers.arena[rank]) + kP2pError) != 0) acc = __longlong_as_double(0x7ff8000000000000LL);
        scal_out[threadIdx.x] = acc;
    }
    const volatile unsigned int* in =
        reinterpret_cast<const volatile unsigned int*>(reinterpret_cast<char*>(peers.arena[rank]) + rowmax_offset);
    for (int t = threadIdx.x; t < kp; t += blockDim.x) {
        unsigned int m = 0u;
        for (int p = 0; p < world; ++p) m = max(m, in[static_cast<int64_t>(p) * kp + t]);   // values >= 0: uint order
        rowmax[t] = m;
    }
}

// this rank's Gram (slot `rank` of its own gram_in) -> slot `rank` of every peer's gram_in
__global__ void __launch_bounds__(256)
p2p_push_gram_kernel(P2pPeers peers, int world, int rank, int64_t gram_offset, int gram_elems) {
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(peers.arena[rank]) + gram_offset) +
                        static_cast<int64_t>(rank) * (gram_elems / 4);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < gram_elems / 4; e += gridDim.x * blockDim.x) {
        const float4 v = src[e];
#pragma unroll
        for (int p = 0; p < kP2pMaxWorld; ++p)
            if (p < world && p != rank)
                (reinterpret_cast<float4*>(reinterpret_cast<char*>(peers.arena[p]) + gram_offset) +
                 static_cast<int64_t>(rank) * (gram_elems / 4))[e] = v;
    }
}

// wait for every rank's numerator tiles + Gram of `epoch`; gram_out = sum_p gram_in[p] (rank order)
__global__ void __launch_bounds__(256)
p2p_wait_sum_gram_kernel(P2pPeers peers, int world, int rank, int flag_offset, int epoch, int64_t gram_offset,
                         int gram_elems, float* __restrict__ gram_out) {
    wait_all(peers, world, rank, flag_offset, epoch);
    const float* in = reinterpret_cast<const float*>(reinterpret_cast<const char*>(peers.arena[rank]) + gram_offset);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < gram_elems; e += gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int p = 0; p < world; ++p) acc += ld_peer(in + static_cast<int64_t>(p) * gram_elems + e);
        gram_out[e] = acc;
    }
}

// wait only: every rank has raised flag row `flag_offset` to `epoch`
__global__ void __launch_bounds__(32)
p2p_wait_kernel(P2pPeers peers, int world, int rank, int flag_offset, int epoch) {
    wait_all(peers, world, rank, flag_offset, epoch);
}

// wait for every peer's slice, then scal_out[i] = sum_p scal[p][i] in rank order, rowmax_out[t] = max_p rowmax_in[p][t]
__global__ void __launch_bounds__(256)
p2p_gather_wait_kernel(P2pPeers peers, int world, int rank, int epoch, int64_t scal_offset,
                       double* __restrict__ scal_out, int64_t rowmax_offset, int kp,
                       unsigned int* __restrict__ rowmax_out) {
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
    if (rowmax_out != nullptr) {
        const volatile unsigned int* in =
            reinterpret_cast<const volatile unsigned int*>(reinterpret_cast<char*>(peers.arena[rank]) + rowmax_offset);
        for (int t = threadIdx.x; t < kp; t += blockDim.x) {
            unsigned int m = 0u;
            for (int p = 0; p < world; ++p) m = max(m, in[static_cast<int64_t>(p) * kp + t]);   // values >= 0: uint order
            rowmax_out[t] = m;
        }
    }
}

}  // namespace

int p2p_post(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, cudaStream_t stream,
             const double* scal_src, int64_t scal_offset, const unsigned int* rowmax_src, int64_t rowmax_offset,
             int kp) {
    p2p_post_kernel<<<1, 256, 0, stream>>>(peers, world, rank, flag_offset, epoch, scal_src, scal_offset, rowmax_src,
                                           rowmax_offset, kp);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_post_wait(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, cudaStream_t stream) {
    p2p_post_wait_kernel<<<1, 256, 0, stream>>>(peers, world, rank, flag_offset, epoch);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_post_gather(const P2pPeers& peers, int world, int rank, int epoch, const double* scal_src, int64_t scal_offset,
                    double* scal_out, unsigned int* rowmax, int64_t rowmax_offset, int kp, cudaStream_t stream) {
    p2p_post_gather_kernel<<<1, 256, 0, stream>>>(peers, world, rank, epoch, scal_src, scal_offset, scal_out, rowmax,
                                                  rowmax_offset, kp);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_push_gram(const P2pPeers& peers, int world, int rank, int64_t gram_offset, int gram_elems,
                  cudaStream_t stream) {
    NFB_CHECK(gram_elems % 4 == 0, "Gram size must be a multiple of 4 floats");
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(gram_elems / 4, 256), sm_count())));
    p2p_push_gram_kernel<<<grid, 256, 0, stream>>>(peers, world, rank, gram_offset, gram_elems);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_wait_sum_gram(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, int64_t gram_offset,
                      int gram_elems, float* gram_out, cudaStream_t stream) {
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(gram_elems, 256), sm_count())));
    p2p_wait_sum_gram_kernel<<<grid, 256, 0, stream>>>(peers, world, rank, flag_offset, epoch, gram_offset, gram_elems,
                                                       gram_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_wait(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, cudaStream_t stream) {
    p2p_wait_kernel<<<1, 32, 0, stream>>>(peers, world, rank, flag_offset, epoch);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int p2p_gather_wait(const P2pPeers& peers, int world, int rank, int epoch, int64_t scal_offset, double* scal_out,
                    int64_t rowmax_offset, int kp, unsigned int* rowmax_out, cudaStream_t stream) {
    p2p_gather_wait_kernel<<<1, 256, 0, stream>>>(peers, world, rank, epoch, scal_offset, scal_out, rowmax_offset, kp,
                                                  rowmax_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
