// fp32 -> (fp16 hi, fp16 lo) operand planes for the split-precision tensor-core GEMM.
//
// value * s = hi + lo + r,  s a power of two chosen so that the largest magnitude lands in
// [2^14, 2^15);  hi = rn_fp16(value*s), lo = rn_fp16(value*s - hi)  =>  |r| <= 2^-22 |value*s|
// for values down to 2^-3 of full scale and <= 2^-25 absolute (2^-40 of full scale) below that.
// X is split once per matrix (one global scale: both X and X^T are read from the same planes);
// the factors W^T / H / Gram are re-split after every update with one scale per component row.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

__global__ void absmax_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                              unsigned int* __restrict__ out_bits) {
    float m = 0.0f;
    const int64_t total = rows * cols;
    if (ld == cols) {  // contiguous: no index arithmetic
        for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x)
            m = fmaxf(m, fabsf(x[i]));
    } else {
        for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            const int64_t r = i / cols, c = i - r * cols;
            m = fmaxf(m, fabsf(x[r * ld + c]));
        }
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

__global__ void rowabsmax_kernel(const float* __restrict__ x, int64_t cols, int64_t ld,
                                 const float* __restrict__ col_mul, unsigned int* __restrict__ rowmax_bits) {
    const int64_t r = blockIdx.y;
    float m = 0.0f;
    for (int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; c < cols;
         c += static_cast<int64_t>(gridDim.x) * blockDim.x)
        m = fmaxf(m, fabsf(x[r * ld + c] * (col_mul ? col_mul[c] : 1.0f)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(&rowmax_bits[r], __float_as_uint(m));
}

// one thread = 8 consecutive columns of one row
template <bool kPerRowScale>
__global__ void split_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                             const float* __restrict__ col_mul, const unsigned int* __restrict__ max_bits,
                             __half* __restrict__ hi,
                             __half* __restrict__ lo, int64_t ldp, float* __restrict__ inv_scale,
                             double* __restrict__ sumsq, int* __restrict__ neg_flag) {
    __shared__ double scratch[32];
    const int64_t groups_per_row = ldp / 8;
    const int64_t total = rows * groups_per_row;
    double acc = 0.0, acc_sum = 0.0;
    for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
         g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = g / groups_per_row;
        const int64_t c0 = (g - r * groups_per_row) * 8;
        const float s = pow2_scale_for(__uint_as_float(max_bits[kPerRowScale ? r : 0]));
        if (c0 == 0 && (kPerRowScale || r == 0)) inv_scale[kPerRowScale ? r : 0] = 1.0f / s;
        __align__(16) __half h[8];
        __align__(16) __half l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t c = c0 + i;
            float v = c < cols ? x[r * ld + c] : 0.0f;
            if (col_mul != nullptr && c < cols) v *= col_mul[c];
            acc += static_cast<double>(v) * static_cast<double>(v);
            acc_sum += static_cast<double>(v);
            if (neg_flag != nullptr && v < 0.0f) *neg_flag = 1;
            const float vs = v * s;
            const __half vh = __float2half_rn(vs);
            h[i] = vh;
            l[i] = __float2half_rn(vs - __half2float(vh));
        }
        *reinterpret_cast<uint4*>(hi + r * ldp + c0) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(lo + r * ldp + c0) = *reinterpret_cast<const uint4*>(l);
    }
    if (sumsq != nullptr) {
        const double total_sq = block_sum(acc, scratch);
        if (threadIdx.x == 0) atomicAdd(sumsq, total_sq);
        const double total_sum = block_sum(acc_sum, scratch);
        if (threadIdx.x == 0) atomicAdd(sumsq + 1, total_sum);   // sumsq[1] = plain sum
    }
}

int grid_for(int64_t work_items, int threads, int max_blocks_per_sm = 16) {
    const int64_t blocks = ceil_div64(work_items, threads);
    const int64_t cap = static_cast<int64_t>(sm_count()) * max_blocks_per_sm;
    return static_cast<int>(std::max<int64_t>(1, std::min(blocks, cap)));
}

}  // namespace

int absmax_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, unsigned int* out_bits, cudaStream_t stream) {
    if (rows * cols == 0) return 0;
    absmax_kernel<<<grid_for(rows * cols, 256, 8), 256, 0, stream>>>(x, rows, cols, ld, out_bits);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int rowabsmax_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, const float* col_mul,
                  unsigned int* rowmax_bits, cudaStream_t stream) {
    if (rows * cols == 0) return 0;
    const int bx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(cols, 256 * 4), 64)));
    rowabsmax_kernel<<<dim3(bx, static_cast<unsigned>(rows)), 256, 0, stream>>>(x, cols, ld, col_mul, rowmax_bits);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int split_matrix(const float* x, int64_t rows, int64_t cols, int64_t ld, const unsigned int* max_bits, __half* hi,
                 __half* lo, int64_t ldp, float* inv_scale, double* sumsq, int* neg_flag, cudaStream_t stream) {
    NFB_CHECK(ldp % 8 == 0 && ldp >= cols, "plane leading dimension must be a multiple of 8 and >= cols");
    if (rows == 0) return 0;
    split_kernel<false><<<grid_for(rows * (ldp / 8), 256, 8), 256, 0, stream>>>(x, rows, cols, ld, nullptr, max_bits, hi,
                                                                              lo, ldp, inv_scale, sumsq, neg_flag);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int split_factor(const float* f, int64_t k, int64_t cols, int64_t ld, const float* col_mul,
                 const unsigned int* rowmax_bits, __half* hi, __half* lo, int64_t ldp, float* inv_scale,
                 cudaStream_t stream) {
    NFB_CHECK(ldp % 8 == 0 && ldp >= cols, "plane leading dimension must be a multiple of 8 and >= cols");
    if (k == 0) return 0;
    split_kernel<true><<<grid_for(k * (ldp / 8), 256, 8), 256, 0, stream>>>(f, k, cols, ld, col_mul, rowmax_bits, hi, lo,
                                                                          ldp, inv_scale, nullptr, nullptr);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
