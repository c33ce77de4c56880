// Matrix ingest: probtrackx fdt_matrix2 COO triplets -> dense float32 X, bit-exact with the reference.
//
// Reference arithmetic (NFACT/base/matrix_handling.py:116-158, NFACT/decomp/decomposition/
// matrix_handling.py:73-100), per matrix cell e and subject s, all in float64:
//     d_s[e]  = sum of the duplicate triplets of subject s that hit e      (csc_matrix constructor)
//     m_s[e]  = d_s[e] * (1e8 / waytotal_s)                                (normalise_matrix)
//     acc[e]  = ((m_0[e] + m_1[e]) + m_2[e]) + ...   in subject order      (avg_fdt, sparse +)
//     X[e]    = float32(acc[e] * (1.0 / S))  (group)  |  float32(m_0[e])  (single subject)
//
// Device plan per subject: phase A scatter-adds the raw values into a zeroed fp64 scratch `tmp`
// (fp64 atomics; only duplicates ever collide, and their sum is order-free for integer streamline
// counts -- non-integer `--pd` output has no duplicates), phase B lets the first thread that
// exchanges a non-zero scratch cell back to zero add cell*scale to `acc`.  One add per cell per subject,
// subjects processed in order on one stream => the float64 evaluation order of the reference.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

__global__ void scatter_add_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                                   const double* __restrict__ vals, int64_t nnz, int64_t row0, int64_t nrows_blk,
                                   int64_t nrows_total, int64_t ncols, double* __restrict__ tmp,
                                   int* __restrict__ oob_flag) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nnz;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = rows[i], c = cols[i];
        if (r < 0 || c < 0 || r >= nrows_total || c >= ncols) { *oob_flag = 1; continue; }
        if (r < row0 || r >= row0 + nrows_blk) continue;
        atomicAdd(&tmp[(r - row0) * ncols + c], vals[i]);
    }
}

__global__ void scatter_commit_kernel(const int* __restrict__ rows, const int* __restrict__ cols, int64_t nnz,
                                      int64_t row0, int64_t nrows_blk, int64_t ncols, double scale,
                                      double* __restrict__ tmp, double* __restrict__ acc) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nnz;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = rows[i], c = cols[i];
        if (r < row0 || r >= row0 + nrows_blk || c < 0 || c >= ncols) continue;
        const int64_t e = (r - row0) * ncols + c;
        const unsigned long long old = atomicExch(reinterpret_cast<unsigned long long*>(&tmp[e]), 0ULL);
        const double d = __longlong_as_double(static_cast<long long>(old));
        if (d != 0.0) acc[e] = __dadd_rn(acc[e], __dmul_rn(d, scale));  // sole owner of cell e in this pass
    }
}

__global__ void finalize_kernel(const double* __restrict__ acc, int64_t nelem, double recip, int apply_recip,
                                float* __restrict__ x) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nelem;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double v = acc[i];
        if (apply_recip) v = __dmul_rn(v, recip);
        x[i] = __double2float_rn(v);
    }
}

// ---- single subject straight into the operand planes ------------------------------------------------------
// A probtrackx matrix lists every (seed, target) pair once, so X[r][c] = float32(v * scale) needs neither the
// float64 scratch nor a dense float32 matrix: pass 1 finds max |X| (the planes' scale), pass 2 writes the fp16
// hi / lo halves of every triplet into the zeroed planes.  A one-bit-per-cell mask detects duplicate coordinates
// (the reference SUMS them in float64, scipy csc_matrix): any duplicate raises dup_flag and the caller redoes the
// subject on the exact scatter-add path.
__global__ void coo_absmax_kernel(const double* __restrict__ vals, int64_t nnz, double scale,
                                  unsigned int* __restrict__ out_bits) {
    float m = 0.0f;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nnz;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        m = fmaxf(m, fabsf(__double2float_rn(__dmul_rn(vals[i], scale))));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

__global__ void coo_to_planes_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                                     const double* __restrict__ vals, int64_t nnz, double scale, int64_t nrows,
                                     int64_t ncols, const unsigned int* __restrict__ max_bits,
                                     __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldp,
                                     unsigned int* __restrict__ cell_mask, float* __restrict__ inv_scale,
                                     double* __restrict__ sums, int* __restrict__ flags) {
    __shared__ double scratch[32];
    const float s = pow2_scale_for(__uint_as_float(max_bits[0]));
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = 1.0f / s;
    double acc = 0.0, acc_sum = 0.0;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nnz;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = rows[i], c = cols[i];
        if (r < 0 || c < 0 || r >= nrows || c >= ncols) { flags[0] = 1; continue; }      // out of bounds
        const int64_t cell = r * ldp + c;
        const unsigned int bit = 1u << (cell & 31);
        if (atomicOr(&cell_mask[cell >> 5], bit) & bit) { flags[1] = 1; continue; }       // duplicate coordinate
        const float v = __double2float_rn(__dmul_rn(vals[i], scale));
        acc += static_cast<double>(v) * static_cast<double>(v);
        acc_sum += static_cast<double>(v);
        if (v < 0.0f) flags[2] = 1;
        const float vs = v * s;
        const __half vh = __float2half_rn(vs);
        hi[cell] = vh;
        lo[cell] = __float2half_rn(vs - __half2float(vh));
    }
    const double total_sq = block_sum(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(sums, total_sq);
    const double total_sum = block_sum(acc_sum, scratch);
    if (threadIdx.x == 0) atomicAdd(sums + 1, total_sum);
}

int blocks_for(int64_t items) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(items, 256), sm_count() * 16LL)));
}

}  // namespace

int ingest_scatter(const int* rows, const int* cols, const double* vals, int64_t nnz, int64_t row0,
                   int64_t nrows_blk, int64_t nrows_total, int64_t ncols, double scale, double* tmp, double* acc,
                   int* oob_flag, cudaStream_t stream) {
    if (nnz == 0) return 0;
    scatter_add_kernel<<<blocks_for(nnz), 256, 0, stream>>>(rows, cols, vals, nnz, row0, nrows_blk, nrows_total,
                                                           ncols, tmp, oob_flag);
    NFB_CUDA(cudaGetLastError());
    if (acc == nullptr) return 0;   // single subject: the caller finalises tmp * scale directly
    scatter_commit_kernel<<<blocks_for(nnz), 256, 0, stream>>>(rows, cols, nnz, row0, nrows_blk, ncols, scale, tmp,
                                                              acc);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int ingest_finalize(const double* acc, int64_t nelem, double recip, int apply_recip, float* x, cudaStream_t stream) {
    if (nelem == 0) return 0;
    finalize_kernel<<<blocks_for(nelem), 256, 0, stream>>>(acc, nelem, recip, apply_recip, x);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int ingest_coo_to_planes(const int* rows, const int* cols, const double* vals, int64_t nnz, double scale,
                         int64_t nrows, int64_t ncols, unsigned int* max_bits, __half* hi, __half* lo, int64_t ldp,
                         unsigned int* cell_mask, float* inv_scale, double* sums, int* flags, cudaStream_t stream) {
    // callers zero max_bits, hi, lo, cell_mask, sums[2] and flags[3] on the same stream beforehand
    if (nnz > 0) {
        coo_absmax_kernel<<<blocks_for(nnz), 256, 0, stream>>>(vals, nnz, scale, max_bits);
        NFB_CUDA(cudaGetLastError());
    }
    // launched even for nnz == 0: it writes the planes' inverse scale
    coo_to_planes_kernel<<<blocks_for(std::max<int64_t>(nnz, 1)), 256, 0, stream>>>(
        rows, cols, vals, nnz, scale, nrows, ncols, max_bits, hi, lo, ldp, cell_mask, inv_scale, sums, flags);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
