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

}  // namespace nfb
