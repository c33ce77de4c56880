// Factor-update kernels of the NMF loop.  Both factors are kept component-major on the device
// (W^T [k, m] and H [k, n], fp32) so the W and H half-steps run the same code -- the same symmetry
// sklearn uses when it calls _update_coordinate_descent(X, W, Ht) and then (X.T, Ht, W)
// (sklearn/decomposition/_nmf.py:492-501).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

constexpr float kEpsF32 = 1.1920928955078125e-07f;  // np.finfo(np.float32).eps, sklearn _nmf.py:32

__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, int64_t split_stride, int64_t cols,
                                       int64_t ld, float bias, float* __restrict__ out, int64_t ld_out) {
    const int64_t t = blockIdx.y;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < cols;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < splits; ++s) acc += part[s * split_stride + t * ld + j];
        out[t * ld_out + j] = acc + bias;
    }
}

// out[(j / slice_cols)][t][j % slice_cols] = sum_s part[s][t][j]   (slice-major layout for reduce-scatter)
__global__ void reduce_partials_sliced_kernel(const float* __restrict__ part, int splits, int64_t split_stride,
                                              int64_t k, int64_t cols, int64_t ld, float* __restrict__ out,
                                              int64_t slice_cols) {
    const int64_t t = blockIdx.y;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < cols;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < splits; ++s) acc += part[s * split_stride + t * ld + j];
        const int64_t slice = j / slice_cols;
        out[(slice * k + t) * slice_cols + (j - slice * slice_cols)] = acc;
    }
}

// sklearn/decomposition/_nmf.py:521-626 and :629-723 (beta_loss == 2, gamma == 1); four consecutive
// columns per thread (128-bit loads: every buffer has a leading dimension that is a multiple of 8).
__device__ __forceinline__ float mu_one(float old, float numerator, float denominator, float l1, float l2) {
    if (l1 > 0.0f) denominator = __fadd_rn(denominator, l1);
    if (l2 > 0.0f) denominator = __fadd_rn(denominator, __fmul_rn(l2, old));
    if (denominator == 0.0f) denominator = kEpsF32;
    return __fmul_rn(old, __fdiv_rn(numerator, denominator));
}

__global__ void mu_update_kernel(float* __restrict__ f, int64_t cols, int64_t ld, const float* __restrict__ num,
                                 int num_splits, int64_t num_stride, const float* __restrict__ den, int den_splits,
                                 int64_t den_stride, int64_t ld_part, float l1, float l2,
                                 unsigned int* __restrict__ rowmax_bits, const PeerMirror mirror) {
    const int64_t t = blockIdx.y;
    float vmax = 0.0f;
    const int64_t groups = (cols + 3) / 4;
    for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < groups;
         g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t j = g * 4;
        float4 numerator = make_float4(0.f, 0.f, 0.f, 0.f), denominator = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < num_splits; ++s) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(num + s * num_stride + t * ld_part + j));
            numerator.x += v.x; numerator.y += v.y; numerator.z += v.z; numerator.w += v.w;
        }
        for (int s = 0; s < den_splits; ++s) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(den + s * den_stride + t * ld_part + j));
            denominator.x += v.x; denominator.y += v.y; denominator.z += v.z; denominator.w += v.w;
        }
        float4 cur = *reinterpret_cast<const float4*>(f + t * ld + j);
        // columns >= cols are padding: partial buffers are not written there, keep the zeros
        cur.x = mu_one(cur.x, numerator.x, denominator.x, l1, l2);
        cur.y = (j + 1 < cols) ? mu_one(cur.y, numerator.y, denominator.y, l1, l2) : 0.0f;
        cur.z = (j + 2 < cols) ? mu_one(cur.z, numerator.z, denominator.z, l1, l2) : 0.0f;
        cur.w = (j + 3 < cols) ? mu_one(cur.w, numerator.w, denominator.w, l1, l2) : 0.0f;
        *reinterpret_cast<float4*>(f + t * ld + j) = cur;
        // row-sharded runs: the updated H columns go straight into every peer's copy (the all-gather, fused)
#pragma unroll
        for (int p = 0; p < kGemmMaxPeers - 1; ++p)
            if (p < mirror.n) *reinterpret_cast<float4*>(mirror.f[p] + t * ld + j) = cur;
        vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(cur.x), fabsf(cur.y))), fmaxf(fabsf(cur.z), fabsf(cur.w)));
    }
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0 && vmax > 0.0f) atomicMax(&rowmax_bits[t], __float_as_uint(vmax));
}

__global__ void dot_partials_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int splits,
                                    int64_t b_stride, int64_t ldb, int64_t cols, double* __restrict__ out) {
    __shared__ double scratch[32];
    const int64_t t = blockIdx.y;
    double acc = 0.0;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < cols;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float bsum = 0.0f;
        for (int s = 0; s < splits; ++s) bsum += b[s * b_stride + t * ldb + j];
        acc += static_cast<double>(a[t * lda + j]) * static_cast<double>(bsum);
    }
    const double total = block_sum(acc, scratch);
    if (threadIdx.x == 0 && total != 0.0) atomicAdd(out, total);
}

__global__ void dot_small_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t k, int64_t ld,
                                 double* __restrict__ out) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < k * k; i += blockDim.x) {
        const int64_t t = i / k, u = i - t * k;
        acc += static_cast<double>(a[t * ld + u]) * static_cast<double>(b[t * ld + u]);
    }
    const double total = block_sum(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(out, total);
}

template <typename Tin, typename Tout>
__global__ void transpose_kernel(const Tin* __restrict__ in, int64_t rows, int64_t cols, int64_t ld_in,
                                 Tout* __restrict__ out, int64_t ld_out, int* __restrict__ nonfinite) {
    __shared__ Tout tile[32][33];
    const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 32, r0 = static_cast<int64_t>(blockIdx.y) * 32;
    bool bad = false;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        Tin v = Tin(0);
        if (r < rows && c < cols) v = in[r * ld_in + c];
        bad |= !isfinite(v);
        tile[i][threadIdx.x] = static_cast<Tout>(v);
    }
    if (nonfinite != nullptr && bad) *nonfinite = 1;
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[c * ld_out + r] = tile[threadIdx.x][i];
    }
}

__global__ void reduce_partials_f64_kernel(const float* __restrict__ part, int splits, int64_t split_stride,
                                           int64_t cols, int64_t ld, double* __restrict__ out, int64_t ld_out) {
    const int64_t t = blockIdx.y;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < cols;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < splits; ++s) acc += static_cast<double>(part[s * split_stride + t * ld + j]);
        out[t * ld_out + j] = acc;
    }
}

// 4 x 4 block of the Gram matrix per CTA column slice, float64 products of the fp32 factor.
// partial[z] = F F' over column chunk z: 64 x 64 output tiles (lower triangle of tiles), 256 threads with a 4 x 4
// register tile each, operands converted to float64 on their way into shared memory ([column][row], row stride 65:
// the transposing store and the row reads are both conflict-free enough), 16 DFMA per 8 LDS.64 -- DFMA-bound, where the
// first version (4 x 4 outputs per BLOCK, operands straight from L2) moved 4.5 GB through L2 for a C3 Gram.
constexpr int kGramTile = 64, kGramCols = 32, kGramStride = kGramTile + 1;
__global__ void __launch_bounds__(256)
gram_f64_kernel(const float* __restrict__ f, int k, int64_t cols, int64_t ld, double* __restrict__ gram64,
                int64_t cols_per_block) {
    __shared__ double as[kGramCols * kGramStride], bs[kGramCols * kGramStride];
    const int ti = blockIdx.x, ui = blockIdx.y;
    if (ui > ti) return;                                   // lower triangle of tiles, mirrored below
    const int t0 = ti * kGramTile, u0 = ui * kGramTile;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t j_begin = static_cast<int64_t>(blockIdx.z) * cols_per_block;
    const int64_t j_end = min(cols, j_begin + cols_per_block);
    double acc[4][4] = {};
    const int lc = threadIdx.x & 31, lr = threadIdx.x >> 5;    // loader: column within the chunk, first row
    // the next chunk's operands travel (registers) while the current chunk is multiplied
    float na[8], nb[8];
    auto fetch = [&](int64_t j0) {
        const int64_t j = j0 + lc;
        const bool in_cols = j < j_end;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int t = t0 + lr + 8 * q, u = u0 + lr + 8 * q;
            na[q] = (in_cols && t < k) ? f[static_cast<int64_t>(t) * ld + j] : 0.0f;
            nb[q] = (in_cols && u < k) ? f[static_cast<int64_t>(u) * ld + j] : 0.0f;
        }
    };
    if (j_begin < j_end) fetch(j_begin);
    for (int64_t j0 = j_begin; j0 < j_end; j0 += kGramCols) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            as[lc * kGramStride + lr + 8 * q] = static_cast<double>(na[q]);
            bs[lc * kGramStride + lr + 8 * q] = static_cast<double>(nb[q]);
        }
        __syncthreads();
        if (j0 + kGramCols < j_end) fetch(j0 + kGramCols);
#pragma unroll 8
        for (int c = 0; c < kGramCols; ++c) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = as[c * kGramStride + 4 * ty + i];
                b[i] = bs[c * kGramStride + tx + 16 * i];    // interleaved columns: consecutive lanes, consecutive doubles
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int l = 0; l < 4; ++l) acc[i][l] = fma(a[i], b[l], acc[i][l]);
        }
        __syncthreads();
    }
    // every (tile, column chunk) owns its entries of partial[z][t][u], u <= t: no atomics, the chunks are summed in
    // order by gram_f64_reduce_kernel -> the Gram matrix (and with it the regression) is bit-reproducible
    double* part = gram64 + static_cast<int64_t>(blockIdx.z) * k * k;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int t = t0 + 4 * ty + i, u = u0 + tx + 16 * l;
            if (t < k && u < k && u <= t) part[static_cast<int64_t>(t) * k + u] = acc[i][l];
        }
}

// gram[t][u] = gram[u][t] = sum_z partial[z][t][u]  (u <= t), chunks in order
__global__ void gram_f64_reduce_kernel(const double* __restrict__ partial, int k, int chunks, double* __restrict__ gram64) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    const int t = idx / k, u = idx - t * k;
    if (u > t) return;
    double acc = 0.0;
    for (int z = 0; z < chunks; ++z) acc += partial[static_cast<int64_t>(z) * k * k + idx];
    gram64[static_cast<int64_t>(t) * k + u] = acc;
    gram64[static_cast<int64_t>(u) * k + t] = acc;
}

__global__ void f64_to_f32_rowmax_kernel(const double* __restrict__ in, int64_t cols, int64_t ld_in,
                                         float* __restrict__ out, int64_t ld_out,
                                         unsigned int* __restrict__ rowmax_bits) {
    const int64_t t = blockIdx.y;
    float vmax = 0.0f;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < cols;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float v = __double2float_rn(in[t * ld_in + j]);
        out[t * ld_out + j] = v;
        vmax = fmaxf(vmax, fabsf(v));
    }
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0 && vmax > 0.0f) atomicMax(&rowmax_bits[t], __float_as_uint(vmax));
}

// z[r][:] = (x[r][:] - mean_r) / ||x[r][:] - mean_r||_2   (fp64 statistics; a constant row becomes all zeros and
// sets dead[r] = 1).  One CTA per row.
__global__ void __launch_bounds__(256)
rows_standardize_kernel(const float* __restrict__ x, int64_t cols, int64_t ld, float* __restrict__ z,
                        int64_t ldz, int* __restrict__ dead) {
    __shared__ double scratch[32];
    __shared__ double stat[2];
    const int64_t r = blockIdx.x;
    const float* row = x + r * ld;
    double sum = 0.0;
    for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) sum += static_cast<double>(row[j]);
    sum = block_sum(sum, scratch);
    if (threadIdx.x == 0) stat[0] = sum / static_cast<double>(cols);
    __syncthreads();
    const double mean = stat[0];
    double ss = 0.0;
    for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) {
        const double d = static_cast<double>(row[j]) - mean;
        ss = fma(d, d, ss);
    }
    ss = block_sum(ss, scratch);
    if (threadIdx.x == 0) {
        stat[1] = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
        dead[r] = ss > 0.0 ? 0 : 1;
    }
    __syncthreads();
    const double inv = stat[1];
    for (int64_t j = threadIdx.x; j < cols; j += blockDim.x)
        z[r * ldz + j] = static_cast<float>((static_cast<double>(row[j]) - mean) * inv);
}


// |avg * N(0, 1)| into a factor [k, cols] (ld): counter-based generator (splitmix64 of seed and element index ->
// two uniforms -> Box-Muller), one normal pair per two elements.  The NMF-sso runs draw their random initialisation
// with random_state=None in the reference (nmfsso.py:133), so only the distribution is part of the contract.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__global__ void fill_abs_normal_kernel(float* __restrict__ out, int64_t k, int64_t cols, int64_t ld, float avg,
                                       unsigned long long seed) {
    const int64_t pairs_per_row = (cols + 1) / 2;
    const int64_t total = k * pairs_per_row;
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t t = idx / pairs_per_row, c = 2 * (idx - t * pairs_per_row);
        const unsigned long long bits = splitmix64(splitmix64(seed) ^ static_cast<unsigned long long>(idx));
        const float u1 = (static_cast<float>(static_cast<unsigned int>(bits >> 40)) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
        const float u2 = static_cast<float>(static_cast<unsigned int>(bits) >> 8) * (1.0f / 16777216.0f);           // [0, 1)
        const float rad = sqrtf(-2.0f * __logf(u1));
        float sn, cs;
        __sincosf(6.283185307179586f * u2, &sn, &cs);
        out[t * ld + c] = fabsf(avg * rad * cs);
        if (c + 1 < cols) out[t * ld + c + 1] = fabsf(avg * rad * sn);
    }
}

int col_blocks(int64_t cols, int per_block, int cap) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(cols, per_block), cap)));
}

}  // namespace

int reduce_partials(const float* part, int splits, int64_t split_stride, int64_t k, int64_t cols, int64_t ld,
                    float bias, float* out, int64_t ld_out, cudaStream_t stream) {
    if (k * cols == 0) return 0;
    reduce_partials_kernel<<<dim3(col_blocks(cols, 256, 1024), static_cast<unsigned>(k)), 256, 0, stream>>>(
        part, splits, split_stride, cols, ld, bias, out, ld_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int reduce_partials_sliced(const float* part, int splits, int64_t split_stride, int64_t k, int64_t cols, int64_t ld,
                           float* out, int64_t slice_cols, cudaStream_t stream) {
    if (k * cols == 0) return 0;
    reduce_partials_sliced_kernel<<<dim3(col_blocks(cols, 256, 1024), static_cast<unsigned>(k)), 256, 0, stream>>>(
        part, splits, split_stride, k, cols, ld, out, slice_cols);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int mu_update(float* f, int64_t k, int64_t cols, int64_t ld, const float* num, int num_splits, int64_t num_stride,
              const float* den, int den_splits, int64_t den_stride, int64_t ld_part, float l1, float l2,
              unsigned int* rowmax_bits, cudaStream_t stream, const PeerMirror* mirror) {
    if (k * cols == 0) return 0;
    PeerMirror none{};
    mu_update_kernel<<<dim3(col_blocks(cols, 256 * 4, 1024), static_cast<unsigned>(k)), 256, 0, stream>>>(
        f, cols, ld, num, num_splits, num_stride, den, den_splits, den_stride, ld_part, l1, l2, rowmax_bits,
        mirror ? *mirror : none);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int dot_partials(const float* a, int64_t lda, const float* b, int splits, int64_t b_stride, int64_t ldb, int64_t k,
                 int64_t cols, double* out, cudaStream_t stream) {
    if (k * cols == 0) return 0;
    dot_partials_kernel<<<dim3(col_blocks(cols, 256, 256), static_cast<unsigned>(k)), 256, 0, stream>>>(
        a, lda, b, splits, b_stride, ldb, cols, out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int reduce_partials_f64(const float* part, int splits, int64_t split_stride, int64_t k, int64_t cols, int64_t ld,
                        double* out, int64_t ld_out, cudaStream_t stream) {
    if (k * cols == 0) return 0;
    reduce_partials_f64_kernel<<<dim3(col_blocks(cols, 256, 1024), static_cast<unsigned>(k)), 256, 0, stream>>>(
        part, splits, split_stride, cols, ld, out, ld_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int gram_f64(const float* f, int64_t k, int64_t cols, int64_t ld, double* gram64, cudaStream_t stream) {
    if (k == 0) return 0;
    if (cols == 0) {
        NFB_CUDA(cudaMemsetAsync(gram64, 0, sizeof(double) * k * k, stream));
        return 0;
    }
    const int64_t kt = ceil_div64(k, kGramTile);
    const int64_t tiles = kt * (kt + 1) / 2;
    // enough column chunks to give every SM two blocks, chunks a multiple of the 32-column step
    int64_t jz = std::max<int64_t>(1, std::min<int64_t>(ceil_div64(2 * static_cast<int64_t>(sm_count()), tiles),
                                                        ceil_div64(cols, 4 * kGramCols)));
    const int64_t per_block = round_up64(ceil_div64(cols, jz), kGramCols);
    jz = ceil_div64(cols, per_block);
    double* partial = nullptr;                       // [jz][k][k], stream-ordered scratch
    NFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&partial), sizeof(double) * jz * k * k, stream));
    gram_f64_kernel<<<dim3(static_cast<unsigned>(kt), static_cast<unsigned>(kt), static_cast<unsigned>(jz)), 256, 0,
                      stream>>>(f, static_cast<int>(k), cols, ld, partial, per_block);
    gram_f64_reduce_kernel<<<static_cast<unsigned>(ceil_div64(k * k, 256)), 256, 0, stream>>>(
        partial, static_cast<int>(k), static_cast<int>(jz), gram64);
    const cudaError_t launch_err = cudaGetLastError();
    cudaFreeAsync(partial, stream);
    NFB_CUDA(launch_err);
    return 0;
}

int f64_to_f32_rowmax(const double* in, int64_t k, int64_t cols, int64_t ld_in, float* out, int64_t ld_out,
                      unsigned int* rowmax_bits, cudaStream_t stream) {
    if (k * cols == 0) return 0;
    f64_to_f32_rowmax_kernel<<<dim3(col_blocks(cols, 256, 1024), static_cast<unsigned>(k)), 256, 0, stream>>>(
        in, cols, ld_in, out, ld_out, rowmax_bits);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int rows_standardize(const float* x, int64_t rows, int64_t cols, int64_t ld, float* z, int64_t ldz, int* dead,
                     cudaStream_t stream) {
    if (rows * cols == 0) return 0;
    rows_standardize_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(x, cols, ld, z, ldz, dead);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int dot_small(const float* a, const float* b, int64_t k, int64_t ld, double* out, cudaStream_t stream) {
    if (k == 0) return 0;
    dot_small_kernel<<<1, 1024, 0, stream>>>(a, b, k, ld, out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int transpose_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in, float* out, int64_t ld_out,
                  cudaStream_t stream, int* nonfinite) {
    if (rows * cols == 0) return 0;
    dim3 grid(static_cast<unsigned>(ceil_div64(cols, 32)), static_cast<unsigned>(ceil_div64(rows, 32)));
    transpose_kernel<float, float><<<grid, dim3(32, 8), 0, stream>>>(in, rows, cols, ld_in, out, ld_out, nonfinite);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int transpose_f64_to_f32(const double* in, int64_t rows, int64_t cols, int64_t ld_in, float* out, int64_t ld_out,
                         cudaStream_t stream, int* nonfinite) {
    if (rows * cols == 0) return 0;
    dim3 grid(static_cast<unsigned>(ceil_div64(cols, 32)), static_cast<unsigned>(ceil_div64(rows, 32)));
    transpose_kernel<double, float><<<grid, dim3(32, 8), 0, stream>>>(in, rows, cols, ld_in, out, ld_out, nonfinite);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int transpose_f64(const double* in, int64_t rows, int64_t cols, int64_t ld_in, double* out, int64_t ld_out,
                  cudaStream_t stream) {
    if (rows * cols == 0) return 0;
    dim3 grid(static_cast<unsigned>(ceil_div64(cols, 32)), static_cast<unsigned>(ceil_div64(rows, 32)));
    transpose_kernel<double, double><<<grid, dim3(32, 8), 0, stream>>>(in, rows, cols, ld_in, out, ld_out, nullptr);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int fill_abs_normal(float* out, int64_t k, int64_t cols, int64_t ld, float avg, unsigned long long seed,
                    cudaStream_t stream) {
    if (k * cols == 0) return 0;
    const int64_t pairs = k * ((cols + 1) / 2);
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(pairs, 256), sm_count() * 8)));
    fill_abs_normal_kernel<<<grid, 256, 0, stream>>>(out, k, cols, ld, avg, seed);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
