// Internal host-side launch API of the nfact_b200 CUDA kernels (C++; the public C ABI is
// include/nfact_b200.h).  All pointers are device pointers unless stated otherwise.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfb {

// fp16 hi/lo planes of a pre-scaled fp32 matrix, row-major [rows, ld] (ld % 8 == 0, padding = 0)
struct GemmOperand {
    const __half* hi;
    const __half* lo;
    int64_t rows;
    int64_t ld;
};

// ---- gemm_split.cu --------------------------------------------------------------------------------
int gemm_default_cta_group();
// n_out: output columns per row of A (components); weighs the split-slab traffic against the wave efficiency
int gemm_pick_splits(int64_t m_total, int64_t k_total, int cta_group, int64_t n_out);
// out[s][t][r] (r contiguous, leading dimension ld_out)
//     = scale * col_scale[t] * row_scale[r] * sum_{K in split s} A[r,K] B[t,K]      (null scale vectors -> 1)
//   a_is_mn_major == false : A planes are [m_total rows, K cols]     (X      for X H^T)
//   a_is_mn_major == true  : A planes are [K rows, m_total cols]     (X read as X^T for W^T X)
// scatter (nullable, unsplit launches only): the fused reduce-scatter epilogue of the row-sharded W'X -- row r of
// the result (an H column) goes to  base[r / slice][(src * rows + t) * slice + r % slice]  (peer memory) instead of
// out[t][r]
constexpr int kGemmMaxPeers = 8;
struct GemmScatter {
    float* base[kGemmMaxPeers];
    int64_t slice;     // H columns per rank
    int rows;          // k: component rows per source slab of an inbox
    int src;           // this rank
};
// Workspace of the stream-K tail (gemm_split.cu): ws holds one raw partial tile per worker, cnt two counters per
// worker (zeroed once by the owner of the buffers; the kernel re-arms them itself).  One workspace serves one stream.
struct GemmStreamK {
    float* ws;
    unsigned int* cnt;
    size_t ws_floats;     // capacity of ws
    int workers;          // capacity of cnt / 2
    bool force;           // engage wherever the plan pays even when NFB_STREAMK is not set (the test hook)
};
// floats of workspace a launch with n_store output columns needs at most
size_t gemm_streamk_ws_floats(int n_store);
int gemm_streamk_max_workers();
bool gemm_streamk_enabled();     // NFB_STREAMK=1 switches the stream-K tail on (off by default: measured, no gain)
// an unsplit launch of this shape would run with the stream-K tail (else: whole tiles, or split-K slabs)
bool gemm_streamk_pays(int64_t m_total, int64_t k_total, int cta_group);
int gemm_split(const GemmOperand& a, bool a_is_mn_major, const GemmOperand& b, int64_t m_total, int64_t k_total,
               int n_umma, int n_store, float* out, int64_t ld_out, int splits, const float* col_scale,
               const float* row_scale, float scale, int cta_group, cudaStream_t stream,
               const GemmScatter* scatter = nullptr, const GemmStreamK* sk = nullptr);

// ---- planes.cu ------------------------------------------------------------------------------------
// max |x| over a [rows, cols] fp32 matrix (ld >= cols) -> *out_bits (float bits, must be zeroed first)
int absmax_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, unsigned int* out_bits, cudaStream_t stream);
// per-row max |x| -> rowmax_bits[rows] (must be zeroed first)
// col_mul (nullable, [cols]): the maxima are taken over x[r][c] * col_mul[c]
int rowabsmax_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, const float* col_mul,
                  unsigned int* rowmax_bits, cudaStream_t stream);
// X [rows, cols] fp32 -> hi/lo planes [rows, ldp] scaled by the power of two that maps *max_bits into
// [2^14, 2^15); inv_scale[0] receives 1/scale; sumsq (may be null) points at two doubles that
// accumulate sum x^2 and sum x in fp64.
int split_matrix(const float* x, int64_t rows, int64_t cols, int64_t ld, const unsigned int* max_bits, __half* hi,
                 __half* lo, int64_t ldp, float* inv_scale, double* sumsq, int* neg_flag, cudaStream_t stream);
// factor F [k, cols] fp32 -> planes [>=k rows, ldp] with one power-of-two scale per row
// (from rowmax_bits[k]); inv_scale[k] receives 1/scale_t.  Columns [cols, ldp) are zeroed.
// col_mul (nullable, [cols]) multiplies column c before the split (used to fold the K-side scales of
// the other operand into a Gram matrix).
int split_factor(const float* f, int64_t k, int64_t cols, int64_t ld, const float* col_mul,
                 const unsigned int* rowmax_bits, __half* hi, __half* lo, int64_t ldp, float* inv_scale,
                 cudaStream_t stream);

// the same block of a factor matrix in the memory of up to 7 peers (pointers already offset like the local one)
struct PeerMirror {
    float* f[7];
    int n;
};

// ---- update.cu ------------------------------------------------------------------------------------
// out[t][j] = sum_s part[s][t][j] + bias     (t < k, j < cols)
int reduce_partials(const float* part, int splits, int64_t split_stride, int64_t k, int64_t cols, int64_t ld,
                    float bias, float* out, int64_t ld_out, cudaStream_t stream);
// out[j / slice_cols][t][j % slice_cols] = sum_s part[s][t][j]  (contiguous per-rank chunks for reduce-scatter)
int reduce_partials_sliced(const float* part, int splits, int64_t split_stride, int64_t k, int64_t cols, int64_t ld,
                           float* out, int64_t slice_cols, cudaStream_t stream);
// multiplicative update of one factor (sklearn _multiplicative_update_w/_h, beta = 2):
//   F[t][j] *= (sum_s num[s][t][j]) / den,  den = sum_s dpart[s][t][j] + l1 + l2 F[t][j], den == 0 -> eps
// rowmax_bits[k] (zeroed by the caller) receives the new per-row maxima.
// mirror (nullable): the same block of F in the memory of the other ranks of a row-sharded run; every value the
// kernel writes is written there as well (the all-gather of the H slices, fused into the update).
int mu_update(float* f, int64_t k, int64_t cols, int64_t ld, const float* num, int num_splits, int64_t num_stride,
              const float* den, int den_splits, int64_t den_stride, int64_t ld_part, float l1, float l2,
              unsigned int* rowmax_bits, cudaStream_t stream, const PeerMirror* mirror = nullptr);
// one coordinate-descent sweep over all k components for every column j of F [k, cols]
// (sklearn _cdnmf_fast.pyx with F = W^T):  num = sum_s part[s] - l1, G = gram (+ l2 on the diagonal).
// violation (fp64, accumulated) and rowmax_bits as above.
// den [k][ld_part] = G F from the tensor-core denominator GEMM seeds the per-column gradient.
int cd_sweep(float* f, int64_t k, int64_t cols, int64_t ld, const float* num, int num_splits, int64_t num_stride,
             const float* den, int64_t ld_part, const float* gram, float l1, float l2, double* violation,
             unsigned int* rowmax_bits, cudaStream_t stream, const PeerMirror* mirror = nullptr);
// *out += sum_{t<k, j<cols} a[t][j] * (sum_s b[s][t][j])        (fp64 accumulation)
int dot_partials(const float* a, int64_t lda, const float* b, int splits, int64_t b_stride, int64_t ldb, int64_t k,
                 int64_t cols, double* out, cudaStream_t stream);
// out64[t][j] = sum_s part[s][t][j]  (fp64 accumulation of fp32 partials)
int reduce_partials_f64(const float* part, int splits, int64_t split_stride, int64_t k, int64_t cols, int64_t ld,
                        double* out, int64_t ld_out, cudaStream_t stream);
// gram64[t][u] = sum_j f[t][j] f[u][j] in float64 (k x k, ld = k) straight from the fp32 factor
int gram_f64(const float* f, int64_t k, int64_t cols, int64_t ld, double* gram64, cudaStream_t stream);
// out32[t][j] = (float) in64[t][j]; rowmax_bits as in mu_update
int f64_to_f32_rowmax(const double* in, int64_t k, int64_t cols, int64_t ld_in, float* out, int64_t ld_out,
                      unsigned int* rowmax_bits, cudaStream_t stream);
// z[r][:] = (x[r][:] - mean_r) / ||x[r][:] - mean_r||  (fp64 statistics); dead[r] = 1 for a constant row (z = 0)
int rows_standardize(const float* x, int64_t rows, int64_t cols, int64_t ld, float* z, int64_t ldz, int* dead,
                     cudaStream_t stream);
// out[t][c] = |avg * N(0, 1)|, t < k, c < cols (counter-based generator keyed by seed and element index)
int fill_abs_normal(float* out, int64_t k, int64_t cols, int64_t ld, float avg, unsigned long long seed,
                    cudaStream_t stream);
// *out += sum_{t,u<k} a[t][u] * b[t][u]
int dot_small(const float* a, const float* b, int64_t k, int64_t ld, double* out, cudaStream_t stream);
// transpose [rows, cols] (ld_in) -> [cols, rows] (ld_out); *nonfinite (nullable) is set to 1 when an input
// element is inf or NaN
int transpose_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in, float* out, int64_t ld_out,
                  cudaStream_t stream, int* nonfinite = nullptr);
int transpose_f64_to_f32(const double* in, int64_t rows, int64_t cols, int64_t ld_in, float* out, int64_t ld_out,
                         cudaStream_t stream, int* nonfinite = nullptr);
int transpose_f64(const double* in, int64_t rows, int64_t cols, int64_t ld_in, double* out, int64_t ld_out,
                  cudaStream_t stream);

// ---- postops.cu -----------------------------------------------------------------------------------
// Component maps, component-major [k, ld].  stats: device scratch of 4 * k * n_groups doubles.
// x[t][j] = 0 where x[t][j] < mean + z * std of { x[t][j'] : gid[j'] == gid[j] }  (gid == nullptr: one group);
// samples whose group id is outside [0, n_groups) are left untouched.
template <typename T>
int comp_threshold(T* x, int64_t k, int64_t len, int64_t ld, const int* gid, int n_groups, double z, double* stats,
                   cudaStream_t stream);
// out[t][:] = StandardScaler over the samples of component t (population std, constant components -> scale 1)
template <typename T>
int comp_standardize(const T* x, int64_t k, int64_t len, int64_t ld, T* out, int64_t ld_out, double* stats,
                     cudaStream_t stream);

// winner-takes-all map: out[j] = 1 + first argmax_t z[t][j] of the StandardScaler z-scores (0 below z_thr);
// per_component: the scaler standardises every component over its samples, else every sample over the components
template <typename T>
int comp_wta(const T* x, int64_t k, int64_t len, int64_t ld, int per_component, double z_thr, double* stats,
             long long* out, cudaStream_t stream);

// out[t] = per-component correlation of a[t][:] and b[t][:] as NFACT/stats/stats_component_loadings.py:199-231
// computes it (covariance / (len - 1) over the product of the population standard deviations)
template <typename T>
int comp_loadings(const T* a, const T* b, int64_t k, int64_t len, int64_t ld, double* out, cudaStream_t stream);

// ---- p2p.cu ---------------------------------------------------------------------------------------
// Row-sharded H half-step over NVLink peer memory (CUDA IPC mappings of every rank's arena and H master).  The bulk
// data moves inside the compute kernels (GemmScatter in the W'X epilogue, PeerMirror in the sweep / update); these
// are the small kernels around them.
constexpr int kP2pMaxWorld = 8;
constexpr int kP2pRsFlag = 0, kP2pAgFlag = 16, kP2pError = 32, kP2pGramFlag = 48;   // int offsets into the arena header
constexpr int64_t kP2pScalOffset = 256;                          // byte offset of double scal[16][2]
constexpr int64_t kP2pHeaderBytes = 512;
struct P2pPeers {
    void* arena[kP2pMaxWorld];
    void* hmaster[kP2pMaxWorld];
};
// scal_src[2] / rowmax_src[kp] (nullable) -> slot `rank` of every arena, then arena_p.flag[flag_offset + rank] =
// epoch on every rank p (system-scope release after a fence)
int p2p_post(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, cudaStream_t stream,
             const double* scal_src = nullptr, int64_t scal_offset = 0, const unsigned int* rowmax_src = nullptr,
             int64_t rowmax_offset = 0, int kp = 0);
// this rank's Gram (slot `rank` of its gram_in at byte offset gram_offset) -> the same slot in every peer's arena
int p2p_push_gram(const P2pPeers& peers, int world, int rank, int64_t gram_offset, int gram_elems, cudaStream_t stream);
// wait for flag row `flag_offset` of every rank to reach `epoch`; gram_out = sum_p gram_in[p]
int p2p_wait_sum_gram(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, int64_t gram_offset,
                      int gram_elems, float* gram_out, cudaStream_t stream);
// post + wait in one launch
int p2p_post_wait(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, cudaStream_t stream);
// all-gather round trip in one launch: scal_src[2] (nullable) and rowmax[kp] -> every peer, flag, wait for all,
// scal_out[i] = sum_p scal[p][i] (nullable), rowmax[t] = max_p rowmax_in[p][t]
int p2p_post_gather(const P2pPeers& peers, int world, int rank, int epoch, const double* scal_src, int64_t scal_offset,
                    double* scal_out, unsigned int* rowmax, int64_t rowmax_offset, int kp, cudaStream_t stream);
// wait only
int p2p_wait(const P2pPeers& peers, int world, int rank, int flag_offset, int epoch, cudaStream_t stream);
// wait for every rank's slice of `epoch`; scal_out[i] (nullable) = sum_p scal[p][i]; rowmax_out[t] (nullable) =
// max_p rowmax_in[p][t]
int p2p_gather_wait(const P2pPeers& peers, int world, int rank, int epoch, int64_t scal_offset, double* scal_out,
                    int64_t rowmax_offset, int kp, unsigned int* rowmax_out, cudaStream_t stream);

// ---- ingest.cu ------------------------------------------------------------------------------------
// phase A: tmp[r*ncols+c] += v for every triplet with row in [row0, row0+nrows_blk)  (fp64 atomics)
// phase B: the first thread to exchange a non-zero tmp cell adds cell*scale to acc and leaves tmp zero
//          (skipped when acc == nullptr: single-subject ingest finalises tmp * scale itself).
int ingest_scatter(const int* rows, const int* cols, const double* vals, int64_t nnz, int64_t row0,
                   int64_t nrows_blk, int64_t nrows_total, int64_t ncols, double scale, double* tmp, double* acc,
                   int* oob_flag, cudaStream_t stream);
// X[r][c] = (float)(acc * recip)   (recip == 1.0 -> plain rounding)
int ingest_finalize(const double* acc, int64_t nelem, double recip, int apply_recip, float* x, cudaStream_t stream);

// One subject without duplicate coordinates straight into zeroed operand planes: X[r][c] = float32(v * scale)
// split as planes.cu does.  max_bits[1], cell_mask[ceil(nrows * ldp / 32)], sums[2] (sum x^2, sum x) and
// flags[3] (out of bounds, duplicate coordinate seen -> result invalid, negative value) must be zeroed before.
int ingest_coo_to_planes(const int* rows, const int* cols, const double* vals, int64_t nnz, double scale,
                         int64_t nrows, int64_t ncols, unsigned int* max_bits, __half* hi, __half* lo, int64_t ldp,
                         unsigned int* cell_mask, float* inv_scale, double* sums, int* flags, cudaStream_t stream);

// ---- dot_parse_gpu.cu -----------------------------------------------------------------------------
// fdt_matrix2.dot text (device buffer) -> COO triplets.  count: lines per 128-byte segment (seg_lines[segments]),
// per block (block_lines / block_first[blocks], exclusive scan) and in total; parse: rows0 / cols0 / vals of the
// first n_lines - 1 lines, shape[3] from the last, flags[3] = {malformed line, number outside the exact fast path,
// index not a 32-bit integer}, range[4] = {min row, max row, min col, max col}.
int64_t dot_gpu_segments(int64_t len);
int64_t dot_gpu_blocks(int64_t len);
int dot_gpu_count(const char* txt, int64_t len, unsigned char* seg_lines, unsigned int* block_lines,
                  long long* block_first, long long* total_out, cudaStream_t stream);
int dot_gpu_parse(const char* txt, int64_t len, const unsigned char* seg_lines, const long long* block_first,
                  long long n_lines, int* rows0, int* cols0, double* vals, double* shape, int* flags, int* range,
                  cudaStream_t stream);

// ---- nnls.cu --------------------------------------------------------------------------------------
// Batched non-negative least squares with one shared Gram matrix:
//   for every column j < cols:  min_{h >= 0} 0.5 h'Gh - rhs[:, j]'h ,  G [k,k] fp64, rhs [k, ld] fp32 partials
// sol [k, ld_sol] fp64.  status[0] counts columns that hit the iteration cap.
int nnls_gram_batched(const double* gram, int64_t k, const float* rhs, int rhs_splits, int64_t rhs_stride,
                      int64_t ld_rhs, int64_t cols, double* sol, int64_t ld_sol, int max_outer, int* status,
                      cudaStream_t stream, unsigned long long* steps_total = nullptr,
                      unsigned long long* counter = nullptr);   // counter: device scratch for dynamic scheduling

}  // namespace nfb
