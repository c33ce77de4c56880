/* nfact_b200 -- C ABI of the B200-native NFACT decomposition hot path.
 *
 * libnfact_b200.so is the drop-in boundary: plain pointers and sizes, no torch / numpy types.
 * NFACT itself is pure Python and has no FFI layer; the three seams below are the Python call
 * sites where NFACT hands NumPy arrays to third-party numerics.  Each entry
 * point cites the reference interface it replaces; INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success; otherwise nfb_last_error() describes the failure
 *     (thread-local, valid until the next failing call on that thread)
 *   - "host" pointers are ordinary (pageable or pinned) host memory; "dev" pointers are CUDA device
 *     pointers on the handle's device
 *   - matrices are C-order (row-major); W is [m, k], H is [k, n] exactly like sklearn / NFACT
 *   - all work is issued on the handle's stream; *_host entry points return after the results
 *     have landed in the host buffers
 */
#ifndef NFACT_B200_H
#define NFACT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nfb_handle nfb_handle;   /* device + stream + scratch (+ optional NCCL communicator) */
typedef struct nfb_matrix nfb_matrix;   /* a connectivity matrix X resident in HBM as fp16 hi/lo planes */
typedef struct nfb_dr_components nfb_dr_components;   /* group components prepared for the dual regression */
typedef struct nfb_nmf nfb_nmf;         /* NMF solver state for one (X, k): factors, Grams, workspaces */

/* ---- library / handle ------------------------------------------------------------------------ */
const char* nfb_last_error(void);
const char* nfb_version(void);
/* `stream` is a cudaStream_t (NULL: the library creates its own non-blocking stream). */
int nfb_create(int device, void* stream, nfb_handle** out);
int nfb_destroy(nfb_handle* h);
int nfb_synchronize(nfb_handle* h);
/* frees the per-subject buffers (operand planes, triplet staging) a dual-regression subject loop left cached on
 * this handle for the next subject, and trims the stream-ordered pool the library allocates from */
int nfb_release_cached(nfb_handle* h);
/* Row-sharded multi-GPU: attach an initialised ncclComm_t (one rank per GPU).  When set, the H half-step
 * all-reduces the k x n numerator and the k x k Gram W'W across ranks; W shards stay local.
 * `nccl_allreduce_fn` is the address of ncclAllReduce in the NCCL library that created the communicator. */
int nfb_set_comm(nfb_handle* h, void* nccl_comm, void* nccl_allreduce_fn, int rank, int world_size);
/* helpers so the host language can create the communicator without linking NCCL itself */
int nfb_nccl_unique_id(const char* libnccl_path, void* id_out_128_bytes);
int nfb_nccl_init_rank(nfb_handle* h, const char* libnccl_path, const void* id_128_bytes, int rank, int world_size);

/* ---- ingest -------------------------------------------------------------------------------------
 * replaces  NFACT/base/matrix_handling.py:116-158  load_single_matrix / load_fdt_matrix
 *           NFACT/decomp/decomposition/matrix_handling.py:73-100  avg_fdt
 * rows/cols are 0-based (the caller subtracts 1 exactly like load_single_matrix does), vals are the
 * float64 third column, scale[s] = 1e8 / waytotal_s computed by the caller in float64.
 * group_mode == 0: exactly one subject, X = float32(d * scale).
 * group_mode != 0: X = float32((m_0 + m_1 + ...) * (1.0 / n_subjects)), float64 throughout.
 * The result is bit-identical to the reference (see ingest.cu for the duplicate-coordinate caveat).
 * x_out_host: float32 [nrows, ncols] (may be NULL); x_out_dev: same on the device (may be NULL).   */
int nfb_ingest_coo(nfb_handle* h, int n_subjects, const int32_t* const* rows_host, const int32_t* const* cols_host,
                   const double* const* vals_host, const int64_t* nnz, const double* scale, int64_t nrows,
                   int64_t ncols, int group_mode, float* x_out_host, float* x_out_dev);

/* Same ingest, but the result stays on the device as a resident nfb_matrix (no 4mn-byte round trip through the
 * host): what the per-subject dual-regression pipeline (NFACT/dual_reg/dual_regression/run_dual_regression.py:
 * 150-164, load_fdt_matrix -> run_decomp) needs. */
int nfb_ingest_coo_matrix(nfb_handle* h, int n_subjects, const int32_t* const* rows_host,
                          const int32_t* const* cols_host, const double* const* vals_host, const int64_t* nnz,
                          const double* scale, int64_t nrows, int64_t ncols, int group_mode, nfb_matrix** out);

/* The whole single-subject load on the device: the bytes of fdt_matrix2.dot (decompressed) are uploaded, parsed by
 * the GPU (what np.loadtxt + the index arithmetic of NFACT/base/matrix_handling.py:93-137 do), scaled by
 * `scale` = 1e8 / waytotal and written straight into the resident operand planes; shape_out[2] = (nrows, ncols)
 * from the file's last line.  Numbers are converted exactly (one correctly rounded IEEE operation) when they have
 * at most 19 significant digits, a mantissa <= 2^53 and a decimal exponent within +-22 -- probtrackx's integer
 * counts and %g output.  *fallback_out != 0 and *out == NULL when the file needs the host path instead: 1 = a
 * number outside that fast path (or inf / nan), 2 = duplicate coordinates (summed in float64 by the reference);
 * the caller then uses nfb_dot_parse_coo + nfb_ingest_coo_matrix.  Malformed lines and index errors are errors
 * with the same messages as nfb_dot_parse_coo. */
int nfb_ingest_dot_matrix(nfb_handle* h, const char* text_host, int64_t len, double scale, nfb_matrix** out,
                          int64_t* shape_out, int* fallback_out);

/* The same device parse, but the triplets come back to the host (what the group average, which adds subjects in
 * float64 from host triplets, needs): 0-based int32 rows / cols and float64 values of all lines but the last in
 * page-locked blocks from nfb_host_alloc (the caller releases them with nfb_host_free), their number in *nnz_out,
 * the last line in shape_out[2].  *fallback_out = 1 (and NULL arrays) when the file needs nfb_dot_parse_coo. */
int nfb_dot_parse_coo_gpu(nfb_handle* h, const char* text_host, int64_t len, int32_t** rows0_out, int32_t** cols0_out,
                          double** vals_out, int64_t* nnz_out, int64_t* shape_out, int* fallback_out);

/* Text parser for fdt_matrix2.dot (replaces np.loadtxt at NFACT/base/matrix_handling.py:93-113).
 * `buf` is the (decompressed) file content.  Two calls: count the non-empty lines, then fill three
 * float64 arrays of that length with the row / column / value columns (the last line is the matrix
 * shape, exactly as in the file).  Parsing is correctly rounded (std::from_chars) and multi-threaded;
 * nthreads <= 0 uses all host cores.  A malformed line is an error, like loadtxt's ValueError. */
int nfb_dot_count_lines(const char* buf, int64_t len, int nthreads, int64_t* n_lines_out);
int nfb_dot_parse(const char* buf, int64_t len, int nthreads, int64_t n_lines, double* rows, double* cols,
                  double* vals);
/* The same parse fused with what load_single_matrix does next (NFACT/base/matrix_handling.py:131-137): the
 * first n_lines - 1 lines become 0-based int32 rows0 / cols0 and float64 vals -- exactly the arrays
 * nfb_ingest_coo takes, so they can live in page-locked memory (nfb_host_alloc) -- and the last line becomes
 * shape_out[2] = (nrows, ncols).  Index errors carry scipy's messages ("row index exceeds matrix dimensions" ...). */
int nfb_dot_parse_coo(const char* buf, int64_t len, int nthreads, int64_t n_lines, int32_t* rows0, int32_t* cols0,
                      double* vals, int64_t* shape_out);

/* LZ4 frame decoder for fdt_matrix2.dot.lz4 (replaces lz4.frame.open(matfile, "rb").read() at
 * NFACT/base/matrix_handling.py:108-110; files written by NFACT/config/nfact_config_functions.py:601-604).
 * Two calls: the decoded size of the whole file (all concatenated frames; skippable frames are skipped), then the
 * decode into a caller-owned buffer of at least that size.  Header, block and content checksums (XXH32) are
 * verified; a corrupt or truncated file is an error carrying LZ4F's error name.  Frames with independent blocks
 * are decoded by `nthreads` host threads (<= 0: all cores), linked-block frames in order. */
int nfb_lz4_frame_size(const void* src, int64_t src_len, int nthreads, int64_t* size_out);
int nfb_lz4_frame_decompress(const void* src, int64_t src_len, void* dst, int64_t dst_cap, int nthreads,
                             int64_t* written_out);
/* Diagnostics of the last nfb_lz4_frame_decompress call on the calling thread: how many chunks of linked-block frames
 * were decoded side by side (beyond the first chunk of each frame), and how many of them had to wait for their
 * predecessor because the bytes in front of them could not be reconstructed independently. */
int nfb_lz4_last_chunks(int64_t* parallel_chunks_out, int64_t* deferred_out);

/* ---- resident matrix --------------------------------------------------------------------------
 * X float32 [m, n] -> fp16 hi/lo planes in HBM (4 B/element, read by both X H' and W'X). */
int nfb_matrix_from_host(nfb_handle* h, const float* x_host, int64_t m, int64_t n, nfb_matrix** out);
int nfb_matrix_from_device(nfb_handle* h, const float* x_dev, int64_t m, int64_t n, int64_t ld, nfb_matrix** out);
int nfb_matrix_free(nfb_matrix* x);
int nfb_matrix_shape(const nfb_matrix* x, int64_t* m, int64_t* n);
/* ||X||_F^2 accumulated in float64 from the original float32 values */
int nfb_matrix_sumsq(const nfb_matrix* x, double* out);
/* sum of all elements in float64 (X.mean() * m * n; needed by sklearn's init='random' / nndsvda scaling) */
int nfb_matrix_sum(const nfb_matrix* x, double* out);
/* 1 when any element of X is negative (sklearn's check_non_negative, evaluated during the split pass) */
int nfb_matrix_has_negative(const nfb_matrix* x, int* out);

/* Generic product of the resident matrix with a dense float32 factor given component-major:
 *   trans == 0 : out[t*ld_out + i] = sum_j X[i,j] * b_dev[t*n + j]     (b_dev [n_b, n], out [n_b, m])
 *   trans != 0 : out[t*ld_out + j] = sum_i X[i,j] * b_dev[t*m + i]     (b_dev [n_b, m], out [n_b, n])
 * Used by the NNDSVD initialisation (randomized range finder, sklearn/utils/extmath.py:313-383). */
int nfb_matrix_matmul(nfb_handle* h, const nfb_matrix* x, int trans, const float* b_dev, int64_t n_b,
                      float* out_dev, int64_t ld_out);

/* ---- NMF solve ----------------------------------------------------------------------------------
 * replaces  NFACT/decomp/decomposition/sso/nmfsso.py:32-59  nmf_decomp  ->  sklearn NMF.fit_transform
 *           (sklearn/decomposition/_nmf.py:399-518 coordinate descent, :726-897 multiplicative update,
 *            beta_loss = 'frobenius'), from given initial factors. */
typedef struct nfb_nmf_params {
    int solver;        /* 0 = "cd" (sklearn default, NFACT default), 1 = "mu" */
    int max_iter;      /* sklearn max_iter */
    double tol;        /* sklearn tol */
    double l1_reg_W;   /* sklearn _compute_regularization (_nmf.py:1249-1259), computed by the caller */
    double l1_reg_H;
    double l2_reg_W;
    double l2_reg_H;
    int verbose;
} nfb_nmf_params;

/* One call, host buffers in and out (the reference-facing entry point):
 *   w_host [m, k] float32 in/out (initial guess -> solution), h_host [k, n] float32 in/out.
 *   n_iter_out, recon_err_out (||X - WH||_F, sklearn reconstruction_err_) may be NULL. */
int nfb_nmf_fit_host(nfb_handle* h, const nfb_matrix* x, int k, const nfb_nmf_params* p, float* w_host,
                     float* h_host, int* n_iter_out, double* recon_err_out);

/* Stepwise interface (inputs resident; used by the benchmark and by per-iteration parity tests). */
int nfb_nmf_create(nfb_handle* h, const nfb_matrix* x, int k, const nfb_nmf_params* p, nfb_nmf** out);
int nfb_nmf_free(nfb_nmf* s);
int nfb_nmf_set_factors_host(nfb_nmf* s, const float* w_host /* [m,k] */, const float* h_host /* [k,n] */);
int nfb_nmf_set_factors_dev(nfb_nmf* s, const float* w_dev /* [m,k] */, const float* h_dev /* [k,n] */);
int nfb_nmf_get_factors_host(nfb_nmf* s, float* w_host, float* h_host);
int nfb_nmf_get_factors_dev(nfb_nmf* s, float* w_dev, float* h_dev);
/* Run `n_iters` full iterations (W half-step then H half-step) without convergence checks.
 * violation_out (CD only, may be NULL) receives the projected-gradient violation of the last iteration. */
int nfb_nmf_iterate(nfb_nmf* s, int n_iters, double* violation_out);
/* Full sklearn loop with its stopping rule from the current factors. */
int nfb_nmf_run(nfb_nmf* s, int* n_iter_out, double* recon_err_out);
/* ||X - WH||_F for the current factors (one extra X H' pass). */
int nfb_nmf_error(nfb_nmf* s, double* err_out);
/* How the row-sharded H half-step of this solver exchanges data: *out = 0 single GPU (no exchange),
 * 1 NCCL reduce-scatter / all-gather, 2 NVLink peer-memory kernels (p2p.cu). */
int nfb_nmf_exchange_mode(const nfb_nmf* s, int* out);
/* number of kernels the last iterate/run call launched (for the benchmark's gpu_launches) */
int64_t nfb_launch_count(const nfb_handle* h);
/* Per-kernel timing of the two X-streaming GEMMs (X H' and W'X) with CUDA events on the launch stream.
 * Returns the accumulated duration / count since the previous call, then sets the enable flag. */
int nfb_profile_gemms(nfb_handle* h, int enable, double* total_ms_out, int64_t* launches_out);

/* ---- non-negative dual regression ---------------------------------------------------------------
 * replaces  NFACT/dual_reg/dual_regression/dual_regression_methods.py:92-167  nmf_dual_regression
 *           (scipy.optimize.nnls per column / per row).
 *   g_host   [m, k] group grey-matter components, float64 (NIfTI get_fdata) or float32 (GIFTI / CIFTI
 *            loaders, NFACT/dual_reg/nfact_dr_functions.py:296-353) as said by g_is_f32
 *   gs_out   [m, k] float64  subject grey components      (stage 2)
 *   hs_out   [k, n] float64  subject white components     (stage 1, already transposed like the reference)
 *   n_unconverged_out: number of columns/rows whose solver hit its sweep cap (scipy would raise)
 *   g_nonfinite_out:   1 when g_host holds an inf / NaN (scipy raises ValueError; the outputs are then garbage) */
int nfb_dual_regression_host(nfb_handle* h, const nfb_matrix* x, int k, const void* g_host, int g_is_f32,
                             double* gs_out, double* hs_out, int* n_unconverged_out, int* g_nonfinite_out);

/* The same regression followed, while the results are still in HBM, by the per-subject post-processing of
 * NFACT/dual_reg/dual_regression/run_dual_regression.py:166-186:
 *   zscore != 0: thresholding_components (NFACT/base/matrix_handling.py:241-277) -- white components: entries
 *     below mean + zscore * std of their component -> 0 (:218-238); grey components: the same per (component,
 *     seed), seed_id_host[m] = the seed index of every row (second-to-last column of coords_for_fdt_matrix2,
 *     :184-215), rows with an index outside [0, n_seeds) are left alone;
 *   gs_norm_out / hs_norm_out != NULL: normalise_components (:45-68), StandardScaler per component of the
 *     (thresholded) maps, written to these extra [m, k] / [k, n] float64 arrays. */
int nfb_dual_regression_post_host(nfb_handle* h, const nfb_matrix* x, int k, const void* g_host, int g_is_f32,
                                  double zscore, const int32_t* seed_id_host, int n_seeds, double* gs_out,
                                  double* hs_out, double* gs_norm_out, double* hs_norm_out, int* n_unconverged_out,
                                  int* g_nonfinite_out);

/* The subject loop of nfact_dr (NFACT/dual_reg/dual_regression/set_up_dual_regression.py:186-235) regresses every
 * subject onto the SAME group components: nfb_dr_components_create uploads them once (float32 planes for the
 * G'X product, float64 Gram matrix G'G; g_nonfinite_out as above) and nfb_dual_regression_resident_host is
 * nfb_dual_regression_post_host on the prepared object, so a subject costs no component upload.  The object
 * belongs to the handle it was created on. */
int nfb_dr_components_create(nfb_handle* h, const void* g_host, int g_is_f32, int64_t m, int k,
                             nfb_dr_components** out, int* g_nonfinite_out);
int nfb_dr_components_free(nfb_dr_components* c);
int nfb_dual_regression_resident_host(nfb_handle* h, const nfb_matrix* x, nfb_dr_components* comps, double zscore,
                                      const int32_t* seed_id_host, int n_seeds, double* gs_out, double* hs_out,
                                      double* gs_norm_out, double* hs_norm_out, int* n_unconverged_out);

/* Post-processing of component maps given as host arrays (replaces the numpy / sklearn arithmetic of
 * thresholding :218-238, threshold_grey_components :184-215 and normalise_components :45-68 in
 * NFACT/base/matrix_handling.py).  data: float32 or float64 (is_f64), [k, len] when comp_major (white components)
 * or [len, k] (grey components).  do_threshold: entries below mean + zscore * std of their (component, group)
 * are zeroed IN PLACE; group_id[len] (NULL: one group) assigns every sample to a group in [0, n_groups), samples
 * outside stay untouched.  norm_out (NULL: skip): same shape and dtype, receives the z-scored (StandardScaler)
 * copy of the -- thresholded -- data.  Statistics are float64. */
int nfb_component_postops_host(nfb_handle* h, void* data, int is_f64, int comp_major, int64_t k, int64_t len,
                               const int32_t* group_id, int n_groups, int do_threshold, double zscore,
                               void* norm_out);

/* Winner-takes-all map (replaces create_wta_map, NFACT/decomp/pipes/image_handling.py:216-245: StandardScaler
 * z-scores of every column of the host array, then 1 + the index of the first maximum over the components, 0 where
 * that maximum is below z_thr, compared in the array's dtype).  data: float32 or float64 (is_f64);
 *   comp_major = 1: [k, len] (white components, the reference's axis = 0): every sample is standardised over its k
 *                   component values;
 *   comp_major = 0: [len, k] (grey components, axis = 1): every component is standardised over its len samples.
 * wta_out[len]: int64 like numpy's int. */
int nfb_component_wta_host(nfb_handle* h, const void* data, int is_f64, int comp_major, int64_t k, int64_t len,
                           double z_thr, int64_t* wta_out);

/* Component loadings (replaces __correlating at NFACT/stats/stats_component_loadings.py:199-231): for every
 * component t the correlation between the subject's map and the group's map,
 *   loadings_out[t] = [sum_j (s_tj - mean s_t)(g_tj - mean g_t) / (len - 1)] / (std s_t * std g_t)
 * with population standard deviations, exactly the reference's (mixed) normalisation; a constant map gives NaN
 * like numpy's 0 / 0.  subject / group: host arrays of the same dtype (is_f64) and layout, [k, len] when
 * comp_major else [len, k] (the reference's voxels x components). */
int nfb_component_loadings_host(nfb_handle* h, const void* subject, const void* group, int is_f64, int comp_major,
                                int64_t k, int64_t len, double* loadings_out);

/* ---- NMF-sso similarity --------------------------------------------------------------------------------
 * replaces  NFACT/decomp/decomposition/sso/sso_functions.py:29-68  compute_similairty_matrix:
 * np.corrcoef of the (row-normalised) stacked white-matter components of all NMF-sso runs -- (N k)^2 n
 * multiply-adds, 3.5 TFLOP at C3 -- as one pass of the split-precision GEMM over the standardised rows.
 *   rows_host [r, n] float32;  sim_out_host [r, r] float32 Pearson correlations (unclipped);
 *   dead_out_host [r] (nullable): 1 for a constant row (np.corrcoef gives NaN there; here its row/column is 0). */
int nfb_row_correlation_host(nfb_handle* h, const float* rows_host, int64_t r, int64_t n, float* sim_out_host,
                             int32_t* dead_out_host);

/* ---- NMF-sso with the runs' components resident in HBM ---------------------------------------------------
 * replaces the data movement of  NFACT/decomp/decomposition/sso/nmfsso.py:120-190 (NMFsso.run: every run's W, H
 * returned to the parent), :137 (per-run thresholding), sso_functions.py:29-68 (similarity of the stack) and
 * nmfsso.py:346-357 (centrotype rows as the start of the final solve).  The stacks [runs k, n] / [runs k, m] stay
 * on the device; only the [runs k, runs k] similarity matrix and the final factors travel.
 *   nfb_nmf_init_random : |avg N(0,1)| into both factors of a solver state (the reference draws the runs with
 *                         random_state=None, so only the distribution is part of the contract)
 *   nfb_sso_add_run     : factors of a finished solve -> slot `run` of the stacks, white copy z-thresholded
 *   nfb_sso_similarity_host : Pearson correlations of the stacked white components, dead rows flagged
 *   nfb_sso_seed_final  : gathers the centrotype rows into the (k = n_centroids) state of the final solve
 *   nfb_nmf_factors_all_zero : sklearn's "full of zeros" check of a custom start, on the resident factors */
typedef struct nfb_sso nfb_sso;
int nfb_sso_create(nfb_handle* h, int64_t m, int64_t n, int k, int runs, nfb_sso** out);
int nfb_sso_free(nfb_sso* sso);
int nfb_nmf_init_random(nfb_nmf* s, double avg, uint64_t seed);
int nfb_sso_add_run(nfb_sso* sso, nfb_nmf* s, int run, double zscore);
int nfb_sso_similarity_host(nfb_sso* sso, float* sim_out_host, int32_t* dead_out_host);
int nfb_sso_seed_final(nfb_sso* sso, nfb_nmf* s, const int32_t* centroids_host, int n_centroids);
int nfb_nmf_factors_all_zero(nfb_nmf* s, int32_t* flags_host);

/* ---- page-locked host buffers ------------------------------------------------------------------------
 * Result arrays are hundreds of MB; device-to-host copies into pageable memory run at a few GB/s, into
 * page-locked memory at PCIe speed.  The host layer allocates the arrays it returns (W, H, DR outputs)
 * here; blocks are recycled by size.  Any *_host entry point also accepts ordinary pageable pointers. */
int nfb_host_alloc(int64_t bytes, void** out);
int nfb_host_free(void* p);

/* ---- low-level kernels exported for tests / profiling ------------------------------------------- */
/* out_dev[t*ld_out + r] = sum_K A[r,K] B[t,K] through the split-precision tcgen05 path.
 * a_dev [a_rows, a_cols] float32 (a_is_transposed: A is given as [K, M]); b_dev [n_b, K] float32. */
int nfb_test_gemm(nfb_handle* h, const float* a_dev, int64_t a_rows, int64_t a_cols, int a_is_transposed,
                  const float* b_dev, int64_t n_b, float* out_dev, int64_t ld_out, int cta_group, int splits);

#ifdef __cplusplus
}
#endif
#endif /* NFACT_B200_H */
