// Post-processing of component maps on the device:
//   thresholding               NFACT/base/matrix_handling.py:218-238   x[x < mean + z std] = 0 per component
//   threshold_grey_components  NFACT/base/matrix_handling.py:184-215   the same per (component, seed)
//   normalise_components       NFACT/base/matrix_handling.py:45-68     StandardScaler().fit_transform per component
//   create_wta_map             NFACT/decomp/pipes/image_handling.py:216-245   winner-takes-all map of z-scored maps
// All maps are handled component-major, [k, ld] with the samples of one component contiguous (the layout the
// solvers keep W' and H in).  Statistics are float64 two-pass sums with a fixed reduction tree (deterministic):
// population mean / standard deviation like ndarray.mean() / .std() and sklearn's StandardScaler.
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

constexpr int kStatThreads = 512;

// one CTA per (component t, group g): mean, population std and count of { x[t][j] : gid[j] == g }
template <typename T>
__global__ void __launch_bounds__(kStatThreads)
comp_stats_kernel(const T* __restrict__ x, int64_t len, int64_t ld, const int* __restrict__ gid, int n_groups,
                  double* __restrict__ mean_out, double* __restrict__ std_out, double* __restrict__ var_out,
                  long long* __restrict__ count_out) {
    __shared__ double scratch[32];
    __shared__ double bcast[2];
    const int t = blockIdx.x, g = blockIdx.y;
    const T* row = x + static_cast<int64_t>(t) * ld;
    double s = 0.0, c = 0.0;
    for (int64_t j = threadIdx.x; j < len; j += kStatThreads) {
        if (gid == nullptr || gid[j] == g) {
            s += static_cast<double>(row[j]);
            c += 1.0;
        }
    }
    s = block_sum(s, scratch);
    c = block_sum(c, scratch);
    if (threadIdx.x == 0) {
        bcast[0] = c > 0.0 ? s / c : 0.0;
        bcast[1] = c;
    }
    __syncthreads();
    const double mean = bcast[0], cnt = bcast[1];
    double q = 0.0;
    for (int64_t j = threadIdx.x; j < len; j += kStatThreads) {
        if (gid == nullptr || gid[j] == g) {
            const double d = static_cast<double>(row[j]) - mean;
            q = fma(d, d, q);
        }
    }
    q = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        const int64_t o = static_cast<int64_t>(t) * n_groups + g;
        const double var = cnt > 0.0 ? q / cnt : 0.0;
        mean_out[o] = mean;
        var_out[o] = var;
        std_out[o] = sqrt(var);
        count_out[o] = static_cast<long long>(cnt);
    }
}

// x[t][j] = 0 where x[t][j] < T(mean + z std) of its (component, group); samples of no group are left alone
template <typename T>
__global__ void comp_threshold_kernel(T* __restrict__ x, int64_t k, int64_t len, int64_t ld,
                                      const int* __restrict__ gid, int n_groups, const double* __restrict__ mean,
                                      const double* __restrict__ sd, const long long* __restrict__ count, double z) {
    const int64_t total = k * len;
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t t = idx / len, j = idx - t * len;
        const int g = gid ? gid[j] : 0;
        if (g < 0 || g >= n_groups) continue;
        const int64_t o = t * n_groups + g;
        if (count[o] == 0) continue;
        // the reference forms the threshold in the array's own dtype (numpy scalar arithmetic) and compares there
        const T thr = static_cast<T>(mean[o] + z * sd[o]);
        T* p = x + t * ld + j;
        if (*p < thr) *p = static_cast<T>(0);
    }
}

// StandardScaler: out = (x - mean) / scale, scale = 1 for constant components
// (sklearn/preprocessing/_data.py: _is_constant_feature + _handle_zeros_in_scale)
template <typename T>
__global__ void comp_standardize_kernel(const T* __restrict__ x, int64_t k, int64_t len, int64_t ld,
                                        const double* __restrict__ mean, const double* __restrict__ var,
                                        T* __restrict__ out, int64_t ld_out) {
    const int64_t total = k * len;
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t t = idx / len, j = idx - t * len;
        const double mu = mean[t], v = var[t], n = static_cast<double>(len);
        const double eps = DBL_EPSILON;
        const double bound = n * eps * v + (n * mu * eps) * (n * mu * eps);
        double scale = sqrt(v);
        if (v <= bound || scale < 10.0 * eps) scale = 1.0;
        // X -= mean_; X /= scale_  -- two in-place operations, each rounded to the array's dtype
        const T centred = static_cast<T>(static_cast<double>(x[t * ld + j]) - mu);
        out[t * ld_out + j] = static_cast<T>(static_cast<double>(centred) / scale);
    }
}

// StandardScaler's scale for a feature of n values with the given mean / population variance: constant features
// (sklearn/preprocessing/_data.py _is_constant_feature) and vanishing scales (_handle_zeros_in_scale) get 1
__device__ __forceinline__ double scaler_scale(double mu, double v, double n) {
    const double eps = DBL_EPSILON;
    const double bound = n * eps * v + (n * mu * eps) * (n * mu * eps);
    const double scale = sqrt(v);
    return (v <= bound || scale < 10.0 * eps) ? 1.0 : scale;
}

// winner-takes-all map, one thread per sample j of the component-major map x[k][ld]:
//   z[t][j] = T(T(x[t][j] - mean) / scale)  (the two roundings of StandardScaler.transform on an array of dtype T)
//   out[j]  = 1 + first argmax_t z[t][j], or 0 when that maximum is below T(z_thr).
// per_component != 0: the scaler's features are the components (grey maps [len, k] on the host: statistics over the
// samples, precomputed by comp_stats_kernel); otherwise its features are the samples (white maps [k, len] on the
// host: mean / variance over the k components of sample j, summed here in component order).
template <typename T>
__global__ void comp_wta_kernel(const T* __restrict__ x, int64_t k, int64_t len, int64_t ld, int per_component,
                                const double* __restrict__ mean, const double* __restrict__ var, double z_thr,
                                long long* __restrict__ out) {
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < len;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double mu = 0.0, scale = 1.0;
        if (!per_component) {
            double s = 0.0;
            for (int64_t t = 0; t < k; ++t) s += static_cast<double>(x[t * ld + j]);
            mu = s / static_cast<double>(k);
            double q = 0.0;
            for (int64_t t = 0; t < k; ++t) {
                const double d = static_cast<double>(x[t * ld + j]) - mu;
                q = fma(d, d, q);
            }
            scale = scaler_scale(mu, q / static_cast<double>(k), static_cast<double>(k));
        }
        T best = static_cast<T>(0);
        int64_t arg = 0;
        for (int64_t t = 0; t < k; ++t) {
            if (per_component) {
                mu = mean[t];
                scale = scaler_scale(mu, var[t], static_cast<double>(len));
            }
            const T centred = static_cast<T>(static_cast<double>(x[t * ld + j]) - mu);
            const T z = static_cast<T>(static_cast<double>(centred) / scale);
            if (t == 0 || z > best) {      // strict: the first maximum wins, like np.argmax
                best = z;
                arg = t;
            }
        }
        out[j] = best < static_cast<T>(z_thr) ? 0LL : static_cast<long long>(arg + 1);
    }
}

// one CTA per component t: r_t = [sum_j (a_tj - mean a_t)(b_tj - mean b_t) / (len - 1)] / (std a_t * std b_t), population
// standard deviations -- the mixed normalisation of NFACT/stats/stats_component_loadings.py:227-231, kept as is
template <typename T>
__global__ void __launch_bounds__(kStatThreads)
comp_loadings_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t len, int64_t ld,
                     double* __restrict__ out) {
    __shared__ double scratch[32];
    __shared__ double bcast[2];
    const int t = blockIdx.x;
    const T* ra = a + static_cast<int64_t>(t) * ld;
    const T* rb = b + static_cast<int64_t>(t) * ld;
    double sa = 0.0, sb = 0.0;
    for (int64_t j = threadIdx.x; j < len; j += kStatThreads) {
        sa += static_cast<double>(ra[j]);
        sb += static_cast<double>(rb[j]);
    }
    sa = block_sum(sa, scratch);
    sb = block_sum(sb, scratch);
    if (threadIdx.x == 0) {
        bcast[0] = sa / static_cast<double>(len);
        bcast[1] = sb / static_cast<double>(len);
    }
    __syncthreads();
    const double ma = bcast[0], mb = bcast[1];
    double qab = 0.0, qaa = 0.0, qbb = 0.0;
    for (int64_t j = threadIdx.x; j < len; j += kStatThreads) {
        const double da = static_cast<double>(ra[j]) - ma, db = static_cast<double>(rb[j]) - mb;
        qab = fma(da, db, qab);
        qaa = fma(da, da, qaa);
        qbb = fma(db, db, qbb);
    }
    qab = block_sum(qab, scratch);
    qaa = block_sum(qaa, scratch);
    qbb = block_sum(qbb, scratch);
    if (threadIdx.x == 0) {
        const double n = static_cast<double>(len);
        out[t] = (qab / (n - 1.0)) / (sqrt(qaa / n) * sqrt(qbb / n));     // 0 / 0 -> NaN like numpy
    }
}

int elementwise_grid(int64_t total) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(total, 256), static_cast<int64_t>(sm_count()) * 16)));
}

}  // namespace

template <typename T>
int comp_threshold(T* x, int64_t k, int64_t len, int64_t ld, const int* gid, int n_groups, double z, double* stats,
                   cudaStream_t stream) {
    if (k * len == 0 || n_groups <= 0) return 0;
    const int64_t ng = static_cast<int64_t>(k) * n_groups;
    double* mean = stats;
    double* sd = stats + ng;
    double* var = stats + 2 * ng;
    long long* count = reinterpret_cast<long long*>(stats + 3 * ng);
    comp_stats_kernel<T><<<dim3(static_cast<unsigned>(k), static_cast<unsigned>(n_groups)), kStatThreads, 0, stream>>>(
        x, len, ld, gid, n_groups, mean, sd, var, count);
    NFB_CUDA(cudaGetLastError());
    comp_threshold_kernel<T><<<elementwise_grid(k * len), 256, 0, stream>>>(x, k, len, ld, gid, n_groups, mean, sd,
                                                                           count, z);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int comp_standardize(const T* x, int64_t k, int64_t len, int64_t ld, T* out, int64_t ld_out, double* stats,
                     cudaStream_t stream) {
    if (k * len == 0) return 0;
    double* mean = stats;
    double* sd = stats + k;
    double* var = stats + 2 * k;
    long long* count = reinterpret_cast<long long*>(stats + 3 * k);
    comp_stats_kernel<T><<<dim3(static_cast<unsigned>(k), 1), kStatThreads, 0, stream>>>(x, len, ld, nullptr, 1, mean,
                                                                                        sd, var, count);
    NFB_CUDA(cudaGetLastError());
    comp_standardize_kernel<T><<<elementwise_grid(k * len), 256, 0, stream>>>(x, k, len, ld, mean, var, out, ld_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int comp_loadings(const T* a, const T* b, int64_t k, int64_t len, int64_t ld, double* out, cudaStream_t stream) {
    if (k == 0) return 0;
    comp_loadings_kernel<T><<<static_cast<unsigned>(k), kStatThreads, 0, stream>>>(a, b, len, ld, out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int comp_wta(const T* x, int64_t k, int64_t len, int64_t ld, int per_component, double z_thr, double* stats,
             long long* out, cudaStream_t stream) {
    if (k * len == 0) return 0;
    double* mean = stats;
    double* sd = stats + k;
    double* var = stats + 2 * k;
    long long* count = reinterpret_cast<long long*>(stats + 3 * k);
    if (per_component) {
        comp_stats_kernel<T><<<dim3(static_cast<unsigned>(k), 1), kStatThreads, 0, stream>>>(x, len, ld, nullptr, 1,
                                                                                            mean, sd, var, count);
        NFB_CUDA(cudaGetLastError());
    }
    comp_wta_kernel<T><<<elementwise_grid(len), 256, 0, stream>>>(x, k, len, ld, per_component, mean, var, z_thr, out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

template int comp_loadings<float>(const float*, const float*, int64_t, int64_t, int64_t, double*, cudaStream_t);
template int comp_loadings<double>(const double*, const double*, int64_t, int64_t, int64_t, double*, cudaStream_t);
template int comp_wta<float>(const float*, int64_t, int64_t, int64_t, int, double, double*, long long*, cudaStream_t);
template int comp_wta<double>(const double*, int64_t, int64_t, int64_t, int, double, double*, long long*, cudaStream_t);
template int comp_threshold<float>(float*, int64_t, int64_t, int64_t, const int*, int, double, double*, cudaStream_t);
template int comp_threshold<double>(double*, int64_t, int64_t, int64_t, const int*, int, double, double*, cudaStream_t);
template int comp_standardize<float>(const float*, int64_t, int64_t, int64_t, float*, int64_t, double*, cudaStream_t);
template int comp_standardize<double>(const double*, int64_t, int64_t, int64_t, double*, int64_t, double*,
                                      cudaStream_t);

}  // namespace nfb
