/* Plain-C part of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see nfact_oracle.py).
 *
 * Two sequential kernels that numpy cannot express without changing the
 * floating-point evaluation order:
 *
 *   oracle_cd_sweep_f32  -- sklearn/decomposition/_cdnmf_fast.pyx:8-38 (float32 instance),
 *                           reached from NFACT/decomp/decomposition/sso/nmfsso.py:54-56
 *   oracle_nnls_f64      -- Lawson & Hanson NNLS (Solving Least Squares Problems, 1995,
 *                           ch. 23), the algorithm behind scipy.optimize.nnls
 *                           (scipy/optimize/_nnls.py), reached from
 *                           NFACT/dual_reg/dual_regression/dual_regression_methods.py:139,153,264
 *
 * Build:  gcc -O2 -ffp-contract=off -shared -fPIC -o liboracle_c.so oracle_c.c -lm
 * (-ffp-contract=off: the sklearn wheel targets baseline x86-64 without FMA, so the
 *  mul and the add of the gradient accumulation round separately.)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* One coordinate-descent sweep over all components for every row of W (in place).
 * W [n_samples, k] C order, HHt [k, k], XHt [n_samples, k].  Returns the violation,
 * accumulated sequentially in float32 like the Cython code. */
float oracle_cd_sweep_f32(float *W, const float *HHt, const float *XHt, long n_samples, long k)
{
    float violation = 0.0f;
    for (long t = 0; t < k; ++t) {
        for (long i = 0; i < n_samples; ++i) {
            float grad = -XHt[i * k + t];
            for (long r = 0; r < k; ++r)
                grad += HHt[t * k + r] * W[i * k + r];
            float pg = (W[i * k + t] == 0.0f) ? (grad < 0.0f ? grad : 0.0f) : grad;
            violation += fabsf(pg);
            float hess = HHt[t * k + t];
            if (hess != 0.0f) {
                float v = W[i * k + t] - grad / hess;
                W[i * k + t] = v > 0.0f ? v : 0.0f;
            }
        }
    }
    return violation;
}

/* Least squares on a column subset by Householder QR (fresh factorisation).
 * A [m, n] C order; cols[p] selects the passive columns; z[p] receives the solution. */
static void ls_subset(const double *A, const double *b, long m, long n, const long *cols, long p,
                      double *work, double *rhs, double *z)
{
    /* work: m x p column-major copy */
    for (long c = 0; c < p; ++c)
        for (long i = 0; i < m; ++i)
            work[c * m + i] = A[i * n + cols[c]];
    memcpy(rhs, b, (size_t)m * sizeof(double));
    for (long c = 0; c < p; ++c) {
        double *col = work + c * m;
        double norm = 0.0;
        for (long i = c; i < m; ++i) norm += col[i] * col[i];
        norm = sqrt(norm);
        if (norm == 0.0) continue;
        double alpha = col[c] > 0.0 ? -norm : norm;
        double v0 = col[c] - alpha;
        /* v = (v0, col[c+1..]) ; beta = 2 / (v'v) */
        double vtv = v0 * v0;
        for (long i = c + 1; i < m; ++i) vtv += col[i] * col[i];
        if (vtv == 0.0) continue;
        double beta = 2.0 / vtv;
        for (long c2 = c + 1; c2 < p; ++c2) {
            double *o = work + c2 * m;
            double dot = v0 * o[c];
            for (long i = c + 1; i < m; ++i) dot += col[i] * o[i];
            dot *= beta;
            o[c] -= dot * v0;
            for (long i = c + 1; i < m; ++i) o[i] -= dot * col[i];
        }
        double dot = v0 * rhs[c];
        for (long i = c + 1; i < m; ++i) dot += col[i] * rhs[i];
        dot *= beta;
        rhs[c] -= dot * v0;
        for (long i = c + 1; i < m; ++i) rhs[i] -= dot * col[i];
        col[c] = alpha; /* R diagonal; the sub-diagonal keeps v but is not read again */
    }
    for (long c = p - 1; c >= 0; --c) {
        double acc = rhs[c];
        for (long c2 = c + 1; c2 < p; ++c2) acc -= work[c2 * m + c] * z[c2];
        double diag = work[c * m + c];
        z[c] = diag != 0.0 ? acc / diag : 0.0;
    }
}

/* Lawson-Hanson active-set NNLS.  Returns 1 on success, 3 when maxiter inner iterations
 * were used up (scipy raises RuntimeError for that code).  x [n] out, *rnorm = ||Ax-b||. */
int oracle_nnls_f64(const double *A, const double *b, long m, long n, long maxiter,
                    double *x, double *rnorm)
{
    char *passive = (char *)calloc((size_t)n, 1);
    long *cols = (long *)malloc((size_t)n * sizeof(long));
    double *w = (double *)malloc((size_t)n * sizeof(double));
    double *s = (double *)calloc((size_t)n, sizeof(double));
    double *z = (double *)malloc((size_t)n * sizeof(double));
    double *resid = (double *)malloc((size_t)m * sizeof(double));
    double *work = (double *)malloc((size_t)m * (size_t)n * sizeof(double));
    double *rhs = (double *)malloc((size_t)m * sizeof(double));
    long iter = 0, np = 0;
    int info = 1;
    memset(x, 0, (size_t)n * sizeof(double));

    for (;;) {
        /* dual vector w = A'(b - A x) */
        for (long i = 0; i < m; ++i) {
            double acc = b[i];
            for (long j = 0; j < n; ++j) acc -= A[i * n + j] * x[j];
            resid[i] = acc;
        }
        for (long j = 0; j < n; ++j) {
            double acc = 0.0;
            for (long i = 0; i < m; ++i) acc += A[i * n + j] * resid[i];
            w[j] = acc;
        }
        if (np == n || np >= m) break;
        int accepted = 0;
        while (!accepted) {
            long t = -1;
            double wmax = 0.0;
            for (long j = 0; j < n; ++j)
                if (!passive[j] && w[j] > wmax) { wmax = w[j]; t = j; }
            if (t < 0) break;
            /* candidate t joins P; accept only if its least-squares coefficient is positive */
            passive[t] = 1;
            long p = 0;
            for (long j = 0; j < n; ++j) if (passive[j]) cols[p++] = j;
            ls_subset(A, b, m, n, cols, p, work, rhs, z);
            double zt = 0.0;
            for (long c = 0; c < p; ++c) if (cols[c] == t) zt = z[c];
            if (zt > 0.0) { accepted = 1; np = p; }
            else { passive[t] = 0; w[t] = 0.0; }
        }
        if (!accepted) break;
        /* inner loop: step towards z until it is feasible */
        for (;;) {
            long p = 0;
            for (long j = 0; j < n; ++j) if (passive[j]) cols[p++] = j;
            memset(s, 0, (size_t)n * sizeof(double));
            for (long c = 0; c < p; ++c) s[cols[c]] = z[c];
            int feasible = 1;
            for (long c = 0; c < p; ++c) if (z[c] <= 0.0) { feasible = 0; break; }
            if (feasible) break;
            if (++iter > maxiter) { info = 3; break; }
            double alpha = 2.0;
            long jmin = -1;
            for (long c = 0; c < p; ++c) {
                long j = cols[c];
                if (s[j] <= 0.0) {
                    double ratio = x[j] / (x[j] - s[j]);
                    if (ratio < alpha) { alpha = ratio; jmin = j; }
                }
            }
            for (long j = 0; j < n; ++j) x[j] += alpha * (s[j] - x[j]);
            /* every passive index driven to zero returns to the active set Z */
            for (long c = 0; c < p; ++c) {
                long j = cols[c];
                if (j == jmin || x[j] <= 0.0) { passive[j] = 0; x[j] = 0.0; }
            }
            p = 0;
            for (long j = 0; j < n; ++j) if (passive[j]) cols[p++] = j;
            np = p;
            if (p == 0) { memset(s, 0, (size_t)n * sizeof(double)); break; }
            ls_subset(A, b, m, n, cols, p, work, rhs, z);
        }
        if (info == 3) break;
        memcpy(x, s, (size_t)n * sizeof(double));
    }
    double acc2 = 0.0;
    for (long i = 0; i < m; ++i) {
        double acc = b[i];
        for (long j = 0; j < n; ++j) acc -= A[i * n + j] * x[j];
        acc2 += acc * acc;
    }
    *rnorm = sqrt(acc2);
    free(passive); free(cols); free(w); free(s); free(z); free(resid); free(work); free(rhs);
    return info;
}
