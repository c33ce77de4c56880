// Batched non-negative least squares on a shared Gram matrix (the nfact_dr hot loop).
//
// The reference solves, for every target column j (stage 1) / seed row i (stage 2),
//     h_j = argmin_{h >= 0} || A h - x_j ||_2        scipy.optimize.nnls(A, x_j)  (float64, Lawson-Hanson)
// (NFACT/dual_reg/dual_regression/dual_regression_methods.py:137-163).  Every one of those problems shares
// A, so they share the k x k Gram matrix G = A'A and differ only in the right-hand side r_j = A'x_j:
//     h_j = argmin_{h >= 0} 0.5 h'Gh - r_j'h .
// G and all r_j come out of the tensor-core GEMMs; this kernel solves the n (or m) independent k-dim
// QPs in float64, one warp per problem, by exact coordinate minimisation (the same projected
// Newton step as the NMF coordinate-descent sweep) with the gradient g = Gh - r kept incrementally:
// a coordinate whose value does not move costs O(1), a move costs one column of G.  Full sweeps detect
// support changes; between them only the current support is swept.  The minimiser of a strictly convex
// QP is unique, so on full-column-rank A the fixed point equals the active-set solution of scipy;
// convergence is declared on the KKT residual max_t |pg_t| <= tol * max|r|.
#include <algorithm>
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <int KPL>  // coordinates per lane: k <= 32 * KPL
__global__ void __launch_bounds__(128)
nnls_cd_kernel(const double* __restrict__ gram, int k, const float* __restrict__ rhs, int rhs_splits,
               int64_t rhs_stride, int64_t ld_rhs, int64_t cols, double* __restrict__ sol, int64_t ld_sol,
               int max_outer, double tol, int* __restrict__ status, unsigned long long* __restrict__ steps_total) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int64_t num_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;

    for (int64_t j = warp_global; j < cols; j += num_warps) {
        double h[KPL], g[KPL];
        double rmax = 0.0;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int t = lane + 32 * i;
            double r = 0.0;
            if (t < k)
                for (int s = 0; s < rhs_splits; ++s) r += static_cast<double>(rhs[s * rhs_stride + t * ld_rhs + j]);
            h[i] = 0.0;
            g[i] = -r;
            rmax = fmax(rmax, fabs(r));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        const double thresh = tol * rmax;

        // one exact coordinate step on component t; returns |projected gradient| before the step
        auto step = [&](int t) -> double {
            const int owner = t & 31, slot = t >> 5;
            double gt = 0.0, ht = 0.0;
#pragma unroll
            for (int i = 0; i < KPL; ++i)
                if (i == slot) { gt = g[i]; ht = h[i]; }
            gt = shfl_d(gt, owner);
            ht = shfl_d(ht, owner);
            const double gtt = __ldg(gram + static_cast<int64_t>(t) * k + t);
            const double pg = (ht == 0.0) ? fmin(gt, 0.0) : gt;
            if (gtt > 0.0) {
                const double hn = fmax(ht - gt / gtt, 0.0);
                const double delta = hn - ht;
                if (delta != 0.0) {  // warp-uniform
                    const double* gcol = gram + static_cast<int64_t>(t) * k;  // G symmetric: row t == column t
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int u = lane + 32 * i;
                        if (u < k) g[i] = fma(delta, __ldg(gcol + u), g[i]);
                        if (i == slot && lane == owner) h[i] = hn;
                    }
                }
            }
            return fabs(pg);
        };

        bool converged = (rmax == 0.0);
        unsigned long long nsteps = 0;
        for (int outer = 0; outer < max_outer && !converged; ++outer) {
            double viol = 0.0;
            for (int t = 0; t < k; ++t) viol = fmax(viol, step(t));
            nsteps += k;
            if (viol <= thresh) { converged = true; break; }
            // sweeps restricted to the current support
            for (int inner = 0; inner < 64; ++inner) {
                double iv = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    unsigned mask = __ballot_sync(0xffffffffu, h[i] > 0.0);
                    while (mask) {
                        const int bit = __ffs(mask) - 1;
                        mask &= mask - 1;
                        iv = fmax(iv, step(32 * i + bit));
                        ++nsteps;
                    }
                }
                if (iv <= thresh) break;
            }
        }
        if (!converged && lane == 0) atomicAdd(status, 1);
        if (lane == 0 && steps_total != nullptr) atomicAdd(steps_total, nsteps);
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int t = lane + 32 * i;
            if (t < k) sol[t * ld_sol + j] = h[i];
        }
    }
}


// ---- fast path: packed lower triangle of G (float64) resident in shared memory ----------------------
// One CTA per SM, 24 warps, columns handed out dynamically (their sweep counts differ widely).  The
// coordinate step has its slot (register index) as a compile-time constant, multiplies by a precomputed
// 1/G[t][t] and reads row/column t of G from shared memory.

constexpr int kInitialSweeps = 6;          // Gauss-Seidel passes before the first CG phase
constexpr double kFaceReduction = 1e-2;    // residual reduction asked of one CG phase
constexpr int kGreedyBudget = 8;           // greedy coordinate steps per problem, in multiples of k
constexpr int kGreedyBatch = 3;            // steps taken per pass over the gradient keys

template <int KPL, int kNnlsWarps>
__global__ void __launch_bounds__(kNnlsWarps * 32, 1)
nnls_smem_kernel(const double* __restrict__ gram, int k, const float* __restrict__ rhs, int rhs_splits,
                 int64_t rhs_stride, int64_t ld_rhs, int64_t cols, double* __restrict__ sol, int64_t ld_sol,
                 int max_outer, double tol, int* __restrict__ status, unsigned long long* __restrict__ steps_total,
                 unsigned long long* __restrict__ next_col, int initial_sweeps, double face_reduction,
                 int scale_diag, int greedy_budget, int greedy_batch,
                 const unsigned long long* __restrict__ flagged) {
    // flagged != nullptr: clean-up pass behind nnls_greedy_kernel -- only the columns that kernel marked (negative
    // row 0 of their solution) are finished, by sweeps + CG from the state its greedy steps left (what the greedy
    // budget did not finish is ill-conditioned: more greedy steps are the wrong tool); nothing to do (the usual case)
    // costs one global load
    if (flagged != nullptr && *flagged == 0ULL) return;
    extern __shared__ double smd[];
    const int ntri = k * (k + 1) / 2;
    double* g_tri = smd;                                 // k (k + 1) / 2 packed lower triangle, row-major
    double* zeros = smd + ntri;                          // k zeros: "column t" of the padding rows u >= k
    double* inv_diag = zeros + k;                        // k  (0 where G[t][t] <= 0)
    double* scl = inv_diag + k;                          // k  symmetric diagonal scaling s_t = 1 / sqrt(G[t][t])
    const int lane = threadIdx.x & 31;
    // The problems are solved in the variables h' = h / s with G' = S G S (unit diagonal), r' = S r: the
    // constraint h >= 0 and the coordinate steps are invariant under a positive diagonal scaling, and plain CG
    // on G' is Jacobi-preconditioned CG on G -- the component norms of NMF maps spread over orders of magnitude.
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const double d = gram[static_cast<int64_t>(t) * k + t];
        scl[t] = (scale_diag && d > 0.0) ? 1.0 / sqrt(d) : 1.0;
        zeros[t] = 0.0;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
        const int r = idx / k, c = idx - r * k;
        if (c <= r) g_tri[r * (r + 1) / 2 + c] = gram[idx] * scl[r] * scl[c];
    }
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const double d = gram[static_cast<int64_t>(t) * k + t] * scl[t] * scl[t];
        inv_diag[t] = d > 0.0 ? 1.0 / d : 0.0;
    }
    // G[u][t] for this lane's coordinates u = lane + 32 i:
    //   u >= t : g_tri[ro[i] + t]  with ro[i] = u (u + 1) / 2  (padding rows point at the zero region)
    //   u <  t : g_tri[t (t + 1) / 2 + u]  -- a contiguous run of row t, addressed with immediate offsets
    int ro[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int u = lane + 32 * i;
        ro[i] = u < k ? u * (u + 1) / 2 : ntri;
    }
    // q[i] += dt * G[lane + 32 i][t] for all i, t = 32 I + o
    auto axpy_col = [&](auto slot_tag, int o, double dt, double (&q)[KPL]) {
        constexpr int I = decltype(slot_tag)::value;
        const int t = 32 * I + o;
        const double* rowt = g_tri + t * (t + 1) / 2 + lane;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            if (i < I) q[i] = fma(dt, rowt[32 * i], q[i]);
            else if (i == I) q[i] = fma(dt, lane < o ? rowt[32 * i] : g_tri[ro[i] + t], q[i]);
            else q[i] = fma(dt, g_tri[ro[i] + t], q[i]);
        }
    };
    __syncthreads();

    unsigned pending = 0u;                 // flagged mode: marked columns of the current chunk of 32
    unsigned long long chunk_base = 0;
    for (;;) {
        unsigned long long j = 0;
        if (flagged == nullptr) {
            if (lane == 0) j = atomicAdd(next_col, 1ULL);
            j = __shfl_sync(0xffffffffu, j, 0);
            if (j >= static_cast<unsigned long long>(cols)) break;
        } else {
            bool out_of_work = false;
            while (pending == 0u) {
                unsigned long long c = 0;
                if (lane == 0) c = atomicAdd(next_col, 1ULL);
                chunk_base = __shfl_sync(0xffffffffu, c, 0) * 32ULL;
                if (chunk_base >= static_cast<unsigned long long>(cols)) { out_of_work = true; break; }
                const unsigned long long jj = chunk_base + lane;
                const bool marked = jj < static_cast<unsigned long long>(cols) && sol[jj] < 0.0;
                pending = __ballot_sync(0xffffffffu, marked);
            }
            if (out_of_work) break;
            j = chunk_base + (__ffs(pending) - 1);
            pending &= pending - 1;
        }
        double h[KPL], g[KPL];
        double rmax = 0.0;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int t = lane + 32 * i;
            double r = 0.0;
            if (t < k)
                for (int s = 0; s < rhs_splits; ++s) r += static_cast<double>(rhs[s * rhs_stride + t * ld_rhs + j]);
            if (t < k) r *= scl[t];
            h[i] = 0.0;
            g[i] = t < k ? -r : 0.0;         // padding slots: +0, they stay +0 (their Gram entries are zeros)
            rmax = fmax(rmax, fabs(r));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        const double thresh = tol * rmax;
        unsigned long long nsteps = 0;
        if (flagged != nullptr) {
            // warm start: h' = h / s from the greedy kernel's output, g = G'h' - r'
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int t = lane + 32 * i;
                double v = t < k ? sol[t * ld_sol + j] : 0.0;
                if (t == 0) v = v <= -1.0e-300 ? -v : 0.0;          // the marker: -h_0, or -DBL_MIN for h_0 == 0
                h[i] = t < k ? v / scl[t] : 0.0;
            }
            auto add_columns = [&](auto slot_tag) {
                constexpr int I = decltype(slot_tag)::value;
                unsigned mask = __ballot_sync(0xffffffffu, h[I] > 0.0);
                while (mask) {
                    const int o = __ffs(mask) - 1;
                    mask &= mask - 1;
                    axpy_col(slot_tag, o, shfl_d(h[I], o), g);
                }
            };
            [&]<int... Is>(std::integer_sequence<int, Is...>) {
                (add_columns(std::integral_constant<int, Is>{}), ...);
            }(std::make_integer_sequence<int, KPL>{});
        }

        // exact coordinate step on component t = 32 * I + o (I compile-time); returns |projected gradient|
        auto step = [&](auto slot_tag, int o) -> double {
            constexpr int I = decltype(slot_tag)::value;
            const int t = 32 * I + o;
            const double gt = shfl_d(g[I], o);
            const double ht = shfl_d(h[I], o);
            const double inv = inv_diag[t];
            const double pg = (ht == 0.0) ? fmin(gt, 0.0) : gt;
            const double hn = fmax(fma(-gt, inv, ht), 0.0);
            const double delta = inv != 0.0 ? hn - ht : 0.0;
            if (delta != 0.0) {  // warp-uniform
                axpy_col(slot_tag, o, delta, g);
                if (lane == o) h[I] = hn;
            }
            return fabs(pg);
        };

        bool converged = (rmax == 0.0);
        // ---- greedy (Gauss-Southwell) coordinate descent ------------------------------------------------------
        // Every step takes the coordinate with the largest projected gradient (unit diagonal after the scaling, so
        // that is also the largest decrease of the objective) and minimises over it exactly.  On NMF-type problems
        // the support is found and solved to the KKT tolerance in 2-3 k column updates, against 5-10 k for cyclic
        // sweeps + CG (simulated on C3-shaped problems, scratch/nnls_sim7.py): a cyclic sweep spends most of its
        // updates on coordinates that enter the support only to leave it again.  The selection costs no shared-
        // memory traffic: a 24-bit key (sign-stripped high word of the gradient, truncated) + the coordinate index
        // per slot, one redux.sync per step.  The sweeps + CG loop below stays as the fallback for whatever the
        // budget did not finish (ill-conditioned Gram matrices).
        bool greedy_used = flagged != nullptr;
        if (greedy_budget > 0 && !converged && flagged == nullptr) {
            greedy_used = true;
            // stop at half the tolerance: the truncated keys under-estimate by < 2^-11, the exact check follows
            const unsigned stop_key = static_cast<unsigned>(__double2hiint(0.5 * thresh)) & 0x7fffff00u;
            auto greedy_step = [&](auto slot_tag, int o) -> bool {
                constexpr int I = decltype(slot_tag)::value;
                const int t = 32 * I + o;
                const double gt = shfl_d(g[I], o);
                const double ht = shfl_d(h[I], o);
                const double inv = inv_diag[t];
                const double hn = fmax(fma(-gt, inv, ht), 0.0);
                const double delta = inv != 0.0 ? hn - ht : 0.0;
                if (delta != 0.0) {  // warp-uniform
                    axpy_col(slot_tag, o, delta, g);
                    if (lane == o) h[I] = hn;
                }
                return delta != 0.0;
            };
            for (int left = greedy_budget; left > 0;) {
                // this lane's best candidate.  Projected gradient: |g| on the support (h > 0), max(-g, 0) off it;
                // the padding slots hold h = 0, g = +0 and never qualify
                unsigned mine = 0u;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int ghi = __double2hiint(g[i]);
                    const bool take = __double2hiint(h[i]) > 0 || ghi < 0;
                    const unsigned key = (static_cast<unsigned>(ghi) & 0x7fffff00u) | static_cast<unsigned>(32 * i + lane);
                    mine = max(mine, take ? key : 0u);
                }
                // one key pass serves up to greedy_batch steps: the best candidates of distinct lanes, largest
                // first.  The later ones were chosen on a gradient that is one or two updates old; the step itself
                // is always exact on the current gradient (a candidate that no longer moves costs no update).
                bool any_moved = false, done = false;
                for (int b = 0; b < greedy_batch && left > 0; ++b) {
                    const unsigned best = __reduce_max_sync(0xffffffffu, mine);
                    if ((best & 0x7fffff00u) <= stop_key) { done = (b == 0); break; }
                    const int t = static_cast<int>(best & 0xffu);
                    if (lane == (t & 31)) mine = 0u;
                    bool moved = false;
                    [&]<int... Is>(std::integer_sequence<int, Is...>) {
                        ((((t >> 5) == Is) ? (void)(moved = greedy_step(std::integral_constant<int, Is>{}, t & 31)) : (void)0), ...);
                    }(std::make_integer_sequence<int, KPL>{});
                    any_moved |= moved;
                    ++nsteps;
                    --left;
                }
                if (done || !any_moved) break;   // converged on fresh keys / steps lost to rounding: exact check below
            }
            double viol = 0.0;
#pragma unroll
            for (int i = 0; i < KPL; ++i) viol = fmax(viol, h[i] > 0.0 ? fabs(g[i]) : fmax(-g[i], 0.0));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) viol = fmax(viol, __shfl_xor_sync(0xffffffffu, viol, o));
            converged = viol <= thresh;
        }
        for (int outer = 0; outer < max_outer && !converged; ++outer) {
            // Gauss-Seidel sweeps over the coordinates that can move: the support (h > 0) and the candidates to
            // enter it (g < 0).  Everything else has a zero projected gradient and a step there is a no-op, so a
            // sweep costs |P| + |E| column updates instead of k visits.  The sweeps find the support -- they are
            // as cheap as a CG iteration and change many coordinates at once --, several of them run before the
            // first CG phase (from h = 0 the first pass is only a rough guess), one between CG phases.
            auto sweep = [&](auto slot_tag) {
                constexpr int I = decltype(slot_tag)::value;
                unsigned mask = __ballot_sync(0xffffffffu, h[I] > 0.0 || g[I] < 0.0);
                nsteps += __popc(mask);
                while (mask) {
                    const int o = __ffs(mask) - 1;
                    mask &= mask - 1;
                    step(slot_tag, o);
                }
            };
            const int n_sweeps = (outer == 0 && !greedy_used) ? initial_sweeps : 1;
            for (int sw = 0; sw < n_sweeps; ++sw) {
                [&]<int... Is>(std::integer_sequence<int, Is...>) {
                    (sweep(std::integral_constant<int, Is>{}), ...);
                }(std::make_integer_sequence<int, KPL>{});
            }
            // KKT residual of the state the sweeps left: max_t |projected gradient|
            double viol = 0.0;
#pragma unroll
            for (int i = 0; i < KPL; ++i) viol = fmax(viol, h[i] > 0.0 ? fabs(g[i]) : fmax(-g[i], 0.0));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) viol = fmax(viol, __shfl_xor_sync(0xffffffffu, viol, o));
            if (viol <= thresh) { converged = true; break; }
            // ---- conjugate gradients on the current support P = {h > 0} --------------------------------
            // Between two sweeps the support is fixed and the problem is the linear system
            // G_PP z = r_P: CG needs O(sqrt(cond)) products with G where coordinate sweeps need O(cond).
            // The step is cut at the first component that would turn negative; that component leaves P and CG
            // restarts on the reduced support; the next sweep re-examines what may enter P.  A phase only
            // reduces the residual on its face by face_reduction (in norm): solving a face that is still going
            // to change to full accuracy is wasted work, and the last phases reach the tolerance anyway.
            double d[KPL], q[KPL];
            bool in_p[KPL];
            double rr = 0.0;
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                in_p[i] = h[i] > 0.0;
                d[i] = in_p[i] ? -g[i] : 0.0;
                rr = fma(d[i], d[i], rr);
            }
            rr = warp_sum(rr);
            const double rr_stop = fmax(0.01 * thresh * thresh, face_reduction * face_reduction * rr);
            for (int it = 0; it < 2 * k && rr > rr_stop; ++it) {
#pragma unroll
                for (int i = 0; i < KPL; ++i) q[i] = 0.0;
                auto accumulate = [&](auto slot_tag) {
                    constexpr int I = decltype(slot_tag)::value;
                    unsigned mask = __ballot_sync(0xffffffffu, d[I] != 0.0);
                    nsteps += __popc(mask);
                    while (mask) {
                        const int o = __ffs(mask) - 1;
                        mask &= mask - 1;
                        axpy_col(slot_tag, o, shfl_d(d[I], o), q);
                    }
                };
                [&]<int... Is>(std::integer_sequence<int, Is...>) {
                    (accumulate(std::integral_constant<int, Is>{}), ...);
                }(std::make_integer_sequence<int, KPL>{});
                double dq = 0.0, amax = 1.0e300;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    dq = fma(d[i], q[i], dq);
                    if (in_p[i] && d[i] < 0.0) amax = fmin(amax, -h[i] / d[i]);
                }
                dq = warp_sum(dq);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) amax = fmin(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                if (!(dq > 0.0)) break;                       // direction in the null space: leave it to the sweeps
                double alpha = rr / dq;
                const bool hit = alpha >= amax;
                if (hit) alpha = amax;
                double rr_new = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const double h_new = fma(alpha, d[i], h[i]);
                    g[i] = fma(alpha, q[i], g[i]);
                    // a component that reached the bound (exactly, or within rounding of the cut step) leaves P
                    const bool blocked = in_p[i] && d[i] < 0.0 && (h_new <= 0.0 || (hit && -h[i] / d[i] <= amax));
                    h[i] = blocked ? 0.0 : h_new;
                    if (in_p[i]) rr_new = fma(g[i], g[i], rr_new);
                }
                if (hit) {
                    // the blocking component(s) left P: restart CG on the reduced support instead of paying a
                    // full sweep per dropped component (the sweep is only needed to let components ENTER P)
                    rr = 0.0;
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        in_p[i] = in_p[i] && h[i] > 0.0;
                        d[i] = in_p[i] ? -g[i] : 0.0;
                        rr = fma(d[i], d[i], rr);
                    }
                    rr = warp_sum(rr);
                    continue;
                }
                rr_new = warp_sum(rr_new);
                const double beta = rr_new / rr;
#pragma unroll
                for (int i = 0; i < KPL; ++i) d[i] = in_p[i] ? fma(beta, d[i], -g[i]) : 0.0;
                rr = rr_new;
            }
        }
        if (!converged && lane == 0) atomicAdd(status, 1);
        if (lane == 0 && steps_total != nullptr) atomicAdd(steps_total, nsteps);
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int t = lane + 32 * i;
            if (t < k) sol[t * ld_sol + j] = h[i] * scl[t];
        }
    }
}


// ---- greedy-only kernel: 32 warps per SM ------------------------------------------------------------------
// The greedy (Gauss-Southwell) phase of nnls_smem_kernel on its own.  It needs only the gradient and the solution
// (no CG vectors), so it fits 64 registers per thread and runs 32 warps per SM instead of 16: a coordinate step is
// one dependent chain (key pass -> redux -> shuffle -> 7 LDS.64 + DFMA), and what hides a dependent chain is more
// chains per SM.  The scaled Gram matrix G' = S G S lives in shared memory as 32 x 32 blocks (row stride 33) of the
// block lower triangle, diagonal blocks in full: row t of G' restricted to block column i <= I is the contiguous run
// [o][lane] of block (I, i), its continuation below the diagonal block is column o of block (i, I), i > I,
// [lane][o] -- both conflict-free, both addressed as one per-step base + an immediate, no per-lane row offsets and
// no select on the diagonal.  Problems the step budget does not finish are marked (NaN in row 0 of their solution)
// and counted; nnls_smem_kernel solves exactly those afterwards.
constexpr int kGBlk = 32 * 33;     // doubles per padded block

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__host__ __device__ constexpr int greedy_block_off(int R, int C, int last, int rl33) {
    return R < last ? (R * (R + 1) / 2 + C) * kGBlk : (last * (last + 1) / 2) * kGBlk + C * rl33;
}
__host__ inline size_t greedy_smem_bytes(int k) {
    const int nb = (k + 31) / 32, rows_last = k - 32 * (nb - 1);
    return (static_cast<size_t>(greedy_block_off(nb - 1, nb, nb - 1, rows_last * 33)) + k) * sizeof(double);
}

// byte offsets of the blocks (last block row, C): their row count depends on k, so they are not immediates -- handed
// to the kernel as arguments (constant bank operands) instead of being rebuilt from k in every step
struct GreedyLastOff { int v[8]; };

template <int NB, int kWarps>      // NB = ceil(k / 32): block rows = coordinates per lane
__global__ void __launch_bounds__(kWarps * 32, 1)
nnls_greedy_kernel(const double* __restrict__ gram, int k, const float* __restrict__ rhs, int rhs_splits,
                   int64_t rhs_stride, int64_t ld_rhs, int64_t cols, double* __restrict__ sol, int64_t ld_sol,
                   double tol, unsigned long long* __restrict__ steps_total, unsigned long long* __restrict__ counters,
                   int budget, int batch, const GreedyLastOff last_off) {
    extern __shared__ double smd[];
    constexpr int kLast = NB - 1;
    const int rows_last = k - 32 * kLast;
    const int rl33 = rows_last * 33;
    double* scl = smd + greedy_block_off(kLast, NB, kLast, rl33);      // k scales s_t = 1 / sqrt(G[t][t])
    const int lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const double d = gram[static_cast<int64_t>(t) * k + t];
        scl[t] = d > 0.0 ? 1.0 / sqrt(d) : 1.0;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
        const int r = idx / k, c = idx - r * k;
        const int R = r >> 5, C = c >> 5;
        if (C <= R) smd[greedy_block_off(R, C, kLast, rl33) + (r & 31) * 33 + (c & 31)] = gram[idx] * scl[r] * scl[c];
    }
    __syncthreads();
    const bool last_valid = lane < rows_last;          // the last slot holds coordinates 32 (NB - 1) + lane < k only
    const uint32_t sa_lane = smem_u32(smd) + lane * 8, sb_lane = smem_u32(smd) + lane * (33 * 8);

    for (;;) {
        unsigned long long j = 0;
        if (lane == 0) j = atomicAdd(counters, 1ULL);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= static_cast<unsigned long long>(cols)) break;
        double h[NB], g[NB];
        double rmax = 0.0;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int t = lane + 32 * i;
            double r = 0.0;
            if (t < k) {
                for (int s = 0; s < rhs_splits; ++s) r += static_cast<double>(rhs[s * rhs_stride + t * ld_rhs + j]);
                r *= scl[t];
            }
            h[i] = 0.0;
            g[i] = t < k ? -r : 0.0;         // padding slots: +0, they stay +0 and never qualify
            rmax = fmax(rmax, fabs(r));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        const double thresh = tol * rmax;
        const unsigned stop_key = static_cast<unsigned>(__double2hiint(0.5 * thresh)) & 0x7fffff00u;

        // exact minimisation over coordinate t = 32 I + o (unit diagonal): every lane forms the step of its own
        // slot-I coordinate, lane o's is broadcast; g += delta * G'[:, t].  A step that does not move (a candidate
        // picked on a stale gradient, or a move lost to rounding) costs no update; the budget bounds the loop.
        auto step = [&](auto slot_tag, int o) {
            constexpr int I = decltype(slot_tag)::value;
            // max(h - g, 0) on the sign bit: three integer operations instead of the six of a NaN-aware fmax
            const double hm = h[I] - g[I];
            const int keep = ~(__double2hiint(hm) >> 31);
            const double hn = __hiloint2double(__double2hiint(hm) & keep, __double2loint(hm) & keep);
            const double delta = shfl_d(hn - h[I], o);
            if (delta == 0.0) return;                             // warp-uniform
            if (lane == o) h[I] = hn;
            const uint32_t pa = sa_lane + o * (33 * 8);           // [o][lane] of the blocks (I, i), i <= I
            const uint32_t pb = sb_lane + o * 8;                  // [lane][o] of the blocks (i, I), i > I
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                double v = 0.0;
                if (i < kLast || last_valid) {
                    if constexpr (I == kLast) v = lds_f64(pa + last_off.v[i]);                 // blocks (last, i)
                    else if (i == kLast) v = lds_f64(pb + last_off.v[I]);                      // block (last, I)
                    else v = i <= I ? lds_f64(pa + 8 * greedy_block_off(I, i, kLast, rl33))
                                    : lds_f64(pb + 8 * greedy_block_off(i, I, kLast, rl33));
                }
                g[i] = fma(delta, v, g[i]);
            }
        };

        int left = rmax != 0.0 ? budget : 0;
        while (left > 0) {
            // this lane's best candidate: |g| on the support (h > 0), max(-g, 0) off it, as a 24-bit key
            // (sign-stripped, truncated high word) + the coordinate index
            unsigned mine = 0u;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int ghi = __double2hiint(g[i]);
                const bool take = __double2hiint(h[i]) > 0 || ghi < 0;
                const unsigned key = (static_cast<unsigned>(ghi) & 0x7fffff00u) | static_cast<unsigned>(32 * i + lane);
                mine = max(mine, take ? key : 0u);
            }
            // one key pass serves up to `batch` steps: the best candidates of distinct lanes, largest first (the
            // later ones chosen on a gradient that is one or two updates old; the step itself is always exact)
            bool done = false;
            for (int b = 0; b < batch && left > 0; ++b) {
                const unsigned best = __reduce_max_sync(0xffffffffu, mine);
                if ((best & 0x7fffff00u) <= stop_key) { done = (b == 0); break; }
                const int o = static_cast<int>(best & 31u);
                if (lane == o) mine = 0u;
                switch ((best >> 5) & 7u) {
                    case 0: step(std::integral_constant<int, 0>{}, o); break;
                    case 1: if constexpr (NB > 1) step(std::integral_constant<int, 1>{}, o); break;
                    case 2: if constexpr (NB > 2) step(std::integral_constant<int, 2>{}, o); break;
                    case 3: if constexpr (NB > 3) step(std::integral_constant<int, 3>{}, o); break;
                    case 4: if constexpr (NB > 4) step(std::integral_constant<int, 4>{}, o); break;
                    case 5: if constexpr (NB > 5) step(std::integral_constant<int, 5>{}, o); break;
                    default: if constexpr (NB > 6) step(std::integral_constant<int, 6>{}, o); break;
                }
                --left;
            }
            if (done) break;                     // converged on fresh keys: the exact check follows
        }
        const unsigned nsteps = rmax != 0.0 ? static_cast<unsigned>(budget - left) : 0u;
        double viol = 0.0;
#pragma unroll
        for (int i = 0; i < NB; ++i) viol = fmax(viol, h[i] > 0.0 ? fabs(g[i]) : fmax(-g[i], 0.0));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) viol = fmax(viol, __shfl_xor_sync(0xffffffffu, viol, o));
        const bool converged = viol <= thresh;
        if (lane == 0) {
            if (steps_total != nullptr) atomicAdd(steps_total, static_cast<unsigned long long>(nsteps));
            if (!converged) atomicAdd(counters + 1, 1ULL);
        }
        // an unfinished problem keeps its state for the clean-up pass and is marked by the sign of row 0
        // (solutions are >= 0): -h_0 for h_0 > 0, -DBL_MIN for h_0 == 0
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int t = lane + 32 * i;
            if (t < k) {
                double v = h[i] * scl[t];
                if (!converged && t == 0) v = v > 0.0 ? -v : -2.2250738585072014e-308;
                sol[t * ld_sol + j] = v;
            }
        }
    }
}

}  // namespace

int nnls_gram_batched(const double* gram, int64_t k, const float* rhs, int rhs_splits, int64_t rhs_stride,
                      int64_t ld_rhs, int64_t cols, double* sol, int64_t ld_sol, int max_outer, int* status,
                      cudaStream_t stream, unsigned long long* steps_total, unsigned long long* counter) {
    if (k == 0 || cols == 0) return 0;
    NFB_CHECK(k <= 512, "n_components > 512 is not supported by the NNLS kernel");
    const size_t smem_fast = (static_cast<size_t>(k) * (k + 1) / 2 + 3 * k) * sizeof(double);
    if (smem_fast <= 220 * 1024 && k <= 224 && counter != nullptr) {
        // counter[0]: column hand-out of the first kernel, [1]: columns the greedy kernel left unfinished,
        // [2]: chunk hand-out of the clean-up pass
        NFB_CUDA(cudaMemsetAsync(counter, 0, 3 * sizeof(unsigned long long), stream));
        // 16 warps per SM (128 registers, a few spilled values at k > 192) measured faster than 12 spill-free ones.
        // The knobs are read on every call so that one process can compare settings (tests/gpu_nnls_probe.py).
        const int warps_env = getenv("NFB_NNLS_WARPS") ? atoi(getenv("NFB_NNLS_WARPS")) : 0;
        const bool wide = warps_env ? warps_env >= 16 : true;
        const int kNnlsWarps = wide ? 16 : 12;
        const int sweeps0 = getenv("NFB_NNLS_SWEEPS") ? std::max(1, atoi(getenv("NFB_NNLS_SWEEPS"))) : kInitialSweeps;
        const double face_red = getenv("NFB_NNLS_FACE") ? atof(getenv("NFB_NNLS_FACE")) : kFaceReduction;
        const int scale_diag = getenv("NFB_NNLS_SCALE") ? atoi(getenv("NFB_NNLS_SCALE")) : 1;
        // greedy coordinate steps before the sweeps + CG loop takes over, in multiples of k (0: no greedy phase)
        const int greedy_mult = getenv("NFB_NNLS_GREEDY") ? std::max(0, atoi(getenv("NFB_NNLS_GREEDY"))) : kGreedyBudget;
        const int greedy_budget = scale_diag ? greedy_mult * static_cast<int>(k) : 0;   // the keys assume a unit diagonal
        const int greedy_batch = getenv("NFB_NNLS_BATCH") ? std::max(1, atoi(getenv("NFB_NNLS_BATCH"))) : kGreedyBatch;
        // NFB_NNLS_SPLIT=0: everything in the one 16-warp kernel (greedy phase + sweeps / CG), as before the
        // 32-warp greedy kernel existed
        const bool split = !(getenv("NFB_NNLS_SPLIT") && atoi(getenv("NFB_NNLS_SPLIT")) == 0);
        const size_t smem_greedy = greedy_smem_bytes(static_cast<int>(k));
        const unsigned long long* flagged = nullptr;
        unsigned long long* fast_counter = counter;
        if (split && greedy_budget > 0 && smem_greedy <= 226 * 1024) {
            const int gw = warps_env > 16 ? std::min(32, warps_env & ~3) : 32;
            GreedyLastOff last_off{};
            {
                const int nb = static_cast<int>((k + 31) / 32), rows_last = static_cast<int>(k) - 32 * (nb - 1);
                for (int c = 0; c < 8; ++c) last_off.v[c] = 8 * greedy_block_off(nb - 1, std::min(c, nb), nb - 1, rows_last * 33);
            }
            const int ggrid = static_cast<int>(std::min<int64_t>(ceil_div64(cols, gw), sm_count()));
#define NFB_NNLS_GREEDY_LAUNCH(NB)                                                                                \
    do {                                                                                                         \
        auto gk = gw >= 32 ? nnls_greedy_kernel<NB, 32> : (gw >= 24 ? nnls_greedy_kernel<NB, 24> : nnls_greedy_kernel<NB, 20>); \
        NFB_CUDA(cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize,                           \
                                      static_cast<int>(smem_greedy)));                                           \
        gk<<<ggrid, (gw >= 32 ? 32 : (gw >= 24 ? 24 : 20)) * 32, smem_greedy, stream>>>(                          \
            gram, static_cast<int>(k), rhs, rhs_splits, rhs_stride, ld_rhs, cols, sol, ld_sol, 1e-10, steps_total, \
            counter, greedy_budget, greedy_batch, last_off);                                                     \
    } while (0)
            switch ((k + 31) / 32) {
                case 1: NFB_NNLS_GREEDY_LAUNCH(1); break;
                case 2: NFB_NNLS_GREEDY_LAUNCH(2); break;
                case 3: NFB_NNLS_GREEDY_LAUNCH(3); break;
                case 4: NFB_NNLS_GREEDY_LAUNCH(4); break;
                case 5: NFB_NNLS_GREEDY_LAUNCH(5); break;
                case 6: NFB_NNLS_GREEDY_LAUNCH(6); break;
                default: NFB_NNLS_GREEDY_LAUNCH(7); break;
            }
#undef NFB_NNLS_GREEDY_LAUNCH
            NFB_CUDA(cudaGetLastError());
            flagged = counter + 1;
            fast_counter = counter + 2;
        }
        const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(cols, kNnlsWarps), sm_count()));
#define NFB_NNLS_FAST(KPL)                                                                                       \
    do {                                                                                                         \
        auto kernel = wide ? nnls_smem_kernel<KPL, 16> : nnls_smem_kernel<KPL, 12>;                              \
        NFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                      static_cast<int>(smem_fast)));                                             \
        kernel<<<grid, kNnlsWarps * 32, smem_fast, stream>>>(gram, static_cast<int>(k), rhs, rhs_splits,         \
                                                             rhs_stride, ld_rhs, cols, sol, ld_sol, max_outer,   \
                                                             1e-10, status, steps_total, fast_counter, sweeps0,  \
                                                             face_red, scale_diag, greedy_budget, greedy_batch,  \
                                                             flagged);                                           \
    } while (0)
        if (k <= 32) NFB_NNLS_FAST(1);
        else if (k <= 64) NFB_NNLS_FAST(2);
        else if (k <= 128) NFB_NNLS_FAST(4);
        else NFB_NNLS_FAST(7);
#undef NFB_NNLS_FAST
        NFB_CUDA(cudaGetLastError());
        return 0;
    }
    const int threads = 128;
    const int64_t warps_needed = cols;
    const int blocks = static_cast<int>(std::max<int64_t>(
        1, std::min<int64_t>(ceil_div64(warps_needed * 32, threads), static_cast<int64_t>(sm_count()) * 8)));
    const double tol = 1e-10;
#define NFB_NNLS_LAUNCH(KPL)                                                                                   \
    nnls_cd_kernel<KPL><<<blocks, threads, 0, stream>>>(gram, static_cast<int>(k), rhs, rhs_splits, rhs_stride, \
                                                        ld_rhs, cols, sol, ld_sol, max_outer, tol, status, steps_total)
    if (k <= 64) NFB_NNLS_LAUNCH(2);
    else if (k <= 128) NFB_NNLS_LAUNCH(4);
    else if (k <= 256) NFB_NNLS_LAUNCH(8);
    else NFB_NNLS_LAUNCH(16);
#undef NFB_NNLS_LAUNCH
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
