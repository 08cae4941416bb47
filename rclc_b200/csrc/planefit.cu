// planefit.cu — robust plane fit of the white board points on device (SURVEY.md §8f "next #1").
//
// Replaces pc_plane_fitting (include/opt/opt_plane_param.hpp:45-88): one residual per point with intensity >= 100,
//   r = |a x + b y + c z + d| / sqrt(a^2 + b^2 + c^2)            (PLANE_FITTING_COST, :24-43, evaluated through AutoDiff)
// each under its own HuberLoss(0.1), parameters (a, b, c, d) from (0.1, 0.1, 1, 1), ceres::Solve with DENSE_QR and
// otherwise default options (50 iterations, function_tolerance 1e-6). The reference allocates one residual block and one
// loss object per point; here an evaluation is ONE streaming pass (16 B per point) that reduces the robustified cost and
// the 4x4 normal equations J'J / J'r, and the last CTA of that pass runs Ceres' trust-region step (the same accept /
// reject / radius schedule as the grid fit, restated for 4 parameters). The step is solved by Cholesky on the normal
// equations instead of QR on the stacked Jacobian: the same minimiser of the same damped quadratic.
#include "ops_dev.cuh"

#include <cfloat>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace rclc {

// Bit-reproducible by construction. The sums over the points (cost, J'J, J'r) are exact, order-free fixed-point sums of terms
// that are each rounded once (common.cuh: Fix128), every other expression is written with `xd` (no FMA contraction) in one fixed
// association, and the step is the Cholesky solution of the 4x4 damped normal equations built from those sums. The CPU checker of
// the tests states the solve the same way — Ceres' own DENSE_QR (Eigen Householder on the stacked Jacobian) has no pinned
// operation order in the reference either — so plane, iteration count and every intermediate agree with it bit for bit, whatever
// the grid shape. That is what the downstream stages need: the projection onto this plane is rounded to float, and a last-bit
// difference there is enough to fork the grid fit's LM path (sensitivity of the chain).
constexpr int kPfThreads = 256;
constexpr int kPfMaxCtas = 592;       // 148 SMs x 4
constexpr int kPfSums = 15;           // cost, 10 J'J (upper triangle, row-major), 4 J'r
constexpr int kPfWords = 2 * kPfSums + 2;     // 128-bit sums, the point count, the "evaluation failed" flag

struct PlaneLm {
    // control
    int phase;                 // 1 evaluate x (initial), 2 evaluate candidate, 3 done
    int termination;           // 0 convergence, 1 no convergence, 2 failure
    int iteration, num_iterations, num_consecutive_invalid, reuse_diagonal, step_is_successful;
    unsigned int counter;      // CTAs of the current pass that have written their partials
    long long n_used;
    double eval_x[4];          // where the next pass evaluates
    double x[4], cand[4], best[4];
    double x_cost, x_norm, radius, decrease_factor, minimum_cost, initial_cost, min_iter_cost, model_cost_change, gmax;
    double A[10], g[4], scale[4], diag[4];
};

struct PlaneEval { double cost; double A[10]; double g[4]; bool valid; };

__device__ bool chol4_solve(const double* A10, const double* l, const double* b, double* y)
{
    // A (upper triangle, row-major: 00 01 02 03 11 12 13 22 23 33) + diag(l^2), Cholesky, solve A y = b
    xd M[4][4];
    int q = 0;
    for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { M[i][j] = xd{A10[q]}; M[j][i] = xd{A10[q]}; ++q; }
    for (int i = 0; i < 4; ++i) M[i][i] = M[i][i] + xd{l[i]} * xd{l[i]};
    xd L[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) L[i][j] = xd{0.0};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j <= i; ++j) {
            xd s = M[i][j];
            for (int k = 0; k < j; ++k) s = s - L[i][k] * L[j][k];
            if (i == j) { if (!(s.v > 0.0)) return false; L[i][i] = xsqrt(s); }
            else L[i][j] = s / L[j][j];
        }
    xd z[4], yy[4];
    for (int i = 0; i < 4; ++i) { xd s = xd{b[i]}; for (int k = 0; k < i; ++k) s = s - L[i][k] * z[k]; z[i] = s / L[i][i]; }
    for (int i = 3; i >= 0; --i) { xd s = z[i]; for (int k = i + 1; k < 4; ++k) s = s - L[k][i] * yy[k]; yy[i] = s / L[i][i]; }
    for (int i = 0; i < 4; ++i) y[i] = yy[i].v;
    return isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]) && isfinite(y[3]);
}

__device__ double norm4(const double* x)
{
    return xsqrt(((xd{x[0]} * xd{x[0]} + xd{x[1]} * xd{x[1]}) + xd{x[2]} * xd{x[2]}) + xd{x[3]} * xd{x[3]}).v;
}

__device__ void plane_take_eval(PlaneLm& s, const PlaneEval& e)
{
    s.x_cost = e.cost;
    for (int k = 0; k < 10; ++k) s.A[k] = e.A[k];
    double gm = 0;
    for (int k = 0; k < 4; ++k) { s.g[k] = e.g[k]; gm = fmax(gm, fabs((xd{s.x[k]} - (xd{s.x[k]} + xd{-e.g[k]})).v)); }
    s.gmax = gm;
}

// TrustRegionMinimizer as a resumable state machine (Ceres' published algorithm: SURVEY.md Appendix B).
__device__ void plane_lm_advance(PlaneLm& s, const PlaneEval& e, double ftol, int max_iters)
{
    const int diag_idx[4] = {0, 4, 7, 9};
    bool terminate = false;
    if (s.phase == 1) {
        if (!e.valid) { s.termination = 2; s.initial_cost = nan(""); s.min_iter_cost = nan(""); s.phase = 3; return; }
        plane_take_eval(s, e);
        for (int k = 0; k < 4; ++k) s.scale[k] = (xd{1.0} / (xd{1.0} + xsqrt(xd{e.A[diag_idx[k]]}))).v;       // jacobi scaling, once
        s.initial_cost = s.x_cost; s.min_iter_cost = s.x_cost; s.minimum_cost = DBL_MAX;
        s.step_is_successful = 1; s.iteration = 0; s.num_iterations = 0; s.radius = 1e4; s.decrease_factor = 2.0;
        s.reuse_diagonal = 0; s.num_consecutive_invalid = 0;
        s.x_norm = norm4(s.x);
    } else {
        const double cand_cost = e.valid ? e.cost : DBL_MAX;
        xd sn = xd{0.0};
        for (int k = 0; k < 4; ++k) sn = sn + (xd{s.x[k]} - xd{s.cand[k]}) * (xd{s.x[k]} - xd{s.cand[k]});
        if (xsqrt(sn).v <= (xd{1e-8} * (xd{s.x_norm} + xd{1e-8})).v) { s.termination = 0; terminate = true; }
        else if (fabs((xd{s.x_cost} - xd{cand_cost}).v) <= (xd{ftol} * xd{s.x_cost}).v) { s.termination = 0; terminate = true; }
        else {
            const double rel = (cand_cost >= DBL_MAX) ? -DBL_MAX : ((xd{s.x_cost} - xd{cand_cost}) / xd{s.model_cost_change}).v;
            if (rel > 1e-3) {
                for (int k = 0; k < 4; ++k) s.x[k] = s.cand[k];
                s.x_norm = norm4(s.x);
                if (!e.valid) { s.termination = 2; terminate = true; }
                else {
                    plane_take_eval(s, e);
                    s.step_is_successful = 1;
                    const xd q = xd{2.0} * xd{rel} - xd{1.0};
                    s.radius = fmin(1e16, (xd{s.radius} / xd{fmax(1.0 / 3.0, (xd{1.0} - (q * q) * q).v)}).v);
                    s.decrease_factor = 2.0;
                    s.reuse_diagonal = 0;
                    s.min_iter_cost = fmin(s.min_iter_cost, s.x_cost);
                }
            } else {
                s.min_iter_cost = fmin(s.min_iter_cost, cand_cost);
                s.radius = (xd{s.radius} / xd{s.decrease_factor}).v;
                s.decrease_factor *= 2.0;
                s.reuse_diagonal = 1;
                s.step_is_successful = 0;
            }
        }
    }
    while (!terminate) {
        if (s.step_is_successful && s.x_cost < s.minimum_cost) { s.minimum_cost = s.x_cost; for (int k = 0; k < 4; ++k) s.best[k] = s.x[k]; }
        s.num_iterations++;
        if (s.iteration >= max_iters) { s.termination = 1; break; }
        if (s.step_is_successful && s.gmax <= 1e-10) { s.termination = 0; break; }
        if (s.radius <= 1e-32) { s.termination = 0; break; }
        s.iteration++;
        s.step_is_successful = 0;
        double As[10], bs[4], l[4], y[4];
        {
            int q = 0;
            for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { As[q] = ((xd{s.A[q]} * xd{s.scale[i]}) * xd{s.scale[j]}).v; ++q; }
        }
        for (int k = 0; k < 4; ++k) bs[k] = (xd{s.g[k]} * xd{s.scale[k]}).v;
        if (!s.reuse_diagonal) for (int k = 0; k < 4; ++k) s.diag[k] = fmin(fmax(As[diag_idx[k]], 1e-6), 1e32);
        for (int k = 0; k < 4; ++k) l[k] = xsqrt(xd{s.diag[k]} / xd{s.radius}).v;
        bool valid = chol4_solve(As, l, bs, y);
        s.reuse_diagonal = 1;
        double step[4], mcc = 0;
        if (valid) {
            for (int k = 0; k < 4; ++k) step[k] = -y[k];
            double M[4][4]; int q = 0;
            for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { M[i][j] = As[q]; M[j][i] = As[q]; ++q; }
            xd sAs = xd{0.0}, sb = xd{0.0};
            for (int i = 0; i < 4; ++i) {
                xd r = xd{0.0};
                for (int j = 0; j < 4; ++j) r = r + xd{M[i][j]} * xd{step[j]};
                sAs = sAs + xd{step[i]} * r;
                sb = sb + xd{step[i]} * xd{bs[i]};
            }
            mcc = (-(sb + xd{0.5} * sAs)).v;               // -(J step)'(r + J step / 2)
            valid = mcc > 0.0;
        }
        if (!valid) {
            if (++s.num_consecutive_invalid >= 5) { s.termination = 2; break; }
            s.radius *= 0.5;
            s.min_iter_cost = fmin(s.min_iter_cost, s.x_cost);
            continue;
        }
        s.num_consecutive_invalid = 0;
        s.model_cost_change = mcc;
        for (int k = 0; k < 4; ++k) { s.cand[k] = (xd{s.x[k]} + xd{step[k]} * xd{s.scale[k]}).v; s.eval_x[k] = s.cand[k]; }
        s.phase = 2;
        return;
    }
    s.phase = 3;
}

// The terms of the points i = first, first + stride, ... of one evaluation at (a, b, c, d), added to the caller's exact sums.
__device__ __forceinline__ void plane_accumulate(const float4* __restrict__ pts, int64_t n, int64_t first, int64_t stride, const double* x, double thd,
                                                 double huber_a, Fix128 (&acc)[kPfSums], unsigned long long& cnt, bool& bad)
{
    const xd a{x[0]}, b{x[1]}, c{x[2]}, d{x[3]};
    const xd nn = (a * a + b * b) + c * c, nrm = xsqrt(nn), den = nn * nrm;
    const xd ha{huber_a}, hb = ha * ha;
    for (int64_t i = first; i < n; i += stride) {
        const float4 p = __ldg(pts + i);
        if ((double)p.w < thd) continue;                                   // opt_plane_param.hpp:56 (float < double)
        const xd X{(double)p.x}, Y{(double)p.y}, Z{(double)p.z};
        const xd e = ((a * X + b * Y) + c * Z) + d;
        const xd ae = xabs(e);
        xd r = ae / nrm;                                                    // :35-38
        const xd sg{(e.v < 0) ? -1.0 : 1.0};                                // ceres abs(Jet): f.a < 0 ? -f : f
        xd j[4] = {(sg * X) / nrm - (ae * a) / den, (sg * Y) / nrm - (ae * b) / den, (sg * Z) / nrm - (ae * c) / den, sg / nrm};
        // an evaluation fails on a non-finite residual or Jacobian; entries beyond 2^13 (products beyond 2^26: 2^21 of them would leave
        // the fixed-point range) count as that too — a plane whose normal has collapsed, never a fit of real coordinates
        if (!(fabs(r.v) < 8192.0 && fabs(j[0].v) < 8192.0 && fabs(j[1].v) < 8192.0 && fabs(j[2].v) < 8192.0 && fabs(j[3].v) < 8192.0)) { bad = true; continue; }
        // HuberLoss + Corrector (rho'' <= 0 branch)
        const xd sq = r * r;
        xd rho0, rho1;
        if (sq.v > hb.v) { const xd rr = xsqrt(sq); rho0 = (xd{2.0} * ha) * rr - hb; rho1 = xd{fmax(DBL_MIN, (ha / rr).v)}; }
        else { rho0 = sq; rho1 = xd{1.0}; }
        const xd w = xsqrt(rho1);
        r = r * w;
#pragma unroll
        for (int k = 0; k < 4; ++k) j[k] = j[k] * w;
        fix128_add_double(acc[0], (xd{0.5} * rho0).v);
        int q = 1;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = u; v < 4; ++v) fix128_add_double(acc[q++], (j[u] * j[v]).v);
#pragma unroll
        for (int u = 0; u < 4; ++u) fix128_add_double(acc[11 + u], (j[u] * r).v);
        ++cnt;
    }
}

// The CTA's share of the sums: warp shuffles, then the warps' partials through shared memory into out[kPfWords] (shared or global).
__device__ __forceinline__ void plane_cta_fold(Fix128 (&acc)[kPfSums], unsigned long long cnt, bool bad, unsigned long long (*s_red)[kPfWords], int* s_bad,
                                               unsigned long long* out)
{
    if (threadIdx.x == 0) *s_bad = 0;
    __syncthreads();
    if (bad) *s_bad = 1;
#pragma unroll
    for (int k = 0; k < kPfSums; ++k) fix128_warp_sum(acc[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < kPfSums; ++k) { s_red[threadIdx.x >> 5][2 * k] = (unsigned long long)acc[k].hi; s_red[threadIdx.x >> 5][2 * k + 1] = acc[k].lo; }
        s_red[threadIdx.x >> 5][2 * kPfSums] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < kPfSums) {
        Fix128 t;
        t.hi = 0; t.lo = 0;
        for (int w = 0; w < kPfThreads / 32; ++w) fix128_add(t, (long long)s_red[w][2 * threadIdx.x], s_red[w][2 * threadIdx.x + 1]);
        out[2 * threadIdx.x] = (unsigned long long)t.hi;
        out[2 * threadIdx.x + 1] = t.lo;
    } else if (threadIdx.x == kPfSums) {
        unsigned long long t = 0;
        for (int w = 0; w < kPfThreads / 32; ++w) t += s_red[w][2 * kPfSums];
        out[2 * kPfSums] = t;
        out[2 * kPfSums + 1] = *s_bad ? 1ull : 0ull;
    }
}

// One evaluation: cost + normal equations at st->eval_x over the points with intensity >= thd; the last CTA adds the per-CTA
// partial sums (integers: any order) and advances the LM state. The shape for large clouds: one launch per evaluation.
__global__ void __launch_bounds__(kPfThreads) k_plane_eval(const float4* __restrict__ pts, int64_t n, double thd, double huber_a, double ftol, int max_iters,
                                                           PlaneLm* __restrict__ st, unsigned long long* __restrict__ partials)
{
    if (st->phase == 3) return;
    Fix128 acc[kPfSums];
#pragma unroll
    for (int k = 0; k < kPfSums; ++k) { acc[k].hi = 0; acc[k].lo = 0; }
    unsigned long long cnt = 0;
    bool bad = false;
    const double x[4] = {st->eval_x[0], st->eval_x[1], st->eval_x[2], st->eval_x[3]};
    plane_accumulate(pts, n, blockIdx.x * (int64_t)kPfThreads + threadIdx.x, (int64_t)gridDim.x * kPfThreads, x, thd, huber_a, acc, cnt, bad);
    __shared__ unsigned long long s_red[kPfThreads / 32][kPfWords];
    __shared__ int s_bad, s_last;
    plane_cta_fold(acc, cnt, bad, s_red, &s_bad, partials + (size_t)blockIdx.x * kPfWords);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    __shared__ double s_tot[kPfSums];
    __shared__ unsigned long long s_cnt, s_anybad;
    if (threadIdx.x < kPfSums) {
        Fix128 t;
        t.hi = 0; t.lo = 0;
        for (unsigned cta = 0; cta < gridDim.x; ++cta)
            fix128_add(t, (long long)__ldcg(partials + (size_t)cta * kPfWords + 2 * threadIdx.x), __ldcg(partials + (size_t)cta * kPfWords + 2 * threadIdx.x + 1));
        s_tot[threadIdx.x] = fix128_to_double(t.hi, t.lo);
    } else if (threadIdx.x == kPfSums) {
        unsigned long long t = 0, bd = 0;
        for (unsigned cta = 0; cta < gridDim.x; ++cta) { t += __ldcg(partials + (size_t)cta * kPfWords + 2 * kPfSums); bd |= __ldcg(partials + (size_t)cta * kPfWords + 2 * kPfSums + 1); }
        s_cnt = t; s_anybad = bd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->counter = 0;
        PlaneEval e;
        e.cost = s_tot[0];
        for (int k = 0; k < 10; ++k) e.A[k] = s_tot[1 + k];
        for (int k = 0; k < 4; ++k) e.g[k] = s_tot[11 + k];
        e.valid = s_anybad == 0ull;
        st->n_used = (long long)s_cnt;
        PlaneLm s = *st;
        plane_lm_advance(s, e, ftol, max_iters);
        *st = s;
    }
}

// The shape for the clouds of the calibration chain (tens of thousands of points): the WHOLE solve in one launch on a thread-block
// cluster. Every CTA sums its share of an evaluation, the partial sums are exchanged through distributed shared memory, and
// every CTA adds them up (integers: the same total everywhere) and advances its own copy of the LM state — identical inputs,
// identical code, so no decision has to be published. Two cluster barriers per evaluation. The sums being exact, the solve
// returns the bits of the one-launch-per-evaluation shape.
__device__ void plane_init_state(PlaneLm& s);
constexpr int kPfCluster = 8;
constexpr int64_t kPfClusterMaxPts = 200000;      // above: one launch per evaluation over up to 592 CTAs
__global__ void __cluster_dims__(kPfCluster, 1, 1) __launch_bounds__(kPfThreads)
k_plane_fit_cluster(const float4* __restrict__ pts, int64_t n, double thd, double huber_a, double ftol, int max_iters, PlaneLm* __restrict__ st_out)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ unsigned long long s_red[kPfThreads / 32][kPfWords];
    __shared__ unsigned long long s_part[kPfWords];
    __shared__ double s_tot[kPfSums];
    __shared__ unsigned long long s_cnt, s_anybad;
    __shared__ int s_bad;
    __shared__ PlaneLm s_lm;
    if (threadIdx.x == 0) plane_init_state(s_lm);
    __syncthreads();
    while (s_lm.phase != 3) {
        Fix128 acc[kPfSums];
#pragma unroll
        for (int k = 0; k < kPfSums; ++k) { acc[k].hi = 0; acc[k].lo = 0; }
        unsigned long long cnt = 0;
        bool bad = false;
        const double x[4] = {s_lm.eval_x[0], s_lm.eval_x[1], s_lm.eval_x[2], s_lm.eval_x[3]};
        plane_accumulate(pts, n, rank * (int64_t)kPfThreads + threadIdx.x, (int64_t)kPfCluster * kPfThreads, x, thd, huber_a, acc, cnt, bad);
        plane_cta_fold(acc, cnt, bad, s_red, &s_bad, s_part);
        cluster.sync();                                                        // every CTA's partial sums are in its shared memory
        if (threadIdx.x < kPfSums) {
            Fix128 t;
            t.hi = 0; t.lo = 0;
            for (int r = 0; r < kPfCluster; ++r) {
                const unsigned long long* rp = cluster.map_shared_rank(s_part, r);
                fix128_add(t, (long long)rp[2 * threadIdx.x], rp[2 * threadIdx.x + 1]);
            }
            s_tot[threadIdx.x] = fix128_to_double(t.hi, t.lo);
        } else if (threadIdx.x == kPfSums) {
            unsigned long long t = 0, bd = 0;
            for (int r = 0; r < kPfCluster; ++r) { const unsigned long long* rp = cluster.map_shared_rank(s_part, r); t += rp[2 * kPfSums]; bd |= rp[2 * kPfSums + 1]; }
            s_cnt = t; s_anybad = bd;
        }
        cluster.sync();                                                        // all reads of the partial sums are done: the next evaluation may overwrite them
        if (threadIdx.x == 0) {
            PlaneEval e;
            e.cost = s_tot[0];
            for (int k = 0; k < 10; ++k) e.A[k] = s_tot[1 + k];
            for (int k = 0; k < 4; ++k) e.g[k] = s_tot[11 + k];
            e.valid = s_anybad == 0ull;
            s_lm.n_used = (long long)s_cnt;
            plane_lm_advance(s_lm, e, ftol, max_iters);
        }
        __syncthreads();
    }
    if (rank == 0 && threadIdx.x == 0) *st_out = s_lm;
}

__device__ void plane_init_state(PlaneLm& s)
{
    memset(&s, 0, sizeof(s));
    const double x0[4] = {0.1, 0.1, 1.0, 1.0};                                 // opt_plane_param.hpp:49
    for (int k = 0; k < 4; ++k) { s.x[k] = x0[k]; s.eval_x[k] = x0[k]; s.best[k] = x0[k]; }
    s.phase = 1;
    s.termination = 1;
}
__global__ void k_plane_init(PlaneLm* st)
{
    PlaneLm s;
    plane_init_state(s);
    *st = s;
}

int dev_plane_fit(rclc_ctx* ctx, const rclc_config* cfg, const float4* d_pts, int64_t n, double plane_out[4], int32_t* iterations, int64_t* n_used)
{
    const int ctas = (int)std::max<int64_t>(1, std::min<int64_t>((n + kPfThreads * 4 - 1) / (kPfThreads * 4), kPfMaxCtas));
    RCLC_TRY(ctx->d_tmp.reserve(sizeof(PlaneLm) + 256 + (size_t)kPfMaxCtas * kPfWords * 8));
    PlaneLm* d_st = ctx->d_tmp.as<PlaneLm>();
    unsigned long long* d_part = reinterpret_cast<unsigned long long*>(ctx->d_tmp.as<char>() + ((sizeof(PlaneLm) + 255) & ~(size_t)255));
    // Ceres defaults for this solve (opt_plane_param.hpp:75-77 sets only DENSE_QR): 50 iterations, function_tolerance 1e-6.
    const int max_iters = 50;
    const double ftol = 1e-6;
    RCLC_TRY(ctx->h_pin2.reserve(sizeof(PlaneLm)));
    PlaneLm* h = ctx->h_pin2.as<PlaneLm>();
    if (n <= kPfClusterMaxPts) {
        k_plane_fit_cluster<<<kPfCluster, kPfThreads, 0, ctx->stream>>>(d_pts, n, cfg->plane_reflactive_thd, cfg->plane_huber_a, ftol, max_iters, d_st);
        ctx->launches += 1;
        RCLC_CUDA(cudaGetLastError());
        RCLC_CUDA(cudaMemcpyAsync(h, d_st, sizeof(PlaneLm), cudaMemcpyDeviceToHost, ctx->stream));
        RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
        RCLC_CHECK(h->phase == 3, RCLC_ERR_NUMERIC, "plane-fit state machine did not terminate");
    } else {
        k_plane_init<<<1, 1, 0, ctx->stream>>>(d_st);
        ctx->launches += 1;
        const int max_launches = 2 * (max_iters + 8) + 8;
        int launched = 0;
        for (;;) {
            for (int r = 0; r < 8; ++r) {
                k_plane_eval<<<ctas, kPfThreads, 0, ctx->stream>>>(d_pts, n, cfg->plane_reflactive_thd, cfg->plane_huber_a, ftol, max_iters, d_st, d_part);
                ctx->launches += 1;
            }
            launched += 8;
            RCLC_CUDA(cudaGetLastError());
            RCLC_CUDA(cudaMemcpyAsync(h, d_st, sizeof(PlaneLm), cudaMemcpyDeviceToHost, ctx->stream));
            RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
            if (h->phase == 3) break;
            RCLC_CHECK(launched < max_launches, RCLC_ERR_NUMERIC, "plane-fit state machine did not terminate");
        }
    }
    // solver.cc: the user state is only overwritten when the solution is usable
    const double x0[4] = {0.1, 0.1, 1.0, 1.0};
    for (int k = 0; k < 4; ++k) plane_out[k] = h->termination == 2 ? x0[k] : h->best[k];
    if (iterations) *iterations = h->num_iterations;
    if (n_used) *n_used = h->n_used;
    return h->termination == 2 ? RCLC_ERR_NUMERIC : RCLC_OK;
}

} // namespace rclc

using namespace rclc;

extern "C" int rclc_plane_fit(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double plane_out[4],
                              int32_t* iterations, int64_t* n_used)
{
    RCLC_CHECK(ctx && cfg && pts && plane_out && n > 0, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    RCLC_TRY(ctx->d_pts.reserve((size_t)n * 16));
    RCLC_TRY(upload_cloud_f4(ctx, pts, stride, ioff, n, ctx->d_pts.as<float4>()));
    return dev_plane_fit(ctx, cfg, ctx->d_pts.as<float4>(), n, plane_out, iterations, n_used);
}
