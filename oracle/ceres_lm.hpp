// oracle/ceres_lm.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product.
//
// Restatement of the Ceres Solver trust-region Levenberg-Marquardt minimiser as the reference
// configures it (include/opt/opt_grid_fitting.h:96-103, include/opt/opt_reprj_calib.h:51-61,
// include/opt/opt_plane_param.hpp:75-82, include/opt/opt_line_fitting.h:89-97).
//
// Ceres is a third-party dependency that is NOT vendored in /root/reference and NOT pinned
// (CMakeLists.txt:12 `find_package(Ceres REQUIRED)`), so this follows the published algorithm of
// Ceres 1.14 / 2.x `TrustRegionMinimizer` + `LevenbergMarquardtStrategy` + `TrustRegionStepEvaluator`
// + `Corrector` + `HuberLoss`.  PARITY UNPINNED against the reference: it holds no
// test or golden vector for any of this. What there is: the minimiser reproduces the two solver logs the Ceres tutorial
// prints, digit for digit (tests/test_oracle_ceres_logs.py), and reaches scipy's minimum on the plane fit and the extrinsic
// solve (tests/test_oracle_scipy_pins.py).
//
// All residual blocks in the reference have exactly ONE residual, which this restatement exploits:
// a problem is `m` scalar residual blocks over one dense parameter vector.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

namespace orc {

enum LinearSolverType { DENSE_QR = 0, DENSE_NORMAL_CHOLESKY = 1 };
enum TerminationType { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2 };

struct LMOptions {
    LinearSolverType linear_solver = DENSE_QR;
    int max_num_iterations = 50;                  // Ceres default
    double function_tolerance = 1e-6;             // Ceres default
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4;
    double max_trust_region_radius = 1e16;
    double min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3;
    double min_lm_diagonal = 1e-6;
    double max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    bool jacobi_scaling = true;
    // The radius update takes pow(2 rho - 1, 3) from libm in Ceres. true: the correctly rounded cube instead (rn::cube) — what a
    // correctly rounding libm returns (glibc <= 2.27 did; glibc 2.39 here differs from it by one ulp in 0.08 % of calls), and a
    // value a GPU can reproduce bit for bit. Set by the extrinsic solve only (rclc_oracle.cpp orc_pose_refine).
    bool correctly_rounded_libm = false;
};

// Correctly rounded stand-ins for the three libm calls of the extrinsic solve (pow(q, 3), sin, cos), in double-double arithmetic:
// the values libm is specified to approximate, independent of the libm at hand. tests/test_oracle_restatement.py counts how often
// this glibc differs from them (by one ulp, in 1e-3 .. 1e-6 of calls) and what that does to the solve.
namespace rn {
struct dd { double h, l; };
inline dd norm(double s, double e) { const double h = s + e; return {h, e - (h - s)}; }
inline dd mul(dd a, dd b) { const double p = a.h * b.h; double e = std::fma(a.h, b.h, -p); e = e + (a.h * b.l + a.l * b.h); return norm(p, e); }
inline dd add(dd a, dd b) { const double s = a.h + b.h, v = s - a.h; double e = (a.h - (s - v)) + (b.h - v); e = e + (a.l + b.l); return norm(s, e); }
inline double cube(double q)
{
    const double q2h = q * q, q2l = std::fma(q, q, -q2h), ph = q2h * q, pl = std::fma(q2h, q, -ph);
    return ph + (pl + q2l * q);
}
// (-1)^k / (2k+1)! and (-1)^k / (2k)! as double-doubles
static const dd kSin[15] = {
    {0x1.0000000000000p+0, 0x0.0p+0}, {-0x1.5555555555555p-3, -0x1.5555555555555p-57}, {0x1.1111111111111p-7, 0x1.1111111111111p-63},
    {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73}, {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80},
    {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87}, {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
    {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120}, {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130},
    {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139}, {-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157}};
static const dd kCos[15] = {
    {0x1.0000000000000p+0, 0x0.0p+0}, {-0x1.0000000000000p-1, 0x0.0p+0}, {0x1.5555555555555p-5, 0x1.5555555555555p-59},
    {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65}, {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76}, {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76},
    {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83}, {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
    {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120}, {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124},
    {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135}, {-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153}};
// sin and cos of 0 <= x <= 0.5 by their Taylor series in double-double (2^-100 and better), rounded once; libm beyond
inline void sincos(double x, double* sn, double* cs)
{
    if (!(x <= 0.5)) { *sn = std::sin(x); *cs = std::cos(x); return; }
    const double uh = x * x;
    const dd u = {uh, std::fma(x, x, -uh)};
    const int n = uh < 0x1p-36 ? 3 : (uh < 0x1p-20 ? 5 : (uh < 0x1p-8 ? 9 : 14));    // terms to 2^-110 of the sum
    dd S = kSin[n], Cc = kCos[n];
    for (int k = n - 1; k >= 0; --k) { S = add(mul(S, u), kSin[k]); Cc = add(mul(Cc, u), kCos[k]); }
    const dd xs = mul(S, dd{x, 0.0});
    *sn = xs.h + xs.l;
    *cs = Cc.h + Cc.l;
}
} // namespace rn

struct LMSummary {
    double initial_cost = 0, final_cost = 0;
    int termination = NO_CONVERGENCE;
    int num_iterations = 0;          // iteration summaries pushed (incl. iteration 0)
    int num_successful_steps = 0;
    int num_cost_evals = 0;          // distinct points at which residuals were evaluated
    int num_jac_evals = 0;           // evaluations with Jacobians
    // One entry per Ceres IterationSummary (iteration 0 included): its `cost` and `trust_region_radius`
    // columns, so that a run can be laid beside a published Ceres log (tests/test_oracle_ceres_logs.py).
    std::vector<double> trace_cost, trace_radius;
};

// HuberLoss::Evaluate (ceres loss_function.cc).
inline void huber_rho(double a, double s, double rho[3])
{
    const double b = a * a;
    if (s > b) {
        const double r = std::sqrt(s);
        rho[0] = 2.0 * a * r - b;
        rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
        rho[2] = -rho[1] / (2.0 * s);
    } else {
        rho[0] = s;
        rho[1] = 1.0;
        rho[2] = 0.0;
    }
}

// ResidualBlock::Evaluate tail for a 1-residual block with loss: returns 0.5*rho(r^2) and applies the
// Corrector to r and its Jacobian row (corrector.cc). With Huber, rho''<=0 always, so the
// "common case" branch is the only one taken; the general branch is kept for fidelity.
inline double apply_loss_1d(double huber_a, double* r, double* jrow, int ncols)
{
    const double sq_norm = (*r) * (*r);
    if (huber_a <= 0.0) return 0.5 * sq_norm;          // loss_function == NULL
    double rho[3];
    huber_rho(huber_a, sq_norm, rho);
    const double sqrt_rho1 = std::sqrt(rho[1]);
    double residual_scaling, alpha_sq_norm;
    if (sq_norm == 0.0 || rho[2] <= 0.0) {
        residual_scaling = sqrt_rho1;
        alpha_sq_norm = 0.0;
    } else {
        const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / sq_norm;
    }
    if (jrow) {
        if (alpha_sq_norm == 0.0) {
            for (int c = 0; c < ncols; ++c) jrow[c] *= sqrt_rho1;
        } else {
            // J = sqrt(rho1) * (J - alpha/s * r r' J); for 1 residual: r r' J = r^2 J
            for (int c = 0; c < ncols; ++c)
                jrow[c] = sqrt_rho1 * (jrow[c] - alpha_sq_norm * (*r) * ((*r) * jrow[c]));
        }
    }
    *r *= residual_scaling;
    return 0.5 * rho[0];
}

// A dense problem: m residuals, n ambient parameters, nl tangent (local) parameters.
struct LMProblem {
    int m = 0, n = 0, nl = 0;
    // Evaluate at x. residuals (m) may be null only together with jac; jac is row-major m x nl in
    // the LOCAL tangent space, already loss-corrected like ProgramEvaluator does. Returns false on
    // evaluation failure (NaN/Inf — ResidualBlock::Evaluate's IsEvaluationValid).
    std::function<bool(const double* x, double* cost, double* residuals, double* jac)> evaluate;
    // x_plus = Plus(x, delta). Null => Euclidean.
    std::function<void(const double* x, const double* delta, double* x_plus)> plus;
};

namespace detail {

// min ||A y - b||, A is (rows x n) row-major, via Householder QR without pivoting (what
// Eigen::HouseholderQR does for ceres DenseQRSolver).  A and b are overwritten.
inline bool householder_solve(std::vector<double>& A, std::vector<double>& b, int rows, int n, double* y)
{
    for (int k = 0; k < n; ++k) {
        double norm2 = 0;
        for (int i = k; i < rows; ++i) norm2 += A[i * n + k] * A[i * n + k];
        const double alpha = A[k * n + k];
        double beta = std::sqrt(norm2);
        if (beta == 0.0) return false;
        if (alpha > 0) beta = -beta;
        // v = x - beta e1, stored in column k (v_k = alpha - beta)
        const double vk = alpha - beta;
        const double vtv = norm2 - alpha * alpha + vk * vk;
        if (vtv == 0.0) return false;
        for (int c = k + 1; c < n; ++c) {
            double dot = vk * A[k * n + c];
            for (int i = k + 1; i < rows; ++i) dot += A[i * n + k] * A[i * n + c];
            const double f = 2.0 * dot / vtv;
            A[k * n + c] -= f * vk;
            for (int i = k + 1; i < rows; ++i) A[i * n + c] -= f * A[i * n + k];
        }
        {
            double dot = vk * b[k];
            for (int i = k + 1; i < rows; ++i) dot += A[i * n + k] * b[i];
            const double f = 2.0 * dot / vtv;
            b[k] -= f * vk;
            for (int i = k + 1; i < rows; ++i) b[i] -= f * A[i * n + k];
        }
        A[k * n + k] = beta;
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < n; ++c) s -= A[k * n + c] * y[c];
        if (A[k * n + k] == 0.0) return false;
        y[k] = s / A[k * n + k];
    }
    return true;
}

// Solve (J'J + D^2) y = J'r by LLT (ceres DenseNormalCholeskySolver, Eigen backend).
inline bool normal_cholesky_solve(const double* J, const double* r, const double* D, int m, int n, double* y)
{
    std::vector<double> L(n * n, 0.0), g(n, 0.0);
    for (int i = 0; i < m; ++i)
        for (int a = 0; a < n; ++a) {
            const double ja = J[i * n + a];
            g[a] += ja * r[i];
            for (int b2 = 0; b2 <= a; ++b2) L[a * n + b2] += ja * J[i * n + b2];
        }
    for (int a = 0; a < n; ++a) L[a * n + a] += D[a] * D[a];
    for (int j = 0; j < n; ++j) {
        double d = L[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) return false;                       // Eigen LLT: NumericalIssue
        d = std::sqrt(d);
        L[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = L[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) {
        double s = g[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
        y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * y[k];
        y[i] = s / L[i * n + i];
    }
    return true;
}

inline bool all_finite(const double* v, int n)
{
    for (int i = 0; i < n; ++i)
        if (!std::isfinite(v[i])) return false;
    return true;
}

} // namespace detail

// TrustRegionMinimizer::Minimize. `x` is in/out (Ceres writes back the best iterate).
inline LMSummary lm_minimize(const LMProblem& P, const LMOptions& opt, double* x_io)
{
    const int m = P.m, n = P.n, nl = P.nl;
    LMSummary sum;
    std::vector<double> x(x_io, x_io + n), cand(n), best(x_io, x_io + n);
    std::vector<double> res(m), jac(m * nl), grad(nl), scale(nl, 1.0), diag(nl), lmdiag(nl);
    std::vector<double> step(nl), delta(nl), model(m), tmpx(n), negg(nl);
    auto plus = [&](const double* a, const double* d, double* out) {
        if (P.plus) P.plus(a, d, out);
        else for (int i = 0; i < n; ++i) out[i] = a[i] + d[i];
    };
    auto norm = [](const std::vector<double>& v) { double s = 0; for (double e : v) s += e * e; return std::sqrt(s); };

    double x_cost = 0, x_norm = norm(x);
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    int num_consecutive_invalid = 0;
    double gmax = 0;

    // EvaluateGradientAndJacobian()
    auto eval_grad_jac = [&](bool first, bool new_point) -> bool {
        if (new_point) sum.num_cost_evals++;
        sum.num_jac_evals++;
        if (!P.evaluate(x.data(), &x_cost, res.data(), jac.data())) return false;
        for (int c = 0; c < nl; ++c) grad[c] = 0;
        for (int i = 0; i < m; ++i)
            for (int c = 0; c < nl; ++c) grad[c] += jac[i * nl + c] * res[i];
        if (opt.jacobi_scaling) {
            if (first) {
                for (int c = 0; c < nl; ++c) {
                    double s = 0;
                    for (int i = 0; i < m; ++i) s += jac[i * nl + c] * jac[i * nl + c];
                    scale[c] = 1.0 / (1.0 + std::sqrt(s));
                }
            }
            for (int i = 0; i < m; ++i)
                for (int c = 0; c < nl; ++c) jac[i * nl + c] *= scale[c];
        }
        for (int c = 0; c < nl; ++c) negg[c] = -grad[c];
        plus(x.data(), negg.data(), tmpx.data());
        gmax = 0;
        for (int i = 0; i < n; ++i) gmax = std::max(gmax, std::fabs(x[i] - tmpx[i]));
        return true;
    };

    // IterationZero()
    if (!eval_grad_jac(true, true)) {
        sum.termination = FAILURE;
        sum.initial_cost = sum.final_cost = std::numeric_limits<double>::quiet_NaN();
        return sum;
    }
    sum.initial_cost = x_cost;
    sum.trace_cost.push_back(x_cost);
    sum.trace_radius.push_back(radius);
    double min_iter_cost = x_cost;          // SetSummaryFinalCost: min over iteration summaries
    double minimum_cost = std::numeric_limits<double>::max();
    bool step_is_successful = true;         // iteration 0 is marked successful
    int iteration = 0;

    for (;;) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue()
        if (step_is_successful) {
            sum.num_successful_steps++;
            if (x_cost < minimum_cost) { minimum_cost = x_cost; best = x; }
        }
        sum.num_iterations++;
        if (iteration >= opt.max_num_iterations) { sum.termination = NO_CONVERGENCE; break; }
        if (step_is_successful && gmax <= opt.gradient_tolerance) { sum.termination = CONVERGENCE; break; }
        if (radius <= opt.min_trust_region_radius) { sum.termination = CONVERGENCE; break; }

        iteration++;
        step_is_successful = false;

        // ComputeTrustRegionStep(): LevenbergMarquardtStrategy::ComputeStep
        if (!reuse_diagonal) {
            for (int c = 0; c < nl; ++c) {
                double s = 0;
                for (int i = 0; i < m; ++i) s += jac[i * nl + c] * jac[i * nl + c];
                diag[c] = std::min(std::max(s, opt.min_lm_diagonal), opt.max_lm_diagonal);
            }
        }
        for (int c = 0; c < nl; ++c) lmdiag[c] = std::sqrt(diag[c] / radius);
        bool solved;
        if (opt.linear_solver == DENSE_QR) {
            std::vector<double> A((m + nl) * nl, 0.0), b(m + nl, 0.0);
            std::memcpy(A.data(), jac.data(), sizeof(double) * m * nl);
            for (int c = 0; c < nl; ++c) A[(m + c) * nl + c] = lmdiag[c];
            std::memcpy(b.data(), res.data(), sizeof(double) * m);
            solved = detail::householder_solve(A, b, m + nl, nl, step.data());
        } else {
            solved = detail::normal_cholesky_solve(jac.data(), res.data(), lmdiag.data(), m, nl, step.data());
        }
        if (solved && !detail::all_finite(step.data(), nl)) solved = false;
        reuse_diagonal = true;
        bool step_is_valid = false;
        double model_cost_change = 0;
        if (solved) {
            for (int c = 0; c < nl; ++c) step[c] = -step[c];
            for (int i = 0; i < m; ++i) {
                double s = 0;
                for (int c = 0; c < nl; ++c) s += jac[i * nl + c] * step[c];
                model[i] = s;
            }
            double dot = 0;
            for (int i = 0; i < m; ++i) dot += model[i] * (res[i] + model[i] / 2.0);
            model_cost_change = -dot;
            step_is_valid = model_cost_change > 0.0;
        }
        if (!step_is_valid) {
            // HandleInvalidStep()
            if (++num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) {
                sum.termination = FAILURE;
                break;
            }
            radius *= 0.5;                      // StepIsInvalid
            reuse_diagonal = true;
            min_iter_cost = std::min(min_iter_cost, x_cost);
            sum.trace_cost.push_back(x_cost);
            sum.trace_radius.push_back(radius);
            continue;
        }
        for (int c = 0; c < nl; ++c) delta[c] = step[c] * scale[c];
        num_consecutive_invalid = 0;

        // ComputeCandidatePointAndEvaluateCost()
        plus(x.data(), delta.data(), cand.data());
        double cand_cost;
        sum.num_cost_evals++;
        if (!P.evaluate(cand.data(), &cand_cost, nullptr, nullptr)) cand_cost = std::numeric_limits<double>::max();

        // ParameterToleranceReached()
        {
            double sn = 0;
            for (int i = 0; i < n; ++i) sn += (x[i] - cand[i]) * (x[i] - cand[i]);
            const double tol = opt.parameter_tolerance * (x_norm + opt.parameter_tolerance);
            if (std::sqrt(sn) <= tol) { sum.termination = CONVERGENCE; break; }
        }
        // FunctionToleranceReached()
        {
            const double cost_change = x_cost - cand_cost;
            if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { sum.termination = CONVERGENCE; break; }
        }
        // IsStepSuccessful(): TrustRegionStepEvaluator::StepQuality (monotonic)
        double rel;
        if (cand_cost >= std::numeric_limits<double>::max()) rel = std::numeric_limits<double>::lowest();
        else rel = (x_cost - cand_cost) / model_cost_change;
        if (rel > opt.min_relative_decrease) {
            // HandleSuccessfulStep()
            x = cand;
            x_norm = norm(x);
            if (!eval_grad_jac(false, false)) { sum.termination = FAILURE; break; }
            step_is_successful = true;
            radius = radius / std::max(1.0 / 3.0, 1.0 - (opt.correctly_rounded_libm ? rn::cube(2.0 * rel - 1.0) : std::pow(2.0 * rel - 1.0, 3)));
            radius = std::min(opt.max_trust_region_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
            min_iter_cost = std::min(min_iter_cost, x_cost);
            sum.trace_cost.push_back(x_cost);
            sum.trace_radius.push_back(radius);
        } else {
            min_iter_cost = std::min(min_iter_cost, cand_cost);   // iteration_summary_.cost = candidate_cost_
            radius = radius / decrease_factor;                    // StepRejected
            decrease_factor *= 2.0;
            reuse_diagonal = true;
            sum.trace_cost.push_back(cand_cost);
            sum.trace_radius.push_back(radius);
        }
    }
    // Final bookkeeping for early `return`s inside the loop: the last successful iterate may not
    // have gone through Finalize (cannot happen: successful steps always loop back), so `best` holds.
    sum.final_cost = std::min(sum.initial_cost, min_iter_cost);
    // solver.cc Minimize(): the user state is only overwritten when Summary::IsSolutionUsable().
    if (sum.termination != FAILURE)
        for (int i = 0; i < n; ++i) x_io[i] = best[i];
    return sum;
}

} // namespace orc
