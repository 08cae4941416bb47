// boardframe.cu — the board frame from the black-block centroids: host code (24 points).
//
// Replaces get_ref_pt (include/utils.h:160-202), calc_laser_coor_v1 (include/pc_process.h:373-462) and its Ceres problem
// line_fitting / LINE_FITTING_COST (include/opt/opt_line_fitting.h:17-108): the link between eular_cluster_iter and
// opt_grid_fitting inside est_board_corner (pc_process.h:583-621). The line fit — three parameters, one Huber(0.1) residual
// per board row, ceres::Solve with DENSE_QR and default options — runs through Ceres' trust-region schedule (Jacobi scaling once,
// LM diagonal clamp, Householder QR on the stacked Jacobian, accept / reject / radius rule, the three stop tests), with the
// analytic Jacobian of sum |d x a/|a|| in place of Ceres' autodiff.
#include "common.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <functional>

namespace rclc {
namespace {

struct V3 { double x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 unit(V3 a) { const double l = norm(a); return {a.x / l, a.y / l, a.z / l}; }

constexpr int kMaxRows = 32;                                  // board_height residuals (one per board row)

// One evaluation of the line-fit problem: robustified cost, and (when asked) the loss-corrected residuals and Jacobian rows.
struct Eval3 { double cost; double r[kMaxRows]; double J[kMaxRows][3]; bool valid; };

// min || A y - b || for the (rows x 3) stacked system by Householder reflections, column by column, no pivoting — what Ceres'
// DENSE_QR (Eigen::HouseholderQR) computes. A and b are overwritten.
bool householder3(double (*A)[3], double* b, int rows, double* y)
{
    for (int k = 0; k < 3; ++k) {
        double norm2 = 0;
        for (int i = k; i < rows; ++i) norm2 += A[i][k] * A[i][k];
        const double alpha = A[k][k];
        double beta = std::sqrt(norm2);
        if (beta == 0.0) return false;
        if (alpha > 0) beta = -beta;
        const double vk = alpha - beta;                           // v = x - beta e1
        const double vtv = norm2 - alpha * alpha + vk * vk;
        if (vtv == 0.0) return false;
        for (int c = k + 1; c < 3; ++c) {
            double dot = vk * A[k][c];
            for (int i = k + 1; i < rows; ++i) dot += A[i][k] * A[i][c];
            const double f = 2.0 * dot / vtv;
            A[k][c] -= f * vk;
            for (int i = k + 1; i < rows; ++i) A[i][c] -= f * A[i][k];
        }
        double dot = vk * b[k];
        for (int i = k + 1; i < rows; ++i) dot += A[i][k] * b[i];
        const double f = 2.0 * dot / vtv;
        b[k] -= f * vk;
        for (int i = k + 1; i < rows; ++i) b[i] -= f * A[i][k];
        A[k][k] = beta;
    }
    for (int k = 2; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < 3; ++c) s -= A[k][c] * y[c];
        if (A[k][k] == 0.0) return false;
        y[k] = s / A[k][k];
    }
    return std::isfinite(y[0]) && std::isfinite(y[1]) && std::isfinite(y[2]);
}

// Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy with DENSE_QR and default options (50 iterations, function_tolerance
// 1e-6, gradient_tolerance 1e-10, parameter_tolerance 1e-8, initial radius 1e4, Jacobi scaling) for m residuals over three
// parameters. The problem is scale-invariant in its parameters (rank 2), so WHICH minimiser of the damped model is taken matters
// to the fourth digit of the answer: the step is the QR solution of the stacked system, like Ceres', not a Cholesky solve of the
// normal equations. x is updated with the best iterate unless the solve fails.
void lm3(const std::function<void(const double*, bool, Eval3&)>& eval, int m, double* x)
{
    Eval3 e, ec;
    e.cost = 0; e.valid = false;
    double grad[3], scale[3] = {1, 1, 1}, diag[3] = {0, 0, 0}, best[3] = {x[0], x[1], x[2]};
    double x_cost = 0, gmax = 0;
    auto take = [&](bool first) {                                 // gradient, Jacobi scaling, max-norm of the projected gradient
        x_cost = e.cost;
        for (int c = 0; c < 3; ++c) grad[c] = 0;
        for (int i = 0; i < m; ++i) for (int c = 0; c < 3; ++c) grad[c] += e.J[i][c] * e.r[i];
        if (first)
            for (int c = 0; c < 3; ++c) { double s = 0; for (int i = 0; i < m; ++i) s += e.J[i][c] * e.J[i][c]; scale[c] = 1.0 / (1.0 + std::sqrt(s)); }
        for (int i = 0; i < m; ++i) for (int c = 0; c < 3; ++c) e.J[i][c] *= scale[c];
        gmax = 0;
        for (int k = 0; k < 3; ++k) gmax = std::max(gmax, std::fabs(x[k] - (x[k] + (-grad[k]))));
    };
    eval(x, true, e);
    if (!e.valid) return;
    take(true);
    auto norm3 = [](const double* v) { double s = 0; for (int k = 0; k < 3; ++k) s += v[k] * v[k]; return std::sqrt(s); };
    double radius = 1e4, dec = 2.0, xn = norm3(x), minimum_cost = DBL_MAX;
    bool reuse = false, success = true, failed = false;
    int invalid = 0;
    for (int it = 0;;) {
        if (success && x_cost < minimum_cost) { minimum_cost = x_cost; for (int k = 0; k < 3; ++k) best[k] = x[k]; }
        if (it >= 50) break;
        if (success && gmax <= 1e-10) break;
        if (radius <= 1e-32) break;
        ++it;
        success = false;
        if (!reuse)
            for (int c = 0; c < 3; ++c) { double s = 0; for (int i = 0; i < m; ++i) s += e.J[i][c] * e.J[i][c]; diag[c] = std::min(std::max(s, 1e-6), 1e32); }
        double A[kMaxRows + 3][3], b[kMaxRows + 3], step[3];
        for (int i = 0; i < m; ++i) { for (int c = 0; c < 3; ++c) A[i][c] = e.J[i][c]; b[i] = e.r[i]; }
        for (int c = 0; c < 3; ++c) { for (int q = 0; q < 3; ++q) A[m + c][q] = 0.0; A[m + c][c] = std::sqrt(diag[c] / radius); b[m + c] = 0.0; }
        bool valid = householder3(A, b, m + 3, step);
        reuse = true;
        double mcc = 0;
        if (valid) {
            for (int c = 0; c < 3; ++c) step[c] = -step[c];
            double dot = 0;
            for (int i = 0; i < m; ++i) {
                double s = 0;
                for (int c = 0; c < 3; ++c) s += e.J[i][c] * step[c];
                dot += s * (e.r[i] + s / 2.0);                    // model_cost_change = -(J step)'(r + J step / 2)
            }
            mcc = -dot;
            valid = mcc > 0.0;
        }
        if (!valid) { if (++invalid >= 5) { failed = true; break; } radius *= 0.5; continue; }
        invalid = 0;
        double cand[3];
        for (int k = 0; k < 3; ++k) cand[k] = x[k] + step[k] * scale[k];
        eval(cand, false, ec);
        const double cc = ec.valid ? ec.cost : DBL_MAX;
        double sn = 0;
        for (int k = 0; k < 3; ++k) sn += (x[k] - cand[k]) * (x[k] - cand[k]);
        if (std::sqrt(sn) <= 1e-8 * (xn + 1e-8)) break;
        if (std::fabs(x_cost - cc) <= 1e-6 * x_cost) break;
        const double rel = cc >= DBL_MAX ? -DBL_MAX : (x_cost - cc) / mcc;
        if (rel > 1e-3) {
            for (int k = 0; k < 3; ++k) x[k] = cand[k];
            xn = norm3(x);
            eval(x, true, e);
            if (!e.valid) { failed = true; break; }
            take(false);
            success = true;
            radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3)));
            dec = 2.0; reuse = false;
        } else { radius = radius / dec; dec *= 2.0; reuse = true; }
    }
    if (!failed) for (int k = 0; k < 3; ++k) x[k] = best[k];
}

} // namespace
} // namespace rclc

using namespace rclc;

extern "C" int rclc_calc_laser_coor(const rclc_config* cfg, const double* cen, int32_t n, const double plane[4], double Tbl[16])
{
    RCLC_CHECK(cfg && cen && plane && Tbl && n >= 4, RCLC_ERR_INVALID, "bad argument");
    const int per_row = cfg->board_width / 2, n_rows = cfg->board_height, origin_id = cfg->board_height / 2;
    RCLC_CHECK(per_row >= 1 && n_rows >= 1 && per_row * n_rows <= n, RCLC_ERR_INVALID, "fewer block centroids than board_width * board_height / 2");
    auto at = [&](int i) { return V3{cen[3 * i], cen[3 * i + 1], cen[3 * i + 2]}; };
    // get_ref_pt (utils.h:160-202): the farthest centroid in the quadrant x > centre, y <= centre
    V3 c{0, 0, 0};
    for (int i = 0; i < n; ++i) { c.x += cen[3 * i]; c.y += cen[3 * i + 1]; c.z += cen[3 * i + 2]; }
    c.x /= n; c.y /= n; c.z /= n;
    double far_d[4] = {0, 0, 0, 0};
    int far_i[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const V3 p = at(i);
        const double d = norm(sub(p, c));
        const int axis = p.x > c.x ? (p.y > c.y ? 0 : 1) : (p.y > c.y ? 2 : 3);
        if (d > far_d[axis]) { far_d[axis] = d; far_i[axis] = i; }
    }
    const int ref_idx = far_i[1];
    const V3 ref0 = at(ref_idx);
    // calc_laser_coor_v1 (pc_process.h:383-417): first guess of the row direction, then all centroids ordered by their distance
    // from the reference row
    std::vector<std::pair<double, int>> dl;
    for (int i = 0; i < n; ++i) if (i != ref_idx) dl.push_back({norm(sub(at(i), ref0)), i});
    std::sort(dl.begin(), dl.end());
    RCLC_CHECK(dl.size() >= 3, RCLC_ERR_INVALID, "too few centroids");
    const V3 ref1 = at(dl[1].second), ref2 = at(dl[2].second);
    const V3 xt = unit((ref1.x < ref2.x && ref1.y < ref2.y) ? sub(ref0, ref1) : sub(ref0, ref2));
    std::vector<std::pair<double, int>> rows;
    for (int i = 0; i < n; ++i) rows.push_back({norm(cross(sub(ref0, at(i)), xt)), i});
    std::sort(rows.begin(), rows.end());
    // line_fitting (opt_line_fitting.h:51-108)
    std::vector<std::vector<V3>> rp((size_t)n_rows);
    std::vector<V3> rc((size_t)n_rows);
    V3 origin_centroid{0, 0, 0};
    int cnt = 0;
    for (int r = 0; r < n_rows; ++r) {
        V3 cc{0, 0, 0};
        for (int k = 0; k < per_row; ++k) { const V3 p = at(rows[(size_t)cnt++].second); rp[(size_t)r].push_back(p); cc.x += p.x; cc.y += p.y; cc.z += p.z; }
        cc.x /= per_row; cc.y /= per_row; cc.z /= per_row;
        rc[(size_t)r] = cc;
        if (r == origin_id) origin_centroid = cc;
    }
    double abc[3] = {xt.x, xt.y, xt.z};
    RCLC_CHECK(n_rows <= kMaxRows, RCLC_ERR_INVALID, "board_height too large");
    lm3([&](const double* x, bool want_jac, Eval3& e) {
        e.cost = 0; e.valid = true;
        const double na = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        const V3 u{x[0] / na, x[1] / na, x[2] / na};
        for (int r = 0; r < n_rows; ++r) {
            double rr = 0, g[3] = {0, 0, 0};
            for (const V3& p : rp[(size_t)r]) {
                const V3 d = sub(p, rc[(size_t)r]);
                const V3 cr = cross(d, u);
                const double l = norm(cr);
                rr += l;
                if (want_jac) { const V3 q = cross(V3{cr.x / l, cr.y / l, cr.z / l}, d); g[0] += q.x; g[1] += q.y; g[2] += q.z; }      // d|d x u| / du
            }
            double j[3] = {0, 0, 0};
            if (want_jac) {
                const double gu = g[0] * u.x + g[1] * u.y + g[2] * u.z;           // du/da = (I - u u^T) / |a|
                j[0] = (g[0] - gu * u.x) / na; j[1] = (g[1] - gu * u.y) / na; j[2] = (g[2] - gu * u.z) / na;
            }
            if (!std::isfinite(rr)) { e.valid = false; return; }
            // HuberLoss(0.1) + Corrector (rho'' <= 0 branch)
            const double sq = rr * rr, bb = 0.1 * 0.1;
            double rho0, rho1;
            if (sq > bb) { const double s = std::sqrt(sq); rho0 = 2.0 * 0.1 * s - bb; rho1 = std::max(DBL_MIN, 0.1 / s); } else { rho0 = sq; rho1 = 1.0; }
            const double w = std::sqrt(rho1);
            for (double& v : j) v *= w;
            rr *= w;
            e.cost += 0.5 * rho0;
            e.r[r] = rr;
            for (int q = 0; q < 3; ++q) e.J[r][q] = j[q];
        }
    }, n_rows, abc);
    V3 bx = unit(V3{abc[0], abc[1], abc[2]});
    V3 bz = unit(V3{plane[0], plane[1], plane[2]});
    V3 by = cross(bx, bz);
    if (bx.x < 0) bx = V3{-bx.x, -bx.y, -bx.z};                                   // pc_process.h:427-432 (each axis on its own)
    if (by.y < 0) by = V3{-by.x, -by.y, -by.z};
    if (bz.z < 0) bz = V3{-bz.x, -bz.y, -bz.z};
    const double s = cfg->gride_side_length;
    const V3 ot = (cfg->board_height % 2 != 0) ? V3{-s * 0.5, -s * 0.5, 0} : V3{s * 0.5, -s * 0.5, 0};      // :444-458
    const V3 o{bx.x * ot.x + by.x * ot.y + bz.x * ot.z + origin_centroid.x, bx.y * ot.x + by.y * ot.y + bz.y * ot.z + origin_centroid.y,
               bx.z * ot.x + by.z * ot.y + bz.z * ot.z + origin_centroid.z};
    const V3 ax[3] = {bx, by, bz};
    for (int r = 0; r < 3; ++r) {
        Tbl[4 * r] = ax[r].x; Tbl[4 * r + 1] = ax[r].y; Tbl[4 * r + 2] = ax[r].z;
        Tbl[4 * r + 3] = -(ax[r].x * o.x + ax[r].y * o.y + ax[r].z * o.z);
    }
    Tbl[12] = 0; Tbl[13] = 0; Tbl[14] = 0; Tbl[15] = 1;
    return RCLC_OK;
}
