// boardframe.cu — the board frame from the black-block centroids: host code (24 points).
//
// Replaces get_ref_pt (include/utils.h:160-202), calc_laser_coor_v1 (include/pc_process.h:373-462) and its Ceres problem
// line_fitting / LINE_FITTING_COST (include/opt/opt_line_fitting.h:17-108): the link between eular_cluster_iter and
// opt_grid_fitting inside est_board_corner (pc_process.h:583-621). The line fit — three parameters, one Huber(0.1) residual
// per board row, ceres::Solve with DENSE_QR and default options — runs through the same trust-region schedule as the device
// solvers of this library (Jacobi scaling once, LM diagonal clamp, accept / reject / radius rule, the three stop tests), with
// the analytic Jacobian of sum |d x a/|a|| in place of Ceres' autodiff and Cholesky on the 3x3 normal equations in place of QR.
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

struct Eval3 { double cost, A[6], g[3]; bool valid; };      // A: J'J upper triangle 00 01 02 11 12 22, g: J'r

bool chol3(const double* A6, const double* l, const double* b, double* y)
{
    const double a00 = A6[0] + l[0] * l[0], a01 = A6[1], a02 = A6[2], a11 = A6[3] + l[1] * l[1], a12 = A6[4], a22 = A6[5] + l[2] * l[2];
    if (!(a00 > 0)) return false;
    const double l00 = std::sqrt(a00), l10 = a01 / l00, l20 = a02 / l00;
    const double d1 = a11 - l10 * l10;
    if (!(d1 > 0)) return false;
    const double l11 = std::sqrt(d1), l21 = (a12 - l20 * l10) / l11;
    const double d2 = a22 - l20 * l20 - l21 * l21;
    if (!(d2 > 0)) return false;
    const double l22 = std::sqrt(d2);
    const double z0 = b[0] / l00, z1 = (b[1] - l10 * z0) / l11, z2 = (b[2] - l20 * z0 - l21 * z1) / l22;
    y[2] = z2 / l22; y[1] = (z1 - l21 * y[2]) / l11; y[0] = (z0 - l10 * y[1] - l20 * y[2]) / l00;
    return std::isfinite(y[0]) && std::isfinite(y[1]) && std::isfinite(y[2]);
}

// Ceres TrustRegionMinimizer with default options (50 iterations, function_tolerance 1e-6, gradient_tolerance 1e-10,
// parameter_tolerance 1e-8, initial radius 1e4) for three parameters. x is updated unless the solve fails.
void lm3(const std::function<Eval3(const double*)>& eval, double* x)
{
    const int di[3] = {0, 3, 5};
    Eval3 e = eval(x);
    if (!e.valid) return;
    double scale[3], diag[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) scale[k] = 1.0 / (1.0 + std::sqrt(e.A[di[k]]));
    double x_cost = e.cost, radius = 1e4, dec = 2.0;
    double xn = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    bool reuse = false;
    int invalid = 0;
    auto gmax = [&](const Eval3& ev) { double m = 0; for (int k = 0; k < 3; ++k) m = std::max(m, std::fabs(x[k] - (x[k] + (-ev.g[k])))); return m; };
    bool success = true;
    for (int it = 0;;) {
        if (it >= 50) return;
        if (success && gmax(e) <= 1e-10) return;
        if (radius <= 1e-32) return;
        ++it;
        success = false;
        double As[6], bs[3], l[3], y[3];
        { int q = 0; for (int i = 0; i < 3; ++i) for (int j = i; j < 3; ++j) { As[q] = e.A[q] * scale[i] * scale[j]; ++q; } }
        for (int k = 0; k < 3; ++k) bs[k] = e.g[k] * scale[k];
        if (!reuse) for (int k = 0; k < 3; ++k) diag[k] = std::min(std::max(As[di[k]], 1e-6), 1e32);
        for (int k = 0; k < 3; ++k) l[k] = std::sqrt(diag[k] / radius);
        bool valid = chol3(As, l, bs, y);
        reuse = true;
        double step[3], mcc = 0;
        if (valid) {
            for (int k = 0; k < 3; ++k) step[k] = -y[k];
            const double sAs = step[0] * (As[0] * step[0] + As[1] * step[1] + As[2] * step[2]) + step[1] * (As[1] * step[0] + As[3] * step[1] + As[4] * step[2]) +
                               step[2] * (As[2] * step[0] + As[4] * step[1] + As[5] * step[2]);
            mcc = -((step[0] * bs[0] + step[1] * bs[1] + step[2] * bs[2]) + 0.5 * sAs);
            valid = mcc > 0;
        }
        if (!valid) { if (++invalid >= 5) return; radius *= 0.5; continue; }
        invalid = 0;
        double cand[3];
        for (int k = 0; k < 3; ++k) cand[k] = x[k] + step[k] * scale[k];
        const Eval3 ec = eval(cand);
        const double cc = ec.valid ? ec.cost : DBL_MAX;
        double sn = 0;
        for (int k = 0; k < 3; ++k) sn += (x[k] - cand[k]) * (x[k] - cand[k]);
        if (std::sqrt(sn) <= 1e-8 * (xn + 1e-8)) return;
        if (std::fabs(x_cost - cc) <= 1e-6 * x_cost) return;
        const double rel = cc >= DBL_MAX ? -DBL_MAX : (x_cost - cc) / mcc;
        if (rel > 1e-3) {
            for (int k = 0; k < 3; ++k) x[k] = cand[k];
            xn = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
            e = ec; x_cost = ec.cost; success = true;
            const double q = 2.0 * rel - 1.0;
            radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - q * q * q));
            dec = 2.0; reuse = false;
        } else { radius /= dec; dec *= 2.0; reuse = true; }
    }
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
    lm3([&](const double* x) {
        Eval3 e;
        e.cost = 0; e.valid = true;
        for (double& v : e.A) v = 0;
        for (double& v : e.g) v = 0;
        const double na = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        const V3 u{x[0] / na, x[1] / na, x[2] / na};
        for (int r = 0; r < n_rows; ++r) {
            double rr = 0, g[3] = {0, 0, 0};
            for (const V3& p : rp[(size_t)r]) {
                const V3 d = sub(p, rc[(size_t)r]);
                const V3 cr = cross(d, u);
                const double l = norm(cr);
                rr += l;
                const V3 q = cross(V3{cr.x / l, cr.y / l, cr.z / l}, d);          // d|d x u| / du
                g[0] += q.x; g[1] += q.y; g[2] += q.z;
            }
            const double gu = g[0] * u.x + g[1] * u.y + g[2] * u.z;               // du/da = (I - u u^T) / |a|
            double j[3] = {(g[0] - gu * u.x) / na, (g[1] - gu * u.y) / na, (g[2] - gu * u.z) / na};
            if (!std::isfinite(rr) || !std::isfinite(j[0]) || !std::isfinite(j[1]) || !std::isfinite(j[2])) { e.valid = false; return e; }
            // HuberLoss(0.1) + Corrector (rho'' <= 0 branch)
            const double sq = rr * rr, bb = 0.1 * 0.1;
            double rho0, rho1;
            if (sq > bb) { const double s = std::sqrt(sq); rho0 = 2.0 * 0.1 * s - bb; rho1 = std::max(DBL_MIN, 0.1 / s); } else { rho0 = sq; rho1 = 1.0; }
            const double w = std::sqrt(rho1);
            rr *= w;
            for (double& v : j) v *= w;
            e.cost += 0.5 * rho0;
            int q = 0;
            for (int a = 0; a < 3; ++a) for (int b2 = a; b2 < 3; ++b2) e.A[q++] += j[a] * j[b2];
            for (int a = 0; a < 3; ++a) e.g[a] += j[a] * rr;
        }
        return e;
    }, abc);
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
