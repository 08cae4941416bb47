// oracle/rclc_oracle.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; see rclc_oracle.h header note).
//
// CPU restatement of the zhijianglu/RCLC calibration hot path, function by function, each citing
// the reference file:line it follows. Third-party arithmetic (PCL, Ceres, Eigen) is restated from
// the published algorithms. PARITY UNPINNED (no reference tests exist).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: the reference is built without -march
// flags, so x86-64 baseline SSE2, no FMA contraction — every a*b+c below is two roundings).
#define RCLC_CONFIG_IMPLEMENTATION
#include "rclc_oracle.h"
#include "ceres_lm.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <random>
#include <map>
#include <cfloat>
#include <vector>

namespace {

struct P4 { float x, y, z, i; };

inline const float* rec(const void* base, int stride, int64_t k)
{
    return reinterpret_cast<const float*>(static_cast<const char*>(base) + k * (int64_t)stride);
}
inline float* rec(void* base, int stride, int64_t k)
{
    return reinterpret_cast<float*>(static_cast<char*>(base) + k * (int64_t)stride);
}
inline float rec_i(const void* base, int stride, int ioff, int64_t k)
{
    float v;
    std::memcpy(&v, static_cast<const char*>(base) + k * (int64_t)stride + ioff, 4);
    return v;
}

std::vector<P4> load_cloud(const void* pts, int stride, int ioff, int64_t n)
{
    std::vector<P4> c(n);
    for (int64_t k = 0; k < n; ++k) {
        const float* r = rec(pts, stride, k);
        c[k] = {r[0], r[1], r[2], rec_i(pts, stride, ioff, k)};
    }
    return c;
}

// ---------------------------------------------------------------------------------------------
// include/utils.h:286-316  generate_std_corners. Expressions kept in the reference's order.
// ---------------------------------------------------------------------------------------------
struct Target { double x, y; bool is_row; };

void generate_std_corners(const rclc_config& c, std::vector<Target>& row_points, std::vector<Target>& col_points)
{
    const int bx = c.board_width, by = c.board_height;
    const int centroid_x_id = bx / 2;
    const int centroid_y_id = by / 2;
    const double s = c.gride_side_length;
    const double row_shift_x = centroid_x_id * s;
    const double row_shift_y = double(centroid_y_id - 1) * s;
    for (int r = 0; r < by - 1; ++r)
        for (int cc = 0; cc < bx; ++cc)
            row_points.push_back({s / 2.0 + cc * s - row_shift_x, r * s - row_shift_y, true});
    const double col_shift_y = double(centroid_y_id) * s;
    const double col_shift_x = double(centroid_x_id - 1) * s;
    for (int cc = 0; cc < bx - 1; ++cc)
        for (int r = 0; r < by; ++r)
            col_points.push_back({cc * s - col_shift_x, s / 2.0 + r * s - col_shift_y, false});
}

// Residual-block order of opt_grid_fitting.h:78-94 : all col targets, then all row targets.
std::vector<Target> problem_targets(const rclc_config& c)
{
    std::vector<Target> rows, cols;
    generate_std_corners(c, rows, cols);
    std::vector<Target> t(cols);
    t.insert(t.end(), rows.begin(), rows.end());
    return t;
}

// ---------------------------------------------------------------------------------------------
// include/opt/board_fitting_cost.h:34-62  radii
// ---------------------------------------------------------------------------------------------
struct Radii { double neighbour_radius, cost_radius2, cost_gradient_radius2; };

Radii make_radii(const rclc_config& c)
{
    Radii r;
    const double factor = std::sqrt(2.0);
    r.neighbour_radius = c.neighbour_radius_rate * c.gride_side_length;
    double cost_radius = c.cost_radius_rate * c.gride_side_length;
    double cost_gradient_radius = cost_radius * factor;
    if (!(r.neighbour_radius > cost_gradient_radius)) {
        cost_gradient_radius = 0.9 * r.neighbour_radius;
        cost_radius = cost_gradient_radius / factor;
    }
    r.cost_radius2 = cost_radius * cost_radius;
    r.cost_gradient_radius2 = cost_gradient_radius * cost_gradient_radius;
    return r;
}

// ---------------------------------------------------------------------------------------------
// board_fitting_cost.h:264-280  update_neighbour_points == pcl::KdTreeFLANN::radiusSearch.
// FLANN (kdtree_single_index.h, L2_Simple<float>): squared distance accumulated in float over
// x,y,z in that order, kept when dist < (float)(radius*radius) (strict). Exhaustive restatement;
// neighbour order = ascending original index (PCL returns them sorted by distance; order only
// affects fp64 summation order inside Evaluate).
// ---------------------------------------------------------------------------------------------
inline bool flann_within(float qx, float qy, float qz, const P4& p, float r2)
{
    float result = 0.0f, diff;
    diff = qx - p.x; result += diff * diff;
    diff = qy - p.y; result += diff * diff;
    diff = qz - p.z; result += diff * diff;
    return result < r2;
}

void build_neighbours(const std::vector<P4>& cloud, const std::vector<Target>& targets, double neighbour_radius,
                      std::vector<std::vector<int>>& nb)
{
    const float r2 = static_cast<float>(neighbour_radius * neighbour_radius);
    nb.assign(targets.size(), {});
    // uniform-grid acceleration with an exact final test (same result as exhaustive search)
    for (size_t t = 0; t < targets.size(); ++t) nb[t].reserve(cloud.size() / 20 + 16);
    const double reach = neighbour_radius * 1.01 + 1e-6;
    for (int k = 0; k < (int)cloud.size(); ++k) {
        const P4& p = cloud[k];
        for (size_t t = 0; t < targets.size(); ++t) {
            if (std::fabs((double)p.x - targets[t].x) > reach || std::fabs((double)p.y - targets[t].y) > reach) continue;
            if (flann_within((float)targets[t].x, (float)targets[t].y, 0.0f, p, r2)) nb[t].push_back(k);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// board_fitting_cost.h:74-262  BOARD_FITTING_COST::Evaluate for one target.
// acc16 (optional): cost sum d,u,l,r | cost cnt d,u,l,r | interval sum d,u,l,r | interval cnt d,u,l,r
// ---------------------------------------------------------------------------------------------
void evaluate_target(const Target& tg, const std::vector<P4>& cloud, const std::vector<int>& nb, const Radii& rad,
                     const double* params, double* residual, double* jac3, double* acc16)
{
    const double theta = (params[2] * double(M_PI)) / (180.0);                              // :80
    // :82,113  Eigen::Quaternion(w,0,0,z).matrix(): tz=2z, twz=tz*w, tzz=tz*z
    const double qw = std::cos(theta / 2.0), qz = std::sin(theta / 2.0);
    const double tz = 2.0 * qz, twz = tz * qw, tzz = tz * qz;
    const double R00 = 1.0 - (0.0 + tzz), R01 = 0.0 - twz, R10 = 0.0 + twz, R11 = 1.0 - (0.0 + tzz);
    const double tx = params[0], ty = params[1];

    double cs[4] = {0, 0, 0, 0}, cc[4] = {0, 0, 0, 0};   // cost intensity / cnt: down, up, left, right
    double is[4] = {0, 0, 0, 0}, ic[4] = {0, 0, 0, 0};   // interval
    const double ref_x = tg.x, ref_y = tg.y;
    for (size_t j = 0; j < nb.size(); ++j) {                                                // :124
        const P4& p = cloud[nb[j]];
        const double neig_x = double(p.x), neig_y = double(p.y), neig_intensity = double(p.i);
        const double vx = neig_x + tx, vy = neig_y + ty;                                     // :133
        const double xt = R00 * vx + R01 * vy;       // third term R02*(0+0) == +0 exactly
        const double yt = R10 * vx + R11 * vy;
        const double dx = std::max(double(10e-15), std::fabs(ref_x - xt));                   // :138
        const double dy = std::max(double(10e-15), std::fabs(ref_y - yt));
        const double distance_2 = dx * dx + dy * dy;
        if (distance_2 > rad.cost_gradient_radius2) continue;                                // :142
        const bool in_cost = distance_2 < rad.cost_radius2;
        const int ud = (yt > ref_y) ? 0 : 1;                                                 // :147 down : up
        const int lr = (xt > ref_x) ? 3 : 2;                                                 // :178 right : left
        if (in_cost) { cs[ud] += neig_intensity; cc[ud] += 1; cs[lr] += neig_intensity; cc[lr] += 1; }
        else         { is[ud] += neig_intensity; ic[ud] += 1; is[lr] += neig_intensity; ic[lr] += 1; }
    }
    const double C_d = cs[0] / cc[0], C_u = cs[1] / cc[1], C_l = cs[2] / cc[2], C_r = cs[3] / cc[3];   // :211-214
    const double dC_d = is[0] / ic[0], dC_u = is[1] / ic[1], dC_l = is[2] / ic[2], dC_r = is[3] / ic[3];
    const double cost_ud = std::fabs(C_d - C_u);
    const double cost_lr = std::fabs(C_r - C_l);
    const double gradient_x = std::fabs(C_r - dC_r) - std::fabs(C_l - dC_l);                 // :224
    const double gradient_y = std::fabs(C_d - dC_d) - std::fabs(C_u - dC_u);
    residual[0] = tg.is_row ? 1.0 - std::tanh((cost_ud) / 100.0) : 1.0 - std::tanh((cost_lr) / 100.0);  // :227-236
    if (jac3) {
        const double dp_dtheta1 = -tx * std::cos(theta) - ty * std::sin(theta);              // :252
        const double dp_dtheta2 = tx * std::sin(theta) - ty * std::cos(theta);
        jac3[0] = gradient_x;
        jac3[1] = gradient_y;
        jac3[2] = gradient_x * dp_dtheta1 + gradient_y * dp_dtheta2;
    }
    if (acc16) {
        for (int k = 0; k < 4; ++k) { acc16[k] = cs[k]; acc16[4 + k] = cc[k]; acc16[8 + k] = is[k]; acc16[12 + k] = ic[k]; }
    }
}

struct GridProblem {
    const std::vector<P4>* cloud;
    std::vector<Target> targets;
    std::vector<std::vector<int>> nb;
    Radii rad;
    double huber_a;
    int64_t evaluate_calls = 0;

    // ProgramEvaluator::Evaluate over the residual blocks (1 residual, 3 params each, shared loss).
    bool evaluate(const double* x, double* cost, double* residuals, double* jac)
    {
        const int m = (int)targets.size();
        double total = 0;
        bool ok = true;
        for (int t = 0; t < m; ++t) {
            double r, j[3];
            evaluate_target(targets[t], *cloud, nb[t], rad, x, &r, jac ? j : nullptr, nullptr);
            evaluate_calls++;
            if (!std::isfinite(r)) ok = false;                                 // IsEvaluationValid
            if (jac && !(std::isfinite(j[0]) && std::isfinite(j[1]) && std::isfinite(j[2]))) ok = false;
            if (!ok) return false;
            total += orc::apply_loss_1d(huber_a, &r, jac ? j : nullptr, 3);
            if (residuals) residuals[t] = r;
            if (jac) { jac[t * 3 + 0] = j[0]; jac[t * 3 + 1] = j[1]; jac[t * 3 + 2] = j[2]; }
        }
        *cost = total;
        return true;
    }
};

// include/utils.h:273-284  SE2_to_SE3<float>: double quaternion matrix cast to float, row-major out.
void se2_to_se3f(double x, double y, double yaw_ang, float T[16])
{
    const double ang = (yaw_ang * double(M_PI)) / double(180.0);
    const double qw = std::cos(ang / 2.0), qz = std::sin(ang / 2.0);
    const double tz = 2.0 * qz, twz = tz * qw, tzz = tz * qz;
    const double R[9] = {1.0 - (0.0 + tzz), 0.0 - twz, 0.0, 0.0 + twz, 1.0 - (0.0 + tzz), 0.0, 0.0, 0.0, 1.0};
    for (int k = 0; k < 16; ++k) T[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[r * 4 + c] = (float)R[r * 3 + c];
    T[3] = (float)x;
    T[7] = (float)y;
    T[11] = 0.0f;
}

// pcl::transformPointCloud(cloud, cloud, Eigen::Matrix4f) — float, ((m0*x + m1*y) + m2*z) + m3
// (PCL <= 1.9 scalar form; PCL's SSE form differs in association only. Unpinned: this is the definition.)
inline void transform_point_f(const float T[16], float& x, float& y, float& z)
{
    const float nx = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
    const float ny = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
    const float nz = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
    x = nx; y = ny; z = nz;
}

void mat4f_mul(const float A[16], const float B[16], float C[16])
{
    float out[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float s = A[r * 4 + 0] * B[0 * 4 + c];
            s = s + A[r * 4 + 1] * B[1 * 4 + c];
            s = s + A[r * 4 + 2] * B[2 * 4 + c];
            s = s + A[r * 4 + 3] * B[3 * 4 + c];
            out[r * 4 + c] = s;
        }
    std::memcpy(C, out, sizeof(out));
}

void mat4d_mul(const double A[16], const double B[16], double C[16])
{
    double out[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[r * 4 + k] * B[k * 4 + c];
            out[r * 4 + c] = s;
        }
    std::memcpy(C, out, sizeof(out));
}

// General 4x4 inverse by cofactors (Eigen::Matrix4d::inverse() is the cofactor form as well).
bool inverse4(const double m[16], double inv_out[16])
{
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0) return false;
    const double id = 1.0 / det;
    for (int i = 0; i < 16; i++) inv_out[i] = inv[i] * id;
    return true;
}

// include/utils.h:47-66  calc_transfered_standard_corners
void transfered_standard_corners(const rclc_config& c, const double T_laser_base[16], double* out)
{
    const double s = c.gride_side_length;
    const double first_corner_x = double(int((c.board_width - 1) / 2)) * s;
    const double first_corner_y = double(int((c.board_height - 1) / 2)) * s;
    int k = 0;
    for (int row = 0; row < c.board_height - 1; ++row)
        for (int col = 0; col < c.board_width - 1; ++col) {
            const double p[4] = {first_corner_x - double(col) * s, first_corner_y - double(row) * s, 0, 1};
            for (int r = 0; r < 3; ++r) {
                double v = 0;
                for (int q = 0; q < 4; ++q) v += T_laser_base[r * 4 + q] * p[q];
                out[k * 3 + r] = v;
            }
            ++k;
        }
}

// Eigen::Quaternion::toRotationMatrix (templated so the pose cost can run it on Jets)
template <typename T>
void quat_to_R(const T& w, const T& x, const T& y, const T& z, T R[9])
{
    const T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w;
    const T txx = tx * x, txy = ty * x, txz = tz * x;
    const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = T(1) - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
    R[3] = txy + twz;          R[4] = T(1) - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = T(1) - (txx + tyy);
}

// ceres::Jet<double, N> — the operations POSE_REPRJ_COST needs.
template <int N>
struct Jet {
    double a;
    double v[N];
    Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
    Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }   // NOLINT
    Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> Jet<N> operator/(const Jet<N>& f, const Jet<N>& g)
{
    // jet.h: g_a_inverse = 1/g.a; f_a_by_g_a = f.a * g_a_inverse; v = (f.v - f_a_by_g_a * g.v) * g_a_inverse
    Jet<N> h;
    const double g_a_inverse = 1.0 / g.a;
    const double f_a_by_g_a = f.a * g_a_inverse;
    h.a = f_a_by_g_a;
    for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
    return h;
}
template <int N> Jet<N> jsqrt(const Jet<N>& f)
{
    Jet<N> h;
    const double tmp = std::sqrt(f.a);
    const double two_a_inverse = 1.0 / (2.0 * tmp);
    h.a = tmp;
    for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
    return h;
}
inline double jsqrt(double f) { return std::sqrt(f); }
template <int N> bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }

// ---------------------------------------------------------------------------------------------
// include/opt/pose_reprj_cost.hpp:28-55  POSE_REPRJ_COST::operator()
// Small fixed-size Eigen products/reductions associate as a0 + (a1 + a2) (redux_novec_unroller).
// ---------------------------------------------------------------------------------------------
template <typename T>
T pose_residual(const T* q, const T* t, const double* pc, const double* img, int C)
{
    T R[9];
    quat_to_R(q[3], q[0], q[1], q[2], R);
    T residual = T(0);
    T depth_max = T(0);
    for (int i = 0; i < C; ++i) {
        const T p0 = T(pc[i * 3 + 0]), p1 = T(pc[i * 3 + 1]), p2 = T(pc[i * 3 + 2]);
        const T X = (R[0] * p0 + (R[1] * p1 + R[2] * p2)) + t[0];
        const T Y = (R[3] * p0 + (R[4] * p1 + R[5] * p2)) + t[1];
        const T Z = (R[6] * p0 + (R[7] * p1 + R[8] * p2)) + t[2];
        const T depth_norm = jsqrt(X * X + (Y * Y + Z * Z));
        if (depth_norm > depth_max) depth_max = depth_norm;
        const T e0 = T(img[i * 2 + 0]) - X / Z;
        const T e1 = T(img[i * 2 + 1]) - Y / Z;
        const T error = jsqrt(e0 * e0 + e1 * e1) / depth_norm;
        residual = residual + error;
    }
    residual = residual * depth_max;
    return residual;
}

// ceres::EigenQuaternionParameterization (storage x,y,z,w)
// rn_libm: sin and cos correctly rounded (ceres_lm.hpp rn::sincos) instead of this libm's
void quat_plus(const double* x, const double* delta, double* out, bool rn_libm)
{
    const double norm_delta = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (norm_delta > 0.0) {
        double sin_nd, cos_nd;
        if (rn_libm) orc::rn::sincos(norm_delta, &sin_nd, &cos_nd);
        else { sin_nd = std::sin(norm_delta); cos_nd = std::cos(norm_delta); }
        const double s = sin_nd / norm_delta;
        const double dw = cos_nd, dx = s * delta[0], dy = s * delta[1], dz = s * delta[2];
        const double xw = x[3], xx = x[0], xy = x[1], xz = x[2];
        // Eigen quaternion product a*b
        const double w = dw * xw - dx * xx - dy * xy - dz * xz;
        const double X = dw * xx + dx * xw + dy * xz - dz * xy;
        const double Y = dw * xy + dy * xw + dz * xx - dx * xz;
        const double Z = dw * xz + dz * xw + dx * xy - dy * xx;
        out[0] = X; out[1] = Y; out[2] = Z; out[3] = w;
    } else {
        for (int i = 0; i < 4; ++i) out[i] = x[i];
    }
}
void quat_plus_jacobian(const double* x, double* J /*4x3 row-major*/)
{
    J[0] = x[3];  J[1] = x[2];   J[2] = -x[1];
    J[3] = -x[2]; J[4] = x[3];   J[5] = x[0];
    J[6] = x[1];  J[7] = -x[0];  J[8] = x[3];
    J[9] = -x[0]; J[10] = -x[1]; J[11] = -x[2];
}

// Residual + 1x6 local Jacobian (AutoDiffCostFunction<POSE_REPRJ_COST,1,4,3> then
// LocalParameterization::MultiplyByJacobian) for one frame.
void pose_block(const double* Tcl, const double* pc, const double* img, int C, double* r, double* j6)
{
    typedef Jet<7> J7;
    J7 q[4], t[3];
    for (int k = 0; k < 4; ++k) q[k] = J7(Tcl[k], k);
    for (int k = 0; k < 3; ++k) t[k] = J7(Tcl[4 + k], 4 + k);
    const J7 res = pose_residual<J7>(q, t, pc, img, C);
    *r = res.a;
    if (j6) {
        double P[12];
        quat_plus_jacobian(Tcl, P);
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += res.v[k] * P[k * 3 + c];
            j6[c] = s;
        }
        for (int c = 0; c < 3; ++c) j6[3 + c] = res.v[4 + c];
    }
}

// ---------------------------------------------------------------------------------------------
// PCL SampleConsensusModelPlane (sac_model_plane.hpp) in float, no FMA.
// ---------------------------------------------------------------------------------------------
bool plane_from_triplet(const float* p0, const float* p1, const float* p2, float coeff[4])
{
    const float a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
    const float b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
    const float d0 = a0 / b0, d1 = a1 / b1, d2 = a2 / b2;
    if ((d0 == d1) && (d2 == d1)) return false;                       // collinearity check
    float c0 = a1 * b2 - a2 * b1;
    float c1 = a2 * b0 - a0 * b2;
    float c2 = a0 * b1 - a1 * b0;
    // model_coefficients.normalize() on a Vector4f with [3]=0: squaredNorm then divide by sqrt.
    // Eigen SSE2 packet reduction: (c0^2 + c2^2) + (c1^2 + 0)
    const float n2 = (c0 * c0 + c2 * c2) + (c1 * c1 + 0.0f);
    const float nrm = std::sqrt(n2);
    c0 = c0 / nrm; c1 = c1 / nrm; c2 = c2 / nrm;
    // d = -1 * coeff.dot(p0) with coeff[3]=0, p0[3]=1:  (c0*x + c2*z) + (c1*y + 0*1)
    const float dot = (c0 * p0[0] + c2 * p0[2]) + (c1 * p0[1] + 0.0f);
    coeff[0] = c0; coeff[1] = c1; coeff[2] = c2; coeff[3] = -1.0f * dot;
    return true;
}
inline float plane_dot(const float c[4], const float* p)
{
    // Eigen Vector4f dot, SSE2 predux: (c0*x + c2*z) + (c1*y + c3*1)
    return (c[0] * p[0] + c[2] * p[2]) + (c[1] * p[1] + c[3] * 1.0f);
}
inline bool plane_inlier(const float c[4], const float* p, double thr)
{
    return (double)std::fabs(plane_dot(c, p)) < thr;                  // std::abs(float) < double threshold
}

// 3x3 symmetric eigen decomposition (Jacobi rotations, double): smallest eigenvalue's vector.
void smallest_eigenvector(const double cov[9], double evec[3])
{
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = cov[r * 3 + c];
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int k = 1; k < 3; ++k) if (a[k][k] < a[best][best]) best = k;
    for (int k = 0; k < 3; ++k) evec[k] = v[k][best];
}

// cvRound: SSE2 cvtsd2si — round-half-even, 0x80000000 ("integer indefinite") for NaN / out of range
inline int cv_round(double v) { return (std::fabs(v) < 2147483648.0) ? (int)std::nearbyint(v) : INT32_MIN; }

} // namespace

extern "C" {

int orc_generate_targets(const rclc_config* cfg, double* targets_xy, int32_t* is_row, int32_t* n_targets)
{
    const std::vector<Target> t = problem_targets(*cfg);
    for (size_t k = 0; k < t.size(); ++k) {
        targets_xy[2 * k] = t[k].x;
        targets_xy[2 * k + 1] = t[k].y;
        is_row[k] = t[k].is_row ? 1 : 0;
    }
    *n_targets = (int32_t)t.size();
    return 0;
}

int orc_gridfit_eval(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                     const double pose[3], double* residuals, double* jac, double* acc)
{
    const std::vector<P4> cloud = load_cloud(pts, stride, ioff, n);
    const std::vector<Target> targets = problem_targets(*cfg);
    const Radii rad = make_radii(*cfg);
    std::vector<std::vector<int>> nb;
    build_neighbours(cloud, targets, rad.neighbour_radius, nb);
    for (size_t t = 0; t < targets.size(); ++t)
        evaluate_target(targets[t], cloud, nb[t], rad, pose, &residuals[t], jac ? &jac[3 * t] : nullptr,
                        acc ? &acc[16 * t] : nullptr);
    return 0;
}

int orc_gridfit_cost(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                     const double pose[3], double* cost)
{
    const std::vector<P4> cloud = load_cloud(pts, stride, ioff, n);
    GridProblem gp;
    gp.cloud = &cloud;
    gp.targets = problem_targets(*cfg);
    gp.rad = make_radii(*cfg);
    gp.huber_a = cfg->gridfit_huber_a;
    build_neighbours(cloud, gp.targets, gp.rad.neighbour_radius, gp.nb);
    return gp.evaluate(pose, cost, nullptr, nullptr) ? 0 : 1;
}

// include/opt/opt_grid_fitting.h:25-132
int orc_gridfit_solve(const rclc_config* cfg, void* pts_io, int stride, int ioff, int64_t n,
                      double* T_bl_io, double* corners_out, orc_gridfit_report* rep)
{
    // :32-38 optional UniformSampling is handled by the caller (shipped config: radius 0 => skipped)
    if (cfg->apply_noise_test) {                                               // :47-56
        float Tn[16];
        se2_to_se3f(cfg->noise_x, cfg->noise_y, cfg->noise_yaml, Tn);
        for (int64_t k = 0; k < n; ++k) { float* r = rec(pts_io, stride, k); transform_point_f(Tn, r[0], r[1], r[2]); }
        double Tnd[16];
        for (int k = 0; k < 16; ++k) Tnd[k] = (double)Tn[k];
        mat4d_mul(Tnd, T_bl_io, T_bl_io);
    }
    std::vector<P4> pc_iter = load_cloud(pts_io, stride, ioff, n);             // :60-61
    float T_base_laser[16];
    for (int k = 0; k < 16; ++k) T_base_laser[k] = (k % 5 == 0) ? 1.0f : 0.0f;

    GridProblem gp;
    gp.cloud = &pc_iter;
    gp.targets = problem_targets(*cfg);
    gp.rad = make_radii(*cfg);
    gp.huber_a = cfg->gridfit_huber_a;
    const int m = (int)gp.targets.size();

    orc::LMOptions opt;                                                        // :96-101
    opt.linear_solver = orc::DENSE_QR;
    opt.max_num_iterations = cfg->gridfit_max_lm_iters;
    opt.function_tolerance = cfg->gridfit_function_tol;

    if (rep) { std::memset(rep, 0, sizeof(*rep)); }
    bool shouldQuit = false;
    int iter = 0;
    while (!shouldQuit) {                                                      // :68
        double T_se2_wl[3] = {0.0, 0.0, 0.0};
        build_neighbours(pc_iter, gp.targets, gp.rad.neighbour_radius, gp.nb); // :75-94 (kd-tree + per-target radius search)
        orc::LMProblem P;
        P.m = m; P.n = 3; P.nl = 3;
        P.evaluate = [&](const double* x, double* cost, double* r, double* j) { return gp.evaluate(x, cost, r, j); };
        const orc::LMSummary sum = orc::lm_minimize(P, opt, T_se2_wl);         // :103
        float T_w1_w0[16];
        se2_to_se3f(T_se2_wl[0], T_se2_wl[1], T_se2_wl[2], T_w1_w0);           // :105-106
        for (auto& p : pc_iter) transform_point_f(T_w1_w0, p.x, p.y, p.z);     // :107
        const double decrease_rate = (sum.initial_cost - sum.final_cost) / sum.initial_cost;   // :113
        iter++;
        if (rep) {
            if (iter == 1) { rep->first_initial_cost = sum.initial_cost; for (int k = 0; k < 3; ++k) rep->se2_first[k] = T_se2_wl[k]; }
            rep->last_final_cost = sum.final_cost;
            rep->pose_evals += sum.num_cost_evals;
            rep->lm_iterations += sum.num_iterations;
            if (sum.termination == orc::FAILURE) rep->status = 1;
        }
        if ((std::fabs(decrease_rate) < 1e-6) || iter > cfg->max_iter_time) {  // :116
            if (decrease_rate > 0) mat4f_mul(T_w1_w0, T_base_laser, T_base_laser);
            shouldQuit = true;
        } else {
            mat4f_mul(T_w1_w0, T_base_laser, T_base_laser);
        }
    }
    double Tbld[16], final_T_base_laser[16], final_T_laser_base[16];
    for (int k = 0; k < 16; ++k) Tbld[k] = (double)T_base_laser[k];
    mat4d_mul(Tbld, T_bl_io, final_T_base_laser);                              // :129
    if (!inverse4(final_T_base_laser, final_T_laser_base)) return 2;           // :130
    transfered_standard_corners(*cfg, final_T_laser_base, corners_out);        // :131
    if (rep) {
        rep->outer_iterations = iter;
        rep->evaluate_calls = gp.evaluate_calls;
        std::memcpy(rep->T_base_laser, T_base_laser, sizeof(T_base_laser));
    }
    return 0;
}

// src/GMMFit.cpp:7-108 + src/GaussianFunction.cpp:7-23
int orc_gmm_fit_1d(const void* intensity, int stride, int64_t n, int K, int iters,
                   double* mu, double* sigma, double* priors_out)
{
    std::vector<double> x(n);
    for (int64_t i = 0; i < n; ++i) x[i] = (double)*rec(intensity, stride, i);   // pc_process.h:605-606 float->double
    std::vector<double> priors(K, 1.0 / double(K));                              // :29-31
    std::vector<double> L((size_t)n * K), P((size_t)n * K);
    for (int itr = 0; itr < iters; ++itr) {                                      // :37
        for (int k = 0; k < K; ++k) {                                            // :46-50
            const double factor = 1.0 / (sigma[k] * std::sqrt(2.0 * M_PI));
            const double sigma_sqr = sigma[k] * sigma[k];
            for (int64_t i = 0; i < n; ++i)
                L[i * K + k] = factor * std::exp(-1.0 / (2 * sigma_sqr) * (x[i] - mu[k]) * (x[i] - mu[k]));
        }
        for (int64_t i = 0; i < n; ++i) {                                        // :59-65
            double p_x = 0;
            for (int k = 0; k < K; ++k) p_x += L[i * K + k] * priors[k];
            for (int k = 0; k < K; ++k) P[i * K + k] = (L[i * K + k] * priors[k]) / p_x;
        }
        std::vector<double> denoms(K);
        for (int k = 0; k < K; ++k) {                                            // :73-93
            double denom = 0, mu_num = 0;
            for (int64_t i = 0; i < n; ++i) denom += P[i * K + k];
            for (int64_t i = 0; i < n; ++i) mu_num += P[i * K + k] * x[i];
            const double mu_new = mu_num / denom;
            double sig_num = 0;
            for (int64_t i = 0; i < n; ++i) sig_num += P[i * K + k] * ((x[i] - mu_new) * (x[i] - mu_new));
            mu[k] = mu_new;
            sigma[k] = std::sqrt(sig_num / denom);
            denoms[k] = denom;
        }
        for (int k = 0; k < K; ++k) priors[k] = denoms[k] / double(n);           // :98-101
    }
    if (priors_out) for (int k = 0; k < K; ++k) priors_out[k] = priors[k];
    return 0;
}

// include/pc_process.h:41-57
int orc_plane_project(void* pts_io, int stride, int64_t n, const double plane[4])
{
    for (int64_t k = 0; k < n; ++k) {
        float* pt = rec(pts_io, stride, k);
        const double d0 = (double)(pt[0] / pt[2]);          // Vector2d(pt.x / pt.z, ...) : float division, then widened
        const double d1 = (double)(pt[1] / pt[2]);
        pt[2] = (float)(-plane[3] / (plane[0] * d0 + plane[1] * d1 + plane[2]));
        pt[0] = (float)((double)pt[2] * d0);                // float * double -> double -> float
        pt[1] = (float)((double)pt[2] * d1);
    }
    return 0;
}

int orc_transform_cloud(void* pts_io, int stride, int64_t n, const float T[16])
{
    for (int64_t k = 0; k < n; ++k) { float* r = rec(pts_io, stride, k); transform_point_f(T, r[0], r[1], r[2]); }
    return 0;
}

int orc_ransac_score(const void* pts, int stride, int64_t n, const int32_t* triplets, int H,
                     double thr, float* planes_out, int32_t* valid_out, int32_t* counts_out)
{
    for (int h = 0; h < H; ++h) {
        float c[4] = {0, 0, 0, 0};
        const bool ok = plane_from_triplet(rec(pts, stride, triplets[3 * h]), rec(pts, stride, triplets[3 * h + 1]),
                                           rec(pts, stride, triplets[3 * h + 2]), c);
        int32_t cnt = 0;
        if (ok)
            for (int64_t k = 0; k < n; ++k) cnt += plane_inlier(c, rec(pts, stride, k), thr) ? 1 : 0;
        for (int q = 0; q < 4; ++q) planes_out[4 * h + q] = c[q];
        valid_out[h] = ok ? 1 : 0;
        counts_out[h] = cnt;
    }
    return 0;
}

int orc_plane_select(const void* pts, int stride, int64_t n, const float plane[4], double thr, uint8_t* mask_out)
{
    for (int64_t k = 0; k < n; ++k) mask_out[k] = plane_inlier(plane, rec(pts, stride, k), thr) ? 1 : 0;
    return 0;
}

// pcl/sample_consensus/impl/ransac.hpp computeModel
int orc_ransac_replay(const int32_t* counts, const int32_t* valid, int H, int64_t n_indices,
                      int max_iterations, double probability, int32_t* best_out, int32_t* iters_out)
{
    int iterations = 0;
    int n_best = -INT32_MAX;
    double k = 1.0;
    int skipped = 0;
    const int max_skip = max_iterations * 10;
    const double log_probability = std::log(1.0 - probability);
    const double one_over_indices = 1.0 / static_cast<double>(n_indices);
    int best = -1, h = 0;
    while (iterations < k && skipped < max_skip) {
        if (h >= H) break;                              // hypothesis list exhausted
        const int cur = h++;
        if (!valid[cur]) { ++skipped; continue; }
        if (counts[cur] > n_best) {
            n_best = counts[cur];
            best = cur;
            const double w = static_cast<double>(n_best) * one_over_indices;
            double p_no_outliers = 1.0 - std::pow(w, 3.0);
            p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
            p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
            k = log_probability / std::log(p_no_outliers);
        }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    *best_out = best;
    *iters_out = iterations;
    return 0;
}

// Exact, order-free sum: terms floored onto the fixed-point grid of 2^-80, added as 128-bit integers, rounded to a double once.
namespace {
struct ExactSum {
    __int128 v = 0;
    void add(double t)
    {
        const double s = t * 65536.0, fl = std::floor(s);
        const __int128 hi = (__int128)(long long)fl;
        const unsigned long long lo = (unsigned long long)((s - fl) * 18446744073709551616.0);
        v += hi * ((__int128)1 << 64) + (__int128)lo;
    }
    double value() const { return std::ldexp((double)v, -80); }       // int128 -> double: to nearest even
};
} // namespace

int orc_plane_refine(const void* pts, int stride, int64_t n, const uint8_t* mask, const float plane_in[4],
                     float plane_out[4], float centroid_out[3])
{
    // computeMeanAndCovarianceMatrix single pass. PCL's accumulation type and order differ by version and are unpinned; the
    // oracle DEFINES the nine sums as exact, order-free sums: every term (x, y, z and the products of two floats, exact in a
    // double) is floored onto the fixed-point grid of 2^-80 and added as a 128-bit integer, the total is rounded to a double
    // once. Normalised like PCL: accu /= n; cov = E[xx'] - mean mean'.
    ExactSum acc128[9];
    auto add = [](ExactSum& a, double t) { a.add(t); };
    int64_t cnt = 0;
    for (int64_t k = 0; k < n; ++k) {
        if (!mask[k]) continue;
        const float* p = rec(pts, stride, k);
        const double x = p[0], y = p[1], z = p[2];
        add(acc128[0], x * x); add(acc128[1], x * y); add(acc128[2], x * z); add(acc128[3], y * y); add(acc128[4], y * z); add(acc128[5], z * z);
        add(acc128[6], x); add(acc128[7], y); add(acc128[8], z);
        ++cnt;
    }
    double accu[9];
    for (int q = 0; q < 9; ++q) accu[q] = acc128[q].value();
    if (cnt <= 3) { for (int q = 0; q < 4; ++q) plane_out[q] = plane_in[q]; centroid_out[0] = centroid_out[1] = centroid_out[2] = 0; return 0; }
    for (int q = 0; q < 9; ++q) accu[q] /= (double)cnt;
    double cov[9];
    cov[0] = accu[0] - accu[6] * accu[6];
    cov[1] = accu[1] - accu[6] * accu[7];
    cov[2] = accu[2] - accu[6] * accu[8];
    cov[4] = accu[3] - accu[7] * accu[7];
    cov[5] = accu[4] - accu[7] * accu[8];
    cov[8] = accu[5] - accu[8] * accu[8];
    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
    double ev[3];
    smallest_eigenvector(cov, ev);
    // sign convention: keep the hemisphere of the input model (eigenvector sign is arbitrary)
    if (ev[0] * plane_in[0] + ev[1] * plane_in[1] + ev[2] * plane_in[2] < 0) { ev[0] = -ev[0]; ev[1] = -ev[1]; ev[2] = -ev[2]; }
    plane_out[0] = (float)ev[0]; plane_out[1] = (float)ev[1]; plane_out[2] = (float)ev[2];
    centroid_out[0] = (float)accu[6]; centroid_out[1] = (float)accu[7]; centroid_out[2] = (float)accu[8];
    const float dot = (plane_out[0] * centroid_out[0] + plane_out[2] * centroid_out[2]) + (plane_out[1] * centroid_out[1] + 0.0f);
    plane_out[3] = -1.0f * dot;
    return 0;
}

int orc_ransac_draw_triplets(int64_t n_indices, int H, int32_t* triplets_out)
{
    // SampleConsensusModel: boost::mt19937 rng(12345), uniform_int<>(0, INT_MAX) => rng()>>1;
    // drawIndexSample: swap(idx[i], idx[i + rnd() % (size - i)]) for i in 0..2 on a persistent shuffled index list.
    std::mt19937 rng(12345u);
    std::vector<int32_t> idx(n_indices);
    for (int64_t i = 0; i < n_indices; ++i) idx[i] = (int32_t)i;
    for (int h = 0; h < H; ++h) {
        for (int i = 0; i < 3; ++i) {
            const uint32_t r = rng() >> 1;
            std::swap(idx[i], idx[i + (r % (n_indices - i))]);
        }
        for (int i = 0; i < 3; ++i) triplets_out[3 * h + i] = idx[i];
    }
    return 0;
}

int orc_pose_eval(const double* pc_corners, const double* img_norm, int F, int C,
                  const double Tcl[7], double* residuals, double* jac6)
{
    for (int f = 0; f < F; ++f)
        pose_block(Tcl, pc_corners + (size_t)f * C * 3, img_norm + (size_t)f * C * 2, C, &residuals[f],
                   jac6 ? &jac6[6 * f] : nullptr);
    return 0;
}

// include/opt/opt_reprj_calib.h:21-66
// libm_mode 0: pow / sin / cos of the solve correctly rounded (ceres_lm.hpp rn::) — the oracle's definition, reproducible on a
// GPU bit for bit; 1: this machine's libm, as a Ceres build here would call it (differs by one ulp in 1e-3 .. 1e-6 of calls; on
// few-frame problems that is enough to end the capped solve elsewhere — tests/test_oracle_restatement.py measures it).
// linear_solver 0: DENSE_NORMAL_CHOLESKY as the reference sets it; 1: DENSE_QR — the same steps up to rounding, for the sensitivity test.
// cost_trace / radius_trace (cap entries, may be null): the iteration log's `cost` and `tr_radius` columns.
int orc_pose_refine_ex(const rclc_config* cfg, const double* pc_corners, const double* img_norm, int F, int C,
                       double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations, int libm_mode, int linear_solver,
                       double* cost_trace, double* radius_trace, int32_t cap, int32_t* n_trace)
{
    const bool rn_libm = libm_mode == 0;
    orc::LMProblem P;
    P.m = F; P.n = 7; P.nl = 6;
    const double huber_a = cfg->pose_huber_a;
    P.evaluate = [&](const double* x, double* cost, double* r, double* j) {
        double total = 0;
        for (int f = 0; f < F; ++f) {
            double rr, j6[6];
            pose_block(x, pc_corners + (size_t)f * C * 3, img_norm + (size_t)f * C * 2, C, &rr, j ? j6 : nullptr);
            if (!std::isfinite(rr)) return false;
            if (j) for (int c = 0; c < 6; ++c) if (!std::isfinite(j6[c])) return false;
            total += orc::apply_loss_1d(huber_a, &rr, j ? j6 : nullptr, 6);
            if (r) r[f] = rr;
            if (j) for (int c = 0; c < 6; ++c) j[f * 6 + c] = j6[c];
        }
        *cost = total;
        return true;
    };
    P.plus = [rn_libm](const double* x, const double* d, double* out) {
        quat_plus(x, d, out, rn_libm);
        for (int k = 0; k < 3; ++k) out[4 + k] = x[4 + k] + d[3 + k];
    };
    orc::LMOptions opt;
    opt.linear_solver = linear_solver == 1 ? orc::DENSE_QR : orc::DENSE_NORMAL_CHOLESKY;      // the reference: DENSE_NORMAL_CHOLESKY (opt_reprj_calib.h:55)
    opt.max_num_iterations = cfg->pose_max_lm_iters;
    opt.function_tolerance = cfg->pose_function_tol;
    opt.correctly_rounded_libm = rn_libm;
    const orc::LMSummary sum = orc::lm_minimize(P, opt, Tcl_io);
    if (std::getenv("RCLC_ORC_DEBUG")) std::fprintf(stderr, "[orc] pose_refine: %d iterations, %d successful steps, %d cost evaluations\n", sum.num_iterations, sum.num_successful_steps, sum.num_cost_evals);
    if (initial_cost) *initial_cost = sum.initial_cost;
    if (final_cost) *final_cost = sum.final_cost;
    if (iterations) *iterations = sum.num_iterations;
    if (n_trace) *n_trace = (int32_t)sum.trace_cost.size();
    for (int32_t i = 0; i < cap && i < (int32_t)sum.trace_cost.size(); ++i) {
        if (cost_trace) cost_trace[i] = sum.trace_cost[(size_t)i];
        if (radius_trace) radius_trace[i] = sum.trace_radius[(size_t)i];
    }
    return sum.termination == orc::FAILURE ? 1 : 0;
}

int orc_pose_refine(const rclc_config* cfg, const double* pc_corners, const double* img_norm, int F, int C,
                    double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations)
{
    return orc_pose_refine_ex(cfg, pc_corners, img_norm, F, C, Tcl_io, initial_cost, final_cost, iterations, 0, 0, nullptr, nullptr, 0, nullptr);
}

// apps/Calibrate.cpp:134-148, include/PinholeCam.h:93-110
int orc_project_colorize(const rclc_config* cfg, const void* pts, int stride, int64_t n, const float T_cl[16],
                         const uint8_t* bgr, uint8_t* rgb_out, int32_t* pix_out, uint8_t* infov_out)
{
    const int W = cfg->cam_width, Hh = cfg->cam_height;
    for (int64_t k = 0; k < n; ++k) {
        const float* p = rec(pts, stride, k);
        float x = p[0], y = p[1], z = p[2];
        transform_point_f(T_cl, x, y, z);                                                 // Calibrate.cpp:134
        const double X = x, Y = y, Z = z;
        const double u = cv_round((X / Z) * cfg->fx + cfg->cx);                           // PinholeCam.h:96-98
        const double v = cv_round((Y / Z) * cfg->fy + cfg->cy);
        const bool in = !(u < 0 || v < 0 || u >= W || v >= Hh);                           // :101-110
        if (pix_out) { pix_out[2 * k] = (int32_t)u; pix_out[2 * k + 1] = (int32_t)v; }
        if (infov_out) infov_out[k] = in ? 1 : 0;
        if (in && rgb_out && bgr) {
            const uint8_t* px = bgr + ((int64_t)(int)v * W + (int)u) * 3;                 // img.at<Vec3b>(pix.y(), pix.x())
            rgb_out[3 * k + 0] = px[2];
            rgb_out[3 * k + 1] = px[1];
            rgb_out[3 * k + 2] = px[0];
        }
    }
    return 0;
}

int orc_quat_to_R(const double q[4], double R[9])
{
    quat_to_R<double>(q[0], q[1], q[2], q[3], R);
    return 0;
}

// include/utils.h:422-442
int orc_R_to_euler(const double R[9], double rpy[3])
{
    const double sy = std::sqrt(R[0] * R[0] + R[3] * R[3]);
    const bool singular = sy < 1e-6;
    if (!singular) {
        rpy[0] = std::atan2(R[7], R[8]);
        rpy[1] = std::atan2(-R[6], sy);
        rpy[2] = std::atan2(R[3], R[0]);
    } else {
        rpy[0] = std::atan2(-R[5], R[4]);
        rpy[1] = std::atan2(-R[6], sy);
        rpy[2] = 0;
    }
    return 0;
}

int orc_inverse4(const double T[16], double Tinv[16]) { return inverse4(T, Tinv) ? 0 : 1; }

// include/opt/opt_plane_param.hpp:24-88
// pc_plane_fitting (include/opt/opt_plane_param.hpp:45-88): PLANE_FITTING_COST (:24-43) through AutoDiff, one HuberLoss(0.1) per
// point, ceres::Solve with DENSE_QR and otherwise default options, from (0.1, 0.1, 1, 1).
//
// The trust-region loop is ceres_lm.hpp's (same control flow, same tests), restated here for this one problem with a DEFINED
// arithmetic: Ceres' DENSE_QR is Eigen's Householder QR on the stacked (m + 4) x 4 Jacobian, whose operation order is not pinned
// by the reference (no Ceres / Eigen version), and any reordering of those m-term dot products moves the solution in its last
// bits — which is enough to change float coordinates after pc_plane_prj and fork the grid fit downstream
// (tests/test_chain_sensitivity.py). So the oracle defines the sums over the points (cost, J'J, J'r) as exact, order-free sums of
// once-rounded terms and takes the step from the 4x4 damped normal equations by Cholesky: the same minimiser of the same
// quadratic model, reproducible bit for bit by any implementation, in any summation order.
int orc_plane_fit(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double plane_out[4],
                  int32_t* iterations, int64_t* n_used)
{
    std::vector<double> xyz;
    for (int64_t k = 0; k < n; ++k) {
        if (rec_i(pts, stride, ioff, k) < cfg->plane_reflactive_thd) continue;            // :56 (float < double)
        const float* p = rec(pts, stride, k);
        xyz.push_back(p[0]); xyz.push_back(p[1]); xyz.push_back(p[2]);
    }
    const int64_t m = (int64_t)(xyz.size() / 3);
    if (n_used) *n_used = m;
    const double huber_a = cfg->plane_huber_a;
    struct Eval { double cost, A[10], g[4]; bool valid; };
    // cost, J'J (upper triangle, row-major) and J'r at x; false on evaluation failure
    auto evaluate = [&](const double* x) {
        Eval ev;
        ExactSum acc[15];
        const double a = x[0], b = x[1], c = x[2], d = x[3];
        const double nn = (a * a + b * b) + c * c;
        const double nrm = std::sqrt(nn), den = nn * nrm;
        const double hb = huber_a * huber_a;
        bool bad = false;
        for (int64_t i = 0; i < m; ++i) {
            const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
            const double e = ((a * X + b * Y) + c * Z) + d;
            const double ae = std::fabs(e);
            double r = ae / nrm;                                                          // :35-38
            const double sg = (e < 0) ? -1.0 : 1.0;                                       // ceres abs(Jet): f.a < 0 ? -f : f
            // Jet: d|e| = sign(e) de ; d(1/nrm) = -(a,b,c)/nrm^3
            double j[4] = {(sg * X) / nrm - (ae * a) / den, (sg * Y) / nrm - (ae * b) / den, (sg * Z) / nrm - (ae * c) / den, sg / nrm};
            // IsEvaluationValid: non-finite residual or Jacobian; entries beyond 2^13 would leave the range of the exact sums and
            // count as that too (a collapsed normal, never a fit of real coordinates)
            if (!(std::fabs(r) < 8192.0 && std::fabs(j[0]) < 8192.0 && std::fabs(j[1]) < 8192.0 && std::fabs(j[2]) < 8192.0 && std::fabs(j[3]) < 8192.0)) { bad = true; continue; }
            // HuberLoss::Evaluate + Corrector (rho'' <= 0: the scaling branch), as orc::apply_loss_1d
            const double sq = r * r;
            double rho0, rho1;
            if (sq > hb) { const double rr = std::sqrt(sq); rho0 = (2.0 * huber_a) * rr - hb; rho1 = std::max(DBL_MIN, huber_a / rr); }
            else { rho0 = sq; rho1 = 1.0; }
            const double w = std::sqrt(rho1);
            r = r * w;
            for (int k = 0; k < 4; ++k) j[k] = j[k] * w;
            acc[0].add(0.5 * rho0);
            int q = 1;
            for (int u = 0; u < 4; ++u) for (int v = u; v < 4; ++v) acc[q++].add(j[u] * j[v]);
            for (int u = 0; u < 4; ++u) acc[11 + u].add(j[u] * r);
        }
        ev.cost = acc[0].value();
        for (int k = 0; k < 10; ++k) ev.A[k] = acc[1 + k].value();
        for (int k = 0; k < 4; ++k) ev.g[k] = acc[11 + k].value();
        ev.valid = !bad;
        return ev;
    };
    auto norm4 = [](const double* v) { return std::sqrt(((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]) + v[3] * v[3]); };
    // (A + diag(l^2)) y = b by Cholesky; A: upper triangle, row-major
    auto chol4 = [](const double* A10, const double* l, const double* b, double* y) {
        double M[4][4], L[4][4] = {{0}};
        int q = 0;
        for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { M[i][j] = A10[q]; M[j][i] = A10[q]; ++q; }
        for (int i = 0; i < 4; ++i) M[i][i] = M[i][i] + l[i] * l[i];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = M[i][j];
                for (int k = 0; k < j; ++k) s = s - L[i][k] * L[j][k];
                if (i == j) { if (!(s > 0.0)) return false; L[i][i] = std::sqrt(s); }
                else L[i][j] = s / L[j][j];
            }
        double z[4];
        for (int i = 0; i < 4; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s = s - L[i][k] * z[k]; z[i] = s / L[i][i]; }
        for (int i = 3; i >= 0; --i) { double s = z[i]; for (int k = i + 1; k < 4; ++k) s = s - L[k][i] * y[k]; y[i] = s / L[i][i]; }
        return std::isfinite(y[0]) && std::isfinite(y[1]) && std::isfinite(y[2]) && std::isfinite(y[3]);
    };
    const int diag_idx[4] = {0, 4, 7, 9};
    const orc::LMOptions opt;                      // :75-77 set only DENSE_QR: 50 iterations, function_tolerance 1e-6, defaults
    double x[4] = {0.1, 0.1, 1, 1}, best[4] = {0.1, 0.1, 1, 1}, cand[4];                  // :49
    double A[10], g[4], scale[4], diag[4], x_cost = 0, gmax = 0;
    auto take = [&](const Eval& ev) {
        x_cost = ev.cost;
        for (int k = 0; k < 10; ++k) A[k] = ev.A[k];
        gmax = 0;
        for (int k = 0; k < 4; ++k) { g[k] = ev.g[k]; gmax = std::max(gmax, std::fabs(x[k] - (x[k] + (-ev.g[k])))); }
    };
    int num_iterations = 0, termination = orc::NO_CONVERGENCE;
    for (int k = 0; k < 4; ++k) plane_out[k] = x[k];
    {
        const Eval e0 = evaluate(x);
        if (!e0.valid) { if (iterations) *iterations = 0; return 1; }
        take(e0);
        for (int k = 0; k < 4; ++k) scale[k] = 1.0 / (1.0 + std::sqrt(e0.A[diag_idx[k]]));      // Jacobi scaling, once
    }
    double x_norm = norm4(x), radius = opt.initial_trust_region_radius, decrease_factor = 2.0, minimum_cost = DBL_MAX;
    bool reuse_diagonal = false, step_is_successful = true;
    int iteration = 0, num_consecutive_invalid = 0;
    for (;;) {
        if (step_is_successful && x_cost < minimum_cost) { minimum_cost = x_cost; for (int k = 0; k < 4; ++k) best[k] = x[k]; }
        num_iterations++;
        if (iteration >= opt.max_num_iterations) { termination = orc::NO_CONVERGENCE; break; }
        if (step_is_successful && gmax <= opt.gradient_tolerance) { termination = orc::CONVERGENCE; break; }
        if (radius <= opt.min_trust_region_radius) { termination = orc::CONVERGENCE; break; }
        iteration++;
        step_is_successful = false;
        double As[10], bs[4], l[4], y[4];
        {
            int q = 0;
            for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { As[q] = (A[q] * scale[i]) * scale[j]; ++q; }
        }
        for (int k = 0; k < 4; ++k) bs[k] = g[k] * scale[k];
        if (!reuse_diagonal) for (int k = 0; k < 4; ++k) diag[k] = std::min(std::max(As[diag_idx[k]], opt.min_lm_diagonal), opt.max_lm_diagonal);
        for (int k = 0; k < 4; ++k) l[k] = std::sqrt(diag[k] / radius);
        bool valid = chol4(As, l, bs, y);
        reuse_diagonal = true;
        double step[4], mcc = 0;
        if (valid) {
            for (int k = 0; k < 4; ++k) step[k] = -y[k];
            double M[4][4];
            int q = 0;
            for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) { M[i][j] = As[q]; M[j][i] = As[q]; ++q; }
            double sAs = 0, sb = 0;
            for (int i = 0; i < 4; ++i) {
                double r = 0;
                for (int j = 0; j < 4; ++j) r = r + M[i][j] * step[j];
                sAs = sAs + step[i] * r;
                sb = sb + step[i] * bs[i];
            }
            mcc = -(sb + 0.5 * sAs);                                                      // -(J step)'(r + J step / 2)
            valid = mcc > 0.0;
        }
        if (!valid) {
            if (++num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) { termination = orc::FAILURE; break; }
            radius *= 0.5;
            continue;
        }
        num_consecutive_invalid = 0;
        for (int k = 0; k < 4; ++k) cand[k] = x[k] + step[k] * scale[k];
        const Eval ec = evaluate(cand);
        const double cand_cost = ec.valid ? ec.cost : DBL_MAX;
        double sn = 0;
        for (int k = 0; k < 4; ++k) sn = sn + (x[k] - cand[k]) * (x[k] - cand[k]);
        if (std::sqrt(sn) <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { termination = orc::CONVERGENCE; break; }
        if (std::fabs(x_cost - cand_cost) <= opt.function_tolerance * x_cost) { termination = orc::CONVERGENCE; break; }
        const double rel = (cand_cost >= DBL_MAX) ? -DBL_MAX : (x_cost - cand_cost) / mcc;
        if (rel > opt.min_relative_decrease) {
            for (int k = 0; k < 4; ++k) x[k] = cand[k];
            x_norm = norm4(x);
            if (!ec.valid) { termination = orc::FAILURE; break; }
            take(ec);
            step_is_successful = true;
            const double q = 2.0 * rel - 1.0;
            radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - (q * q) * q));
            decrease_factor = 2.0;
            reuse_diagonal = false;
        } else {
            radius = radius / decrease_factor;
            decrease_factor *= 2.0;
            reuse_diagonal = true;
        }
    }
    if (termination != orc::FAILURE) for (int k = 0; k < 4; ++k) plane_out[k] = best[k];   // solver.cc: only a usable solution is written back
    if (iterations) *iterations = num_iterations;
    return termination == orc::FAILURE ? 1 : 0;
}

// pcl::UniformSampling (filters/uniform_sampling.hpp, applyFilter)
int orc_uniform_sample(const void* pts, int stride, int64_t n, double leaf_d, int32_t* idx_out, int64_t* n_out)
{
    const float leaf = (float)leaf_d, inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool any = false;
    for (int64_t i = 0; i < n; ++i) {
        const float* p = rec(pts, stride, i);
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        any = true;
        for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], p[d]); mx[d] = std::max(mx[d], p[d]); }
    }
    *n_out = 0;
    if (!any) return 0;
    int64_t mb[3], div[3];
    for (int d = 0; d < 3; ++d) {
        mb[d] = (int64_t)std::floor(mn[d] * inv);
        div[d] = (int64_t)std::floor(mx[d] * inv) - mb[d] + 1;
    }
    struct Leaf { int32_t idx; float dist; };
    std::map<int64_t, Leaf> leaves;                       // ordered: the output order is ascending voxel key
    for (int64_t i = 0; i < n; ++i) {
        const float* p = rec(pts, stride, i);
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        int64_t ijk[3];
        float dist = 0.f;
        for (int d = 0; d < 3; ++d) {
            const float fl = std::floor(p[d] * inv);
            ijk[d] = (int64_t)fl;
            const float c = (fl + 0.5f) * leaf;
            const float e = p[d] - c;
            const float e2 = e * e;
            dist = (d == 0) ? e2 : dist + e2;
        }
        const int64_t key = (ijk[0] - mb[0]) + (ijk[1] - mb[1]) * div[0] + (ijk[2] - mb[2]) * div[0] * div[1];
        auto it = leaves.find(key);
        if (it == leaves.end()) leaves[key] = {(int32_t)i, dist};
        else if (dist < it->second.dist) it->second = {(int32_t)i, dist};
    }
    int64_t k = 0;
    for (auto& kv : leaves) idx_out[k++] = kv.second.idx;
    *n_out = k;
    return 0;
}

// pcl::EuclideanClusterExtraction (segmentation/extract_clusters.hpp) over KdTreeFLANN::radiusSearch
int orc_euclidean_cluster(const void* pts, int stride, int64_t n, double tol, int64_t min_size, int64_t max_size, int32_t* labels_out,
                          int32_t* n_clusters, int32_t* sizes_out)
{
    const float r2 = (float)(tol * tol);                  // kdtree_flann.hpp: radiusSearch passes static_cast<float>(radius * radius)
    const double cell = tol * 1.0001;
    std::vector<int32_t> parent((size_t)n);
    for (int64_t i = 0; i < n; ++i) parent[i] = (int32_t)i;
    auto find = [&](int32_t a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    auto finite = [&](int64_t i) { const float* p = rec(pts, stride, i); return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); };
    struct Key { int64_t x, y, z; bool operator<(const Key& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); } };
    std::map<Key, std::vector<int32_t>> grid;
    auto key_of = [&](int64_t i) { const float* p = rec(pts, stride, i); return Key{(int64_t)std::floor((double)p[0] / cell), (int64_t)std::floor((double)p[1] / cell), (int64_t)std::floor((double)p[2] / cell)}; };
    for (int64_t i = 0; i < n; ++i) if (finite(i)) grid[key_of(i)].push_back((int32_t)i);
    for (int64_t i = 0; i < n; ++i) {
        if (!finite(i)) continue;
        const float* p = rec(pts, stride, i);
        const Key k0 = key_of(i);
        for (int64_t dz = -1; dz <= 1; ++dz) for (int64_t dy = -1; dy <= 1; ++dy) for (int64_t dx = -1; dx <= 1; ++dx) {
            auto it = grid.find(Key{k0.x + dx, k0.y + dy, k0.z + dz});
            if (it == grid.end()) continue;
            for (int32_t j : it->second) {
                if (j >= i) continue;
                const float* q = rec(pts, stride, j);
                const float ex = p[0] - q[0], ey = p[1] - q[1], ez = p[2] - q[2];
                float d = ex * ex;                            // flann L2_Simple: result += diff * diff, in order
                d += ey * ey;
                d += ez * ez;
                if (d < r2) {                                 // flann RadiusResultSet::addPoint: dist < radius
                    const int32_t a = find((int32_t)i), b = find(j);
                    if (a != b) parent[std::max(a, b)] = std::min(a, b);
                }
            }
        }
    }
    std::vector<int32_t> size((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) if (finite(i)) size[find((int32_t)i)]++;
    std::vector<int32_t> roots;
    for (int64_t i = 0; i < n; ++i) if (size[i] >= min_size && size[i] <= max_size && size[i] > 0) roots.push_back((int32_t)i);
    std::stable_sort(roots.begin(), roots.end(), [&](int32_t a, int32_t b) { return size[a] != size[b] ? size[a] > size[b] : a < b; });
    std::vector<int32_t> rank((size_t)n, -1);
    for (size_t k = 0; k < roots.size(); ++k) { rank[roots[k]] = (int32_t)k; if (sizes_out) sizes_out[k] = size[roots[k]]; }
    for (int64_t i = 0; i < n; ++i) labels_out[i] = finite(i) ? rank[find((int32_t)i)] : -1;
    *n_clusters = (int32_t)roots.size();
    return 0;
}

// utils.h:160-202, pc_process.h:373-462, opt/opt_line_fitting.h:17-108
int orc_calc_laser_coor(const rclc_config* cfg, const double* cen, int32_t n, const double plane[4], double Tbl[16])
{
    struct V3 { double x, y, z; };
    auto at = [&](int i) { return V3{cen[3 * i], cen[3 * i + 1], cen[3 * i + 2]}; };
    auto sub = [](V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; };
    auto cross = [](V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; };
    auto norm = [](V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); };
    // get_ref_pt: the farthest centroid of the (x > cx, y <= cy) quadrant
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
    std::vector<std::pair<double, int>> dl;
    for (int i = 0; i < n; ++i) if (i != ref_idx) dl.push_back({norm(sub(at(i), ref0)), i});
    std::sort(dl.begin(), dl.end());
    const V3 ref1 = at(dl[1].second), ref2 = at(dl[2].second);
    V3 xt = (ref1.x < ref2.x && ref1.y < ref2.y) ? sub(ref0, ref1) : sub(ref0, ref2);
    { const double l = norm(xt); xt = V3{xt.x / l, xt.y / l, xt.z / l}; }
    std::vector<std::pair<double, int>> rows;
    for (int i = 0; i < n; ++i) rows.push_back({norm(cross(sub(ref0, at(i)), xt)), i});
    std::sort(rows.begin(), rows.end());
    // line_fitting: one residual per board row, sum over the row's points of |(p - row centroid) x normalize(abc)|, Huber(0.1)
    const int per_row = cfg->board_width / 2, n_rows = cfg->board_height, origin_id = cfg->board_height / 2;
    if (per_row * n_rows > n) return 1;
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
    orc::LMProblem P;
    P.m = n_rows; P.n = 3; P.nl = 3;
    P.evaluate = [&](const double* x, double* cost, double* res, double* jac) {
        const double na = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        const V3 u{x[0] / na, x[1] / na, x[2] / na};
        double total = 0;
        for (int r = 0; r < n_rows; ++r) {
            double rr = 0, g[3] = {0, 0, 0};
            for (const V3& p : rp[(size_t)r]) {
                const V3 d = sub(p, rc[(size_t)r]);
                const V3 cr = cross(d, u);
                const double l = norm(cr);
                rr += l;
                if (jac) { const V3 ch{cr.x / l, cr.y / l, cr.z / l}; const V3 q = cross(ch, d); g[0] += q.x; g[1] += q.y; g[2] += q.z; }      // d|d x u|/du
            }
            double j3[3];
            if (jac) {
                const double gu = g[0] * u.x + g[1] * u.y + g[2] * u.z;                    // du/da = (I - u u^T) / |a|
                j3[0] = (g[0] - gu * u.x) / na; j3[1] = (g[1] - gu * u.y) / na; j3[2] = (g[2] - gu * u.z) / na;
            }
            if (!std::isfinite(rr)) return false;
            total += orc::apply_loss_1d(0.1, &rr, jac ? j3 : nullptr, 3);
            if (res) res[r] = rr;
            if (jac) for (int q = 0; q < 3; ++q) jac[r * 3 + q] = j3[q];
        }
        *cost = total;
        return true;
    };
    orc::LMOptions opt;
    opt.linear_solver = orc::DENSE_QR;
    orc::lm_minimize(P, opt, abc);
    V3 bx{abc[0], abc[1], abc[2]};
    { const double l = norm(bx); bx = V3{bx.x / l, bx.y / l, bx.z / l}; }
    V3 bz{plane[0], plane[1], plane[2]};
    { const double l = norm(bz); bz = V3{bz.x / l, bz.y / l, bz.z / l}; }
    V3 by = cross(bx, bz);
    if (bx.x < 0) bx = V3{-bx.x, -bx.y, -bx.z};
    if (by.y < 0) by = V3{-by.x, -by.y, -by.z};
    if (bz.z < 0) bz = V3{-bz.x, -bz.y, -bz.z};
    const double s = cfg->gride_side_length;
    const V3 ot = (cfg->board_height % 2 != 0) ? V3{-s * 0.5, -s * 0.5, 0} : V3{s * 0.5, -s * 0.5, 0};
    const V3 o{bx.x * ot.x + by.x * ot.y + bz.x * ot.z + origin_centroid.x, bx.y * ot.x + by.y * ot.y + bz.y * ot.z + origin_centroid.y,
               bx.z * ot.x + by.z * ot.y + bz.z * ot.z + origin_centroid.z};
    const V3 ax[3] = {bx, by, bz};
    for (int r = 0; r < 3; ++r) {
        Tbl[4 * r] = ax[r].x; Tbl[4 * r + 1] = ax[r].y; Tbl[4 * r + 2] = ax[r].z;
        Tbl[4 * r + 3] = -(ax[r].x * o.x + ax[r].y * o.y + ax[r].z * o.z);
    }
    Tbl[12] = 0; Tbl[13] = 0; Tbl[14] = 0; Tbl[15] = 1;
    return 0;
}

// The two problems of the Ceres tutorial whose solver logs are printed in the Ceres documentation
// (docs/source/nnls_tutorial.rst: "Hello World!" f(x) = 10 - x from x = 0.5, and Powell's function
// from (3, -1, 0, 1); DENSE_QR there, one residual per block like every block of the reference;
// linear_solver 1 runs the same problem through the DENSE_NORMAL_CHOLESKY branch the pose refinement uses).
// Runs the restated minimiser on them and hands back the per-iteration `cost` and `tr_radius`
// columns, which tests/test_oracle_ceres_logs.py lays beside the published ones.
int orc_lm_tutorial(int which, int linear_solver, double* x_io, int max_iterations, double* cost_trace, double* radius_trace, int32_t cap,
                    int32_t* n_trace, int32_t* termination)
{
    orc::LMProblem P;
    if (which == 0) {
        P.m = 1; P.n = 1; P.nl = 1;
        P.evaluate = [](const double* x, double* cost, double* res, double* jac) {
            const double r = 10.0 - x[0];
            if (res) res[0] = r;
            if (jac) jac[0] = -1.0;
            *cost = 0.5 * r * r;
            return true;
        };
    } else if (which == 1) {
        P.m = 4; P.n = 4; P.nl = 4;
        P.evaluate = [](const double* x, double* cost, double* res, double* jac) {
            const double s5 = std::sqrt(5.0), s10 = std::sqrt(10.0);
            const double f[4] = {x[0] + 10.0 * x[1], s5 * (x[2] - x[3]), (x[1] - 2.0 * x[2]) * (x[1] - 2.0 * x[2]),
                                 s10 * (x[0] - x[3]) * (x[0] - x[3])};
            if (res) for (int i = 0; i < 4; ++i) res[i] = f[i];
            if (jac) {
                for (int i = 0; i < 16; ++i) jac[i] = 0.0;
                jac[0] = 1.0; jac[1] = 10.0;
                jac[4 + 2] = s5; jac[4 + 3] = -s5;
                jac[8 + 1] = 2.0 * (x[1] - 2.0 * x[2]); jac[8 + 2] = -4.0 * (x[1] - 2.0 * x[2]);
                jac[12 + 0] = 2.0 * s10 * (x[0] - x[3]); jac[12 + 3] = -2.0 * s10 * (x[0] - x[3]);
            }
            *cost = 0.5 * (f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3]);
            return true;
        };
    } else {
        return 1;
    }
    orc::LMOptions opt;
    opt.linear_solver = linear_solver == 1 ? orc::DENSE_NORMAL_CHOLESKY : orc::DENSE_QR;
    opt.max_num_iterations = max_iterations;
    const orc::LMSummary sum = orc::lm_minimize(P, opt, x_io);
    const int32_t n = (int32_t)std::min<size_t>(sum.trace_cost.size(), (size_t)cap);
    for (int32_t i = 0; i < n; ++i) { cost_trace[i] = sum.trace_cost[(size_t)i]; radius_trace[i] = sum.trace_radius[(size_t)i]; }
    *n_trace = (int32_t)sum.trace_cost.size();
    *termination = sum.termination;
    return 0;
}

// the correctly rounded stand-ins of ceres_lm.hpp, for tests/test_oracle_restatement.py (which holds them to exact rational arithmetic)
double orc_rn_cube(double q) { return orc::rn::cube(q); }
int orc_rn_sincos(double x, double* sin_out, double* cos_out) { orc::rn::sincos(x, sin_out, cos_out); return 0; }

} // extern "C"
