// pose.cu — K6: final extrinsic refinement (one scalar residual per frame, 7 parameters / 6 DoF).
//
// Replaces POSE_REPRJ_COST::operator() (include/opt/pose_reprj_cost.hpp:28-55, evaluated through
// ceres::AutoDiffCostFunction<.,1,4,3>) and pose_reprj_opt (include/opt/opt_reprj_calib.h:21-66:
// HuberLoss(0.2), EigenQuaternionParameterization, DENSE_NORMAL_CHOLESKY, 1000 iterations,
// function_tolerance 1e-15). The whole solve is ONE kernel launch — this path is latency-bound.
//
// THE ARITHMETIC IS THE ORACLE'S, OPERATION BY OPERATION. On few-frame problems this solve never converges inside its
// iteration cap, and where it stands after 1000 iterations follows the last bit of every step (the oracle itself moves by
// 0.1 - 6 mm when its Cholesky or its FMA contraction is changed, tests/test_pose_sensitivity.py). So nothing here is left to the
// compiler or to a reduction tree: this file is built with -fmad=false (rclc_b200/build.py), every expression is written as
// the CPU oracle writes it (oracle/: the residual on Jets, ceres_lm.hpp), sums over frames run in frame order, and pow(q, 3), sin and cos are taken
// correctly rounded (as the oracle defines them). The result is the oracle's iterate sequence bit for bit — same iteration count,
// same costs, same extrinsic (tests/test_gpu_parity.py::test_pose_refine_is_the_oracles_bit_for_bit).
//
// Shape of an iteration: the frames' residuals and Jacobians are evaluated in parallel (7-wide dual numbers like Jets; one thread
// per frame in the one-CTA shape; (frame, corner) pairs dealt over a 16-CTA thread-block cluster from 64 corner observations up)
// and leave ONE ROW per frame — 0.5 rho, the corrected residual, the corrected local Jacobian — in the leading CTA's shared
// memory (pushed over DSMEM). The BOSS WARP (warp 0 of the leading CTA) then forms J'J, J'r and the cost with one lane per sum,
// frame after frame, judges the step and computes the next one (Cholesky, model cost change, Plus) with the same state in
// every lane; the AIDE WARP (warp 1) takes the gradient check's Plus(x, -g) off its path. The candidate and the stop word reach
// everybody behind one barrier (a cluster barrier in the cluster shape): two cluster barriers per iteration.
#include "ops_dev.cuh"

#include <cfloat>
#include <cstdlib>
#include <cstdio>
#include <cooperative_groups.h>

namespace rclc {

namespace cg = cooperative_groups;

struct J7 {
    double a;
    double v[7];
};
__device__ __forceinline__ J7 jconst(double s) { J7 r; r.a = s; for (int i = 0; i < 7; ++i) r.v[i] = 0; return r; }
__device__ __forceinline__ J7 jvar(double s, int k) { J7 r = jconst(s); r.v[k] = 1.0; return r; }
__device__ __forceinline__ J7 operator+(const J7& f, const J7& g) { J7 h; h.a = f.a + g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
__device__ __forceinline__ J7 operator-(const J7& f, const J7& g) { J7 h; h.a = f.a - g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
__device__ __forceinline__ J7 operator*(const J7& f, const J7& g) { J7 h; h.a = f.a * g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
__device__ __forceinline__ J7 operator/(const J7& f, const J7& g)
{
    J7 h;
    const double gi = 1.0 / g.a, fg = f.a * gi;
    h.a = fg;
    for (int i = 0; i < 7; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
    return h;
}
__device__ __forceinline__ J7 jsqrt(const J7& f)
{
    J7 h;
    const double t = sqrt(f.a), ti = 1.0 / (2.0 * t);
    h.a = t;
    for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] * ti;
    return h;
}

// pose_reprj_cost.hpp:28-55 on Jets. q = (x,y,z,w) storage, Eigen::Quaternion{w,x,y,z}.matrix().
// r: residual; j6: 1x6 local Jacobian (AutoDiff row 1x7 times EigenQuaternionParameterization plus-Jacobian).
__device__ void pose_block_dev(const double* Tcl, const double* pc, const double* img, int C, double& r, double* j6)
{
    J7 q[4], t[3];
    for (int k = 0; k < 4; ++k) q[k] = jvar(Tcl[k], k);
    for (int k = 0; k < 3; ++k) t[k] = jvar(Tcl[4 + k], 4 + k);
    const J7 two = jconst(2.0), one = jconst(1.0);
    const J7 &x = q[0], &y = q[1], &z = q[2], &w = q[3];
    const J7 tx = two * x, ty = two * y, tz = two * z;
    const J7 twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    const J7 R0 = one - (tyy + tzz), R1 = txy - twz, R2 = txz + twy;
    const J7 R3 = txy + twz, R4 = one - (txx + tzz), R5 = tyz - twx;
    const J7 R6 = txz - twy, R7 = tyz + twx, R8 = one - (txx + tyy);
    J7 res = jconst(0.0), depth_max = jconst(0.0);
    for (int i = 0; i < C; ++i) {
        const J7 p0 = jconst(pc[3 * i]), p1 = jconst(pc[3 * i + 1]), p2 = jconst(pc[3 * i + 2]);
        const J7 X = (R0 * p0 + (R1 * p1 + R2 * p2)) + t[0];
        const J7 Y = (R3 * p0 + (R4 * p1 + R5 * p2)) + t[1];
        const J7 Z = (R6 * p0 + (R7 * p1 + R8 * p2)) + t[2];
        const J7 depth_norm = jsqrt(X * X + (Y * Y + Z * Z));
        if (depth_norm.a > depth_max.a) depth_max = depth_norm;
        const J7 e0 = jconst(img[2 * i]) - X / Z, e1 = jconst(img[2 * i + 1]) - Y / Z;
        res = res + jsqrt(e0 * e0 + e1 * e1) / depth_norm;
    }
    res = res * depth_max;
    r = res.a;
    // EigenQuaternionParameterization::ComputeJacobian (4x3), local Jacobian = J_global * plus_jacobian
    const double P[12] = {Tcl[3], Tcl[2], -Tcl[1], -Tcl[2], Tcl[3], Tcl[0], Tcl[1], -Tcl[0], Tcl[3], -Tcl[0], -Tcl[1], -Tcl[2]};
    for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += res.v[k] * P[k * 3 + c];
        j6[c] = s;
    }
    for (int c = 0; c < 3; ++c) j6[3 + c] = res.v[4 + c];
}

// ---- correctly rounded pow(q, 3), sin and cos: oracle/ceres_lm.hpp namespace rn, the same operations ----
struct dd { double h, l; };
__device__ __forceinline__ dd dd_norm(double s, double e) { const double h = s + e; return {h, e - (h - s)}; }
__device__ __forceinline__ dd dd_mul(dd a, dd b) { const double p = a.h * b.h; double e = fma(a.h, b.h, -p); e = e + (a.h * b.l + a.l * b.h); return dd_norm(p, e); }
__device__ __forceinline__ dd dd_add(dd a, dd b) { const double s = a.h + b.h, v = s - a.h; double e = (a.h - (s - v)) + (b.h - v); e = e + (a.l + b.l); return dd_norm(s, e); }
__device__ __forceinline__ double cube_rn(double q)
{
    const double q2h = q * q, q2l = fma(q, q, -q2h), ph = q2h * q, pl = fma(q2h, q, -ph);
    return ph + (pl + q2l * q);
}
__constant__ double c_sin_dd[15][2] = {
    {0x1.0000000000000p+0, 0x0.0p+0}, {-0x1.5555555555555p-3, -0x1.5555555555555p-57}, {0x1.1111111111111p-7, 0x1.1111111111111p-63},
    {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73}, {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80},
    {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87}, {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
    {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120}, {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130},
    {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139}, {-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157}};
__constant__ double c_cos_dd[15][2] = {
    {0x1.0000000000000p+0, 0x0.0p+0}, {-0x1.0000000000000p-1, 0x0.0p+0}, {0x1.5555555555555p-5, 0x1.5555555555555p-59},
    {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65}, {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76}, {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76},
    {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83}, {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
    {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120}, {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124},
    {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135}, {-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153}};
__device__ void sincos_rn(double x, double* sn, double* cs)
{
    if (!(x <= 0.5)) { *sn = sin(x); *cs = cos(x); return; }          // beyond the series' range: the library's (1-2 ulp) values
    const double uh = x * x;
    const dd u = {uh, fma(x, x, -uh)};
    const int n = uh < 0x1p-36 ? 3 : (uh < 0x1p-20 ? 5 : (uh < 0x1p-8 ? 9 : 14));
    dd S = {c_sin_dd[n][0], c_sin_dd[n][1]}, Cc = {c_cos_dd[n][0], c_cos_dd[n][1]};
    for (int k = n - 1; k >= 0; --k) {
        S = dd_add(dd_mul(S, u), dd{c_sin_dd[k][0], c_sin_dd[k][1]});
        Cc = dd_add(dd_mul(Cc, u), dd{c_cos_dd[k][0], c_cos_dd[k][1]});
    }
    const dd xs = dd_mul(S, dd{x, 0.0});
    *sn = xs.h + xs.l;
    *cs = Cc.h + Cc.l;
}

__device__ void quat_plus_dev(const double* x, const double* delta, double* out)
{
    const double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (nd > 0.0) {
        double sin_nd, cos_nd;
        sincos_rn(nd, &sin_nd, &cos_nd);
        const double s = sin_nd / nd;
        const double dw = cos_nd, dx = s * delta[0], dy = s * delta[1], dz = s * delta[2];
        const double xw = x[3], xx = x[0], xy = x[1], xz = x[2];
        out[3] = dw * xw - dx * xx - dy * xy - dz * xz;
        out[0] = dw * xx + dx * xw + dy * xz - dz * xy;
        out[1] = dw * xy + dy * xw + dz * xx - dx * xz;
        out[2] = dw * xz + dz * xw + dx * xy - dy * xx;
    } else {
        for (int i = 0; i < 4; ++i) out[i] = x[i];
    }
}
__device__ void pose_plus(const double* x, const double* d, double* out)
{
    quat_plus_dev(x, d, out);
    for (int k = 0; k < 3; ++k) out[4 + k] = x[4 + k] + d[3 + k];
}

constexpr int kPoseThreads = 256;
struct PoseOut {
    double Tcl[7]; double initial_cost, final_cost; int iterations; int termination;
#ifdef RCLC_POSE_TIMING
    long long t_step, t_pub, t_eval, t_decide;       // thread 0's clock64 in the four parts of an iteration (experiments only)
#endif
};
#ifdef RCLC_POSE_TIMING
#define POSE_TICK(acc) do { const long long now_ = clock64(); acc += now_ - tick_; tick_ = now_; } while (0)
#else
#define POSE_TICK(acc) do { } while (0)
#endif

// A frame's row: ResidualBlock::Evaluate's tail for a one-residual block under HuberLoss (oracle/ceres_lm.hpp apply_loss_1d) —
// 0.5 rho(r^2), the corrected residual, the corrected local Jacobian. A non-finite residual or Jacobian stays non-finite in the
// row and so in the boss's sums, which is where it is noticed.
constexpr int kRow = 8;
__device__ __forceinline__ void pose_row(double r, const double* j, double huber_a, double* row)
{
    const double sq = r * r, bb = huber_a * huber_a;
    double rho0, rho1;
    if (sq > bb) { const double rr = sqrt(sq); rho0 = 2.0 * huber_a * rr - bb; rho1 = fmax(DBL_MIN, huber_a / rr); }
    else { rho0 = sq; rho1 = 1.0; }
    const double w = sqrt(rho1);
    row[0] = 0.5 * rho0;
    row[1] = r * w;
    for (int c = 0; c < 6; ++c) row[2 + c] = j[c] * w;
}

// Evaluate all frames at s_x, one thread per frame: rows[f][kRow]
__device__ void pose_evaluate(const double* pc, const double* img, int F, int C, double huber_a, const double* s_x, double* rows)
{
    for (int f = threadIdx.x; f < F; f += kPoseThreads) {
        double r, j[6], row[kRow];
        pose_block_dev(s_x, pc + (size_t)f * C * 3, img + (size_t)f * C * 2, C, r, j);
        pose_row(r, j, huber_a, row);
        for (int k = 0; k < kRow; ++k) rows[(size_t)f * kRow + k] = row[k];
    }
    __syncthreads();
}

// ---- the cluster shape of the evaluation ----
// q, t -> the nine rotation entries as duals (pose_block_dev's prologue, same expressions)
__device__ __forceinline__ void pose_rot_dev(const double* Tcl, J7* R, J7* t)
{
    J7 q[4];
    for (int k = 0; k < 4; ++k) q[k] = jvar(Tcl[k], k);
    for (int k = 0; k < 3; ++k) t[k] = jvar(Tcl[4 + k], 4 + k);
    const J7 two = jconst(2.0), one = jconst(1.0);
    const J7 &x = q[0], &y = q[1], &z = q[2], &w = q[3];
    const J7 tx = two * x, ty = two * y, tz = two * z;
    const J7 twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = one - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = one - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = one - (txx + tyy);
}

#ifdef RCLC_POSE_TIMING
__device__ long long g_pose_ticks[8];
__device__ long long g_step_ticks[8];
#define STEP_TICK(k) do { const long long now_ = clock64(); if (threadIdx.x == 0) g_step_ticks[k] += now_ - stick_; stick_ = now_; } while (0)
#else
#define STEP_TICK(k) do { } while (0)
#endif
#ifdef RCLC_POSE_TIMING
#define EVAL_TICK(k) do { if (threadIdx.x == 0 && rank == 0) { const long long now_ = clock64(); g_pose_ticks[k] += now_ - etick_; etick_ = now_; } } while (0)
#else
#define EVAL_TICK(k) do { } while (0)
#endif
// One corner of pose_block_dev's loop, stopped at the operands of "res + |e| / depth_norm": the quotient's value fg = |e|.a / d.a,
// 1 / d.a and the dual part's numerators |e|.v - fg * d.v travel, the frame's thread multiplies and adds (Jet operator/ and
// operator+ cut in two: the same operations on the same operands).
// out (component stride os): fg, 1/d.a, (|e|.v - fg * d.v)[7], d.a, d.v[7].
constexpr int kCornerVals = 17;
// stride between a corner's factors in shared memory: odd, so that the eight lanes of a frame (a factor each) fall on different banks
__host__ __device__ constexpr int corner_stride(int S, int C) { return (S * C) | 1; }
__device__ __forceinline__ void pose_corner_dev(const J7* R, const J7* t, const double* pc3, const double* img2, double* out, int os)
{
    const J7 p0 = jconst(pc3[0]), p1 = jconst(pc3[1]), p2 = jconst(pc3[2]);
    const J7 X = (R[0] * p0 + (R[1] * p1 + R[2] * p2)) + t[0];
    const J7 Y = (R[3] * p0 + (R[4] * p1 + R[5] * p2)) + t[1];
    const J7 Z = (R[6] * p0 + (R[7] * p1 + R[8] * p2)) + t[2];
    const J7 depth_norm = jsqrt(X * X + (Y * Y + Z * Z));
    const J7 e0 = jconst(img2[0]) - X / Z, e1 = jconst(img2[1]) - Y / Z;
    const J7 en = jsqrt(e0 * e0 + e1 * e1);
    const double gi = 1.0 / depth_norm.a, fg = en.a * gi;
    out[0] = fg;
    out[os] = gi;
    for (int i = 0; i < 7; ++i) out[(2 + i) * os] = en.v[i] - fg * depth_norm.v[i];
    out[9 * os] = depth_norm.a;
    for (int i = 0; i < 7; ++i) out[(10 + i) * os] = depth_norm.v[i];
}

// One frame from its corners' factors, in corner order: pose_block_dev's running sum, arg-max and epilogue - by ONE thread (the
// form rclc_pose_eval's split test hook runs; the solve uses pose_frame_phase, eight lanes per frame).
__device__ __forceinline__ void pose_frame_from_corners(const double* Tcl, const double* c0, int os, int C, double& r, double* j6)
{
    J7 res = jconst(0.0), depth_max = jconst(0.0);
#pragma unroll 5
    for (int i = 0; i < C; ++i) {
        const double* o = c0 + i;
        double ld[kCornerVals];
#pragma unroll
        for (int k = 0; k < kCornerVals; ++k) ld[k] = o[k * os];      // all loads up front, the arg-max as selects: no branch
        const bool farther = ld[9] > depth_max.a;
        depth_max.a = farther ? ld[9] : depth_max.a;
#pragma unroll
        for (int k = 0; k < 7; ++k) depth_max.v[k] = farther ? ld[10 + k] : depth_max.v[k];
        res.a = res.a + ld[0];
#pragma unroll
        for (int k = 0; k < 7; ++k) res.v[k] = res.v[k] + ld[2 + k] * ld[1];
    }
    res = res * depth_max;
    r = res.a;
    const double P[12] = {Tcl[3], Tcl[2], -Tcl[1], -Tcl[2], Tcl[3], Tcl[0], Tcl[1], -Tcl[0], Tcl[3], -Tcl[0], -Tcl[1], -Tcl[2]};
    for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += res.v[k] * P[k * 3 + c];
        j6[c] = s;
    }
    for (int c = 0; c < 3; ++c) j6[3 + c] = res.v[4 + c];
}

// The two phases are separate routines so that the nine rotation duals (144 registers) are not live while a frame is assembled.
__device__ __noinline__ void pose_corners_phase(const double* Tcl, const double* pc, const double* img, int f0, int fstep, int C, int ntask,
                                                double* s_corner, int os)
{
    J7 R[9], t[3];
    pose_rot_dev(Tcl, R, t);
    for (int tsk = threadIdx.x; tsk < ntask; tsk += kPoseThreads) {        // task = (this CTA's l-th frame f0 + l * fstep, corner c)
        const int l = tsk / C, c = tsk - l * C;
        const size_t fc = (size_t)(f0 + l * fstep) * C + c;
        pose_corner_dev(R, t, pc + fc * 3, img + fc * 2, s_corner + tsk, os);
    }
}

// One frame from its corners' factors by EIGHT lanes, lane `comp` carrying one component of the duals (0: the value, k + 1: the
// derivative k) through pose_frame_from_corners' running sum, arg-max and epilogue - the same operations on the same operands, a
// component each - then all eight form the frame's row and lane `comp` stores entry `comp`. (One thread per frame spends 4.3 k
// cycles on its 35 x 17 shared-memory loads; eight take 4 each per corner.) Called by whole warps; `store` is false for the lanes
// that only keep a warp complete.
__device__ __noinline__ void pose_frame_phase(const double* Tcl, const double* c0, int os, int C, double huber_a, double* dst_row, int comp,
                                              bool store)
{
    constexpr unsigned kAll = 0xffffffffu;
#ifdef RCLC_POSE_TIMING
    long long ft_ = clock64();
#define FRAME_TICK(k) do { const long long n_ = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) g_pose_ticks[k] += n_ - ft_; ft_ = n_; } while (0)
#else
#define FRAME_TICK(k) do { } while (0)
#endif
    const int i_add = comp == 0 ? 0 : (1 + comp) * os, i_dc = (9 + comp) * os;      // comp 0: fg | d.a; k + 1: numerator k | d.v[k]
    // depth_max: the running "if (d > depth_max) depth_max = d" ends on the FIRST corner that attains the largest norm (nothing, if
    // none exceeds 0). Found here without the 35-long chain: every lane scans a stride of the corners, three exchanges settle it
    // (larger norm wins, the smaller index on a tie). The norms are >= +0, so their bit patterns order like their values and
    // the comparisons are integer ones (a NaN norm wins here and loses there; the frame's residual is NaN either way).
    long long best = 0;                                             // bits of +0.0
    int bi = -1;
    for (int i = comp; i < C; i += 8) {
        const long long da = __double_as_longlong(c0[9 * os + i]);
        if (da > best) { best = da; bi = i; }
    }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
        const long long ob = __shfl_xor_sync(kAll, best, o, 8);
        const int oi = __shfl_xor_sync(kAll, bi, o, 8);
        const bool take = ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi));
        best = take ? ob : best;
        bi = take ? oi : bi;
    }
    const double dmax_a = __longlong_as_double(best), dmax_c = bi >= 0 ? c0[i_dc + bi] : 0.0;
    double res = 0.0;
#pragma unroll 5
    for (int i = 0; i < C; ++i) {
        const double* o = c0 + i;
        const double num = o[i_add], gi = comp == 0 ? 1.0 : o[os];
        res = res + num * gi;                                       // comp 0: res.a + fg (times 1.0: exact)
    }
    FRAME_TICK(6);
    // res = res * depth_max (Jet product)
    const double res_a = __shfl_sync(kAll, res, 0, 8);
    const double h = comp == 0 ? res * dmax_a : res_a * dmax_c + res * dmax_a;
    double v[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = __shfl_sync(kAll, h, k + 1, 8);
    const double r = __shfl_sync(kAll, h, 0, 8);
    const double P[12] = {Tcl[3], Tcl[2], -Tcl[1], -Tcl[2], Tcl[3], Tcl[0], Tcl[1], -Tcl[0], Tcl[3], -Tcl[0], -Tcl[1], -Tcl[2]};
    double j[6], row[kRow];
    for (int c = 0; c < 3; ++c) {
        double t = 0;
        for (int k = 0; k < 4; ++k) t += v[k] * P[k * 3 + c];
        j[c] = t;
    }
    for (int c = 0; c < 3; ++c) j[3 + c] = v[4 + c];
    FRAME_TICK(7);
    pose_row(r, j, huber_a, row);
    FRAME_TICK(0);
    double mine = row[7];
#pragma unroll
    for (int k = 6; k >= 0; --k) mine = comp == k ? row[k] : mine;
    if (store) dst_row[comp] = mine;
}

// pose_evaluate over a cluster: frames are dealt round-robin over the CTAs (frame f to rank f % n_cta, so that a handful of frames
// still spreads out), 16 at a time per CTA; the rows go to `rows` — rank 0's shared memory, or global memory — and are complete
// behind the cluster barrier. s_corner: [kCornerVals][corner_stride(S, C)] doubles per CTA.
__device__ void pose_evaluate_cluster(cg::cluster_group& cluster, const double* pc, const double* img, int F, int C, double huber_a,
                                      const double* s_x, double* s_corner, double* rows, int mode)
{
    const int tid = threadIdx.x, rank = (int)cluster.block_rank(), n_cta = (int)cluster.num_blocks();
    const int S = kPoseThreads / n_cta, os = corner_stride(S, C);
#ifdef RCLC_POSE_TIMING
    long long etick_ = clock64();
#endif
    for (int base = 0; base < F; base += kPoseThreads) {
        const int nfb = min(F - base, kPoseThreads);             // frames of this sweep
        const int nfr = nfb > rank ? (nfb - rank + n_cta - 1) / n_cta : 0, ntask = mode == 1 ? 0 : nfr * C;   // mine: base + rank + l * n_cta
        if (tid < ntask) pose_corners_phase(s_x, pc, img, base + rank, n_cta, C, ntask, s_corner, os);
        EVAL_TICK(1);
        __syncthreads();
        EVAL_TICK(2);
        if (mode == 1) {                                          // test hook: the whole frame by the one-CTA kernel's routine
            if (tid < nfr) {
                const size_t f = (size_t)(base + rank + tid * n_cta);
                double r, j[6], row[kRow];
                pose_block_dev(s_x, pc + f * C * 3, img + f * C * 2, C, r, j);
                pose_row(r, j, huber_a, row);
                for (int k = 0; k < kRow; ++k) rows[f * kRow + k] = row[k];
            }
        } else if (tid < ((nfr * 8 + 31) & ~31)) {                // eight lanes per frame, whole warps
            const int l = min(tid >> 3, nfr - 1);
            const size_t f = (size_t)(base + rank + l * n_cta);
            pose_frame_phase(s_x, s_corner + l * C, os, C, huber_a, rows + f * kRow, tid & 7, (tid >> 3) < nfr);
        }
        EVAL_TICK(3);
        __syncthreads();
    }
    EVAL_TICK(4);
    cluster.sync();                                               // every frame's row is in place
    EVAL_TICK(5);
}

__device__ __forceinline__ constexpr int sym6(int a, int b) { return a <= b ? a * 6 - a * (a - 1) / 2 + (b - a) : b * 6 - b * (b - 1) / 2 + (a - b); }

// Eigen's LLT and the two substitutions as oracle/ceres_lm.hpp normal_cholesky_solve restates them - every element by the oracle's
// operations in the oracle's order - with the work of a column dealt over the lanes of the (converged) warp: lane r < 6 owns row r
// of L, so a column's quotients are taken side by side instead of in a row, and lane 6 carries the right-hand side as a seventh
// row: its "L[6][j]" = (b[j] - sum_k L[j][k] y[k]) / L[j][j] IS the forward substitution, term for term. Lanes whose quotient is
// not needed divide 0 by the pivot (a quotient of leftovers can send the whole warp down the division's slow path: 8 k cycles
// per solve instead of 3 k). U: packed upper triangle of the symmetric matrix (sym6), add: what goes onto its diagonal; every lane
// gets the whole solution. false: a pivot is not positive, or the solution not finite.
__device__ __forceinline__ bool chol6_solve_warp(const double* U, const double* add, const double* b, double* y, int lane)
{
    constexpr unsigned kAll = 0xffffffffu;
    const int r = min(lane, 6);
    double Ar[6], Lr[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {                                  // my row (entries right of the diagonal are not used)
        double v = b[c];
#pragma unroll
        for (int rr = 5; rr >= 0; --rr) v = r == rr ? (rr == c ? U[sym6(rr, rr)] + add[rr] : U[sym6(rr, c)]) : v;
        Ar[c] = v;
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = Ar[j];                                          // on lane j: the pivot
#pragma unroll
        for (int k = 0; k < j; ++k) d -= Lr[k] * Lr[k];
        d = __shfl_sync(kAll, d, j);
        ok = ok && (d > 0.0);
        d = sqrt(d);
        double t = Ar[j];                                          // on lane i > j: L[i][j]; on lane 6: y[j] of the forward substitution
#pragma unroll
        for (int k = 0; k < j; ++k) t -= Lr[k] * __shfl_sync(kAll, Lr[k], j);
        t = r > j ? t : 1.0;                                       // (not 0: a zero operand is the slow path too)
        Lr[j] = r == j ? d : t / d;
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = __shfl_sync(kAll, Lr[i], 6);
#pragma unroll
    for (int i = 5; i >= 0; --i) {                                 // backward: column i of L sits in the lanes below the diagonal
        double t = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) t -= __shfl_sync(kAll, Lr[i], k) * y[k];
        y[i] = t / __shfl_sync(kAll, Lr[i], i);
    }
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 6; ++i) fin = fin && isfinite(y[i]);
    return fin;
}

// Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseNormalCholeskySolver as oracle/ceres_lm.hpp restates them, in
// the oracle's operations and order (see the head of this file). kCluster: the evaluation is shared by the CTAs of a cluster.
// kGlobalRows / g_rows: the frames' rows in global memory, when 2 x F x kRow doubles do not fit shared memory (a template parameter,
// so that the boss's loops over the rows are shared-memory loads in the usual case, not generic ones).
template <bool kCluster, bool kGlobalRows>
__device__ __forceinline__ void pose_refine_body(const double* __restrict__ pc, const double* __restrict__ img, int F, int C,
                                                 double huber_a, double ftol, int max_iters, const double* __restrict__ T0,
                                                 PoseOut* __restrict__ out, double* g_rows, int mode)
{
    __shared__ double s_x[7], s_cand[7], s_gmax;
    __shared__ int s_ctrl[2];        // bit 0: stop; bit 1: the row buffer the next evaluation fills. Two slots used in turn, so that
    int pub = 0;                     // the boss's next word never lands on one some thread has yet to read
    extern __shared__ double s_dyn[];                 // [cluster: per-corner factors] [rows: 2 x F x kRow, unless g_rows]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = kCluster ? (int)cluster.block_rank() : 0;
    const int lane = (int)threadIdx.x & 31;
    const bool boss = rank == 0 && threadIdx.x < 32, aide = rank == 0 && (threadIdx.x >> 5) == 1;
    double* s_corner = s_dyn;
    double* rows_here = s_dyn + (kCluster ? (size_t)kCornerVals * corner_stride(kPoseThreads / (int)cluster.num_blocks(), C) : 0);
    double* rows_lead = rows_here;                    // the leading CTA's rows as this CTA addresses them
    if constexpr (kCluster) rows_lead = cluster.map_shared_rank(rows_here, 0);
    if constexpr (kGlobalRows) { rows_here = g_rows; rows_lead = g_rows; }
    const size_t buf_doubles = (size_t)F * kRow;
    auto evaluate = [&](const double* at, int buf) {
        if constexpr (kCluster) pose_evaluate_cluster(cluster, pc, img, F, C, huber_a, at, s_corner, rows_lead + buf * buf_doubles, mode);
        else pose_evaluate(pc, img, F, C, huber_a, at, rows_here + buf * buf_doubles);
    };
    // aide: the gradient (unscaled, as Ceres takes it before the Jacobi scaling) and max |x - Plus(x, -g)| at the point just evaluated
    auto aide_gmax = [&](const double* at, int buf) {
        if (aide) {
            const double* rw = rows_here + buf * buf_doubles;
            const int c = min(lane, 5);
            double s = 0;
#pragma unroll 8
            for (int f = 0; f < F; ++f) s += rw[f * kRow + 2 + c] * rw[f * kRow + 1];
            double ng[6], xp[7];
            for (int a = 0; a < 6; ++a) ng[a] = -__shfl_sync(0xffffffffu, s, a);
            pose_plus(at, ng, xp);
            double gm = 0;
            for (int k = 0; k < 7; ++k) gm = fmax(gm, fabs(at[k] - xp[k]));
            if (lane == 0) s_gmax = gm;
        }
    };
    auto meet = [&]() { if (boss || aide) asm volatile("bar.sync 2, 64;" ::: "memory"); };
    // the boss's word and candidate reach the other threads (and, in a cluster, the other ranks: pushed, so that nobody reads rank
    // 0's words after the barrier, when the boss may already be writing the next ones)
    auto publish = [&](int word, const double* cand) -> int {
        const int slot = pub & 1;
        ++pub;
        if (boss) {
            if (lane == 0) {
                s_ctrl[slot] = word;
                for (int k = 0; k < 7; ++k) s_cand[k] = cand[k];
            }
            if constexpr (kCluster) {
                __syncwarp();
                const int n_words = ((int)cluster.num_blocks() - 1) * 8;
                for (int p = lane; p < n_words; p += 32) {
                    const int dst = 1 + p / 8, k = p % 8;
                    if (k < 7) cluster.map_shared_rank(s_cand, dst)[k] = s_cand[k];
                    else cluster.map_shared_rank(s_ctrl, dst)[slot] = s_ctrl[slot];
                }
            }
        }
        if constexpr (kCluster) cluster.sync();
        else __syncthreads();
        return s_ctrl[slot];
    };

    // ---- boss state, the same in every lane of the warp ----
    double x[7], cand[7], best[7], A[21], g[6], scale[6], diag[6];   // A: J'J of the Jacobi-scaled Jacobian, packed upper triangle (sym6)
    double x_cost = 0, x_norm = 0, radius = 1e4, decrease_factor = 2.0, gmax = 0, model_cost_change = 0;
    double minimum_cost = DBL_MAX, initial_cost = 0, min_iter_cost = 0;
    int iteration = 0, ninvalid = 0, n_iters = 0, term = 1, cur = 0;   // cur: the row buffer of x; candidates fill the other one
    bool reuse_diagonal = false, step_ok = true, failed = false;
    // this lane's sum: q < 21 the J'J entry (a <= b) of that packed index, 21 .. 26 J'r, 27 the cost
    int ia = 0, ib = 0, la = 0, lb = 0;
    {
        int a = 0, rem = lane;
        while (a < 6 && rem >= 6 - a) { rem -= 6 - a; ++a; }
        if (lane < 21) { la = a; lb = a + rem; ia = 2 + la; ib = 2 + lb; }
        else if (lane < 27) { la = lane - 21; ia = 2 + la; ib = 1; }
    }
    if (threadIdx.x < 7) s_x[threadIdx.x] = T0[threadIdx.x];
    __syncthreads();
    for (int k = 0; k < 7; ++k) { x[k] = s_x[k]; cand[k] = x[k]; best[k] = x[k]; }

    // boss: cost, J'J and J'r of the rows, every sum by one lane in frame order (DenseNormalCholeskySolver's loop over the residual
    // blocks; the Jacobian is Jacobi-scaled entry by entry first, as Ceres scales its columns). first: the scaling itself, from the
    // unscaled column norms (TrustRegionMinimizer::IterationZero)
    double s_cost = 0, s_A[21], s_g[6];
    auto boss_sums = [&](int buf, bool first) {
        const double* rw = rows_here + buf * buf_doubles;
        if (first) {
            const int c = min(lane, 5);
            double s = 0;
#pragma unroll 8
            for (int f = 0; f < F; ++f) { const double j = rw[f * kRow + 2 + c]; s += j * j; }
            const double sc = 1.0 / (1.0 + sqrt(s));
            for (int a = 0; a < 6; ++a) scale[a] = __shfl_sync(0xffffffffu, sc, a);
        }
        double sa = scale[5], sb = scale[5];
        for (int a = 4; a >= 0; --a) { sa = la == a ? scale[a] : sa; sb = lb == a ? scale[a] : sb; }
        if (lane >= 21) sb = 1.0;
        if (lane >= 27) sa = 1.0;
        double t = 0;
#pragma unroll 8
        for (int f = 0; f < F; ++f) {
            const double* p = rw + f * kRow;
            const double u = p[ia] * sa, w = p[ib] * sb;
            t += lane >= 27 ? u : u * w;
        }
        for (int q = 0; q < 21; ++q) s_A[q] = __shfl_sync(0xffffffffu, t, q);
        for (int a = 0; a < 6; ++a) s_g[a] = __shfl_sync(0xffffffffu, t, 21 + a);
        s_cost = __shfl_sync(0xffffffffu, t, 27);
    };
    auto sums_finite = [&]() {
        bool fin = true;
        for (int q = 0; q < 21; ++q) fin = fin && isfinite(s_A[q]);
        for (int a = 0; a < 6; ++a) fin = fin && isfinite(s_g[a]);
        return fin;
    };
    auto take_sums = [&]() { x_cost = s_cost; for (int q = 0; q < 21; ++q) A[q] = s_A[q]; for (int a = 0; a < 6; ++a) g[a] = s_g[a]; };
    // boss: the model's cost change for `step`, a term per frame (one lane each), summed in frame order
    auto model_change = [&](const double* step) -> double {
        const double* rw = rows_here + cur * buf_doubles;
        double* term = rows_here + (cur ^ 1) * buf_doubles;       // the other buffer is free until the next evaluation fills it
        for (int f = lane; f < F; f += 32) {
            const double* p = rw + f * kRow;
            double s = 0;
            for (int c = 0; c < 6; ++c) s += (p[2 + c] * scale[c]) * step[c];
            term[f] = s * (p[1] + s / 2.0);
        }
        __syncwarp();
        double dot = 0;
#pragma unroll 8
        for (int f = 0; f < F; ++f) dot += term[f];
        __syncwarp();
        return -dot;
    };
    // boss: Ceres' "finalize the iteration, then compute the next trust-region step" (invalid steps loop without an evaluation);
    // 0 = cand is to be evaluated, 1 = stop
    auto next_step = [&]() -> int {
        for (;;) {
            if (step_ok && x_cost < minimum_cost) { minimum_cost = x_cost; for (int k = 0; k < 7; ++k) best[k] = x[k]; }
            n_iters++;
            if (iteration >= max_iters) { term = 1; return 1; }
            if (step_ok && gmax <= 1e-10) { term = 0; return 1; }
            if (radius <= 1e-32) { term = 0; return 1; }
            iteration++;
            step_ok = false;
#ifdef RCLC_POSE_TIMING
            long long stick_ = clock64();
#endif
            if (!reuse_diagonal) for (int a = 0; a < 6; ++a) diag[a] = fmin(fmax(A[sym6(a, a)], 1e-6), 1e32);
            double add[6], y[6], step[6];
            {   // lane a takes sqrt(diag[a] / radius) (six quotients and roots side by side instead of in a row)
                double mine = diag[5];
                for (int a = 4; a >= 0; --a) mine = lane == a ? diag[a] : mine;
                const double lm = sqrt(mine / radius);
                for (int a = 0; a < 6; ++a) { const double l = __shfl_sync(0xffffffffu, lm, a); add[a] = l * l; }
            }
            STEP_TICK(0);
            bool valid = chol6_solve_warp(A, add, g, y, lane);
            STEP_TICK(1);
            reuse_diagonal = true;
            double mcc = 0;
            if (valid) {
                for (int a = 0; a < 6; ++a) step[a] = -y[a];
                mcc = model_change(step);
                valid = mcc > 0.0;
            }
            STEP_TICK(2);
            if (!valid) {
                if (++ninvalid >= 5) { term = 2; failed = true; return 1; }
                radius *= 0.5;
                min_iter_cost = fmin(min_iter_cost, x_cost);
                continue;
            }
            ninvalid = 0;
            model_cost_change = mcc;
            double delta[6];
            for (int a = 0; a < 6; ++a) delta[a] = step[a] * scale[a];
            pose_plus(x, delta, cand);
            STEP_TICK(3);
            return 0;
        }
    };

    evaluate(s_x, 0);
    aide_gmax(s_x, 0);
    int stop_word = 1;
    if (boss) {
        boss_sums(0, true);
        if (!isfinite(s_cost) || !sums_finite()) { failed = true; term = 2; initial_cost = nan(""); min_iter_cost = nan(""); }
        else {
            take_sums();
            initial_cost = x_cost; min_iter_cost = x_cost;
            x_norm = 0; for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k];
            x_norm = sqrt(x_norm);
        }
    }
    meet();
    if (boss && !failed) { gmax = s_gmax; stop_word = next_step(); }
#ifdef RCLC_POSE_TIMING
    long long t_step = 0, t_pub = 0, t_eval = 0, t_decide = 0, tick_ = clock64();
#endif
    int word = publish(stop_word | ((cur ^ 1) << 1), cand);
    while (!(word & 1)) {
        POSE_TICK(t_pub);
        // the candidate: cost and Jacobian in one pass (Ceres evaluates the Jacobian again on acceptance: the same numbers)
        const int buf = (word >> 1) & 1;
        evaluate(s_cand, buf);
        aide_gmax(s_cand, buf);
        POSE_TICK(t_eval);
        bool go_on = false, accepted = false;
        if (boss) {
            boss_sums(buf, false);
#ifdef RCLC_POSE_TIMING
            { const long long now_ = clock64(); if (threadIdx.x == 0) g_step_ticks[4] += now_ - tick_; }
#endif
            const double cand_cost = isfinite(s_cost) ? s_cost : DBL_MAX;
            double sn = 0;
            for (int k = 0; k < 7; ++k) sn += (x[k] - cand[k]) * (x[k] - cand[k]);
            if (sqrt(sn) <= 1e-8 * (x_norm + 1e-8)) term = 0;
            else if (fabs(x_cost - cand_cost) <= ftol * x_cost) term = 0;
            else {
                const double rel = (cand_cost >= DBL_MAX) ? -DBL_MAX : (x_cost - cand_cost) / model_cost_change;
                if (rel > 1e-3) {
                    for (int k = 0; k < 7; ++k) x[k] = cand[k];
                    x_norm = 0; for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k];
                    x_norm = sqrt(x_norm);
                    if (!sums_finite()) { term = 2; failed = true; }
                    else {
                        take_sums();
                        cur = buf;
                        step_ok = true;
                        radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - cube_rn(2.0 * rel - 1.0)));
                        decrease_factor = 2.0;
                        reuse_diagonal = false;
                        min_iter_cost = fmin(min_iter_cost, x_cost);
                        go_on = true; accepted = true;
                    }
                } else {
                    min_iter_cost = fmin(min_iter_cost, cand_cost);
                    radius = radius / decrease_factor;
                    decrease_factor *= 2.0;
                    reuse_diagonal = true;
                    step_ok = false;
                    go_on = true;
                }
            }
        }
        POSE_TICK(t_decide);
        meet();
        stop_word = 1;
        if (boss && go_on) {
            if (accepted) gmax = s_gmax;
            stop_word = next_step();
        }
        POSE_TICK(t_step);
        word = publish(stop_word | ((cur ^ 1) << 1), cand);
    }
    if (boss && lane == 0) {
        for (int k = 0; k < 7; ++k) out->Tcl[k] = failed ? T0[k] : best[k];
        out->initial_cost = initial_cost;
        out->final_cost = fmin(initial_cost, min_iter_cost);
        out->iterations = n_iters;
        out->termination = term;
#ifdef RCLC_POSE_TIMING
        out->t_step = t_step; out->t_pub = t_pub; out->t_eval = t_eval; out->t_decide = t_decide;
#endif
    }
    if constexpr (kCluster) cluster.sync();           // nobody leaves while rank 0's shared memory may still be written
}

template <bool kGlobalRows>
__global__ void __launch_bounds__(kPoseThreads) k_pose_refine(const double* __restrict__ pc, const double* __restrict__ img, int F, int C,
                                                              double huber_a, double ftol, int max_iters, const double* __restrict__ T0,
                                                              PoseOut* __restrict__ out, double* g_rows)
{
    pose_refine_body<false, kGlobalRows>(pc, img, F, C, huber_a, ftol, max_iters, T0, out, g_rows, 0);
}

template <bool kGlobalRows>
__global__ void __launch_bounds__(kPoseThreads) k_pose_refine_cluster(const double* __restrict__ pc, const double* __restrict__ img, int F, int C,
                                                                      double huber_a, double ftol, int max_iters,
                                                                      const double* __restrict__ T0, PoseOut* __restrict__ out, double* g_rows,
                                                                      int mode)
{
    pose_refine_body<true, kGlobalRows>(pc, img, F, C, huber_a, ftol, max_iters, T0, out, g_rows, mode);
}

__global__ void k_pose_eval(const double* __restrict__ pc, const double* __restrict__ img, int F, int C, const double* __restrict__ Tcl,
                            double* __restrict__ res, double* __restrict__ jac)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double r, j[6];
    pose_block_dev(Tcl, pc + (size_t)f * C * 3, img + (size_t)f * C * 2, C, r, j);
    res[f] = r;
    for (int c = 0; c < 6; ++c) jac[6 * f + c] = j[c];
}

// debug: the split evaluation (rotation, corners, frame assembly) one thread per frame, scratch in global memory
__global__ void k_pose_eval_split(const double* __restrict__ pc, const double* __restrict__ img, int F, int C, const double* __restrict__ Tcl,
                                  double* __restrict__ res, double* __restrict__ jac, double* __restrict__ scratch)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    J7 R[9], t[3];
    pose_rot_dev(Tcl, R, t);
    double* sc = scratch + (size_t)f * kCornerVals * C;
    for (int i = 0; i < C; ++i) pose_corner_dev(R, t, pc + ((size_t)f * C + i) * 3, img + ((size_t)f * C + i) * 2, sc + i, C);
    double r, j[6];
    pose_frame_from_corners(Tcl, sc, C, C, r, j);
    res[f] = r;
    for (int c = 0; c < 6; ++c) jac[6 * f + c] = j[c];
}

static int upload_pose_inputs(rclc_ctx* ctx, const double* pc, const double* img, int F, int C, const double* Tcl,
                              double** d_pc, double** d_img, double** d_T)
{
    const size_t n_pc = (size_t)F * C * 3, n_img = (size_t)F * C * 2;
    RCLC_TRY(ctx->d_tmp.reserve((n_pc + n_img + 8) * 8));
    *d_pc = ctx->d_tmp.as<double>();
    *d_img = *d_pc + n_pc;
    *d_T = *d_img + n_img;
    RCLC_CUDA(cudaMemcpyAsync(*d_pc, pc, n_pc * 8, cudaMemcpyHostToDevice, ctx->stream));
    RCLC_CUDA(cudaMemcpyAsync(*d_img, img, n_img * 8, cudaMemcpyHostToDevice, ctx->stream));
    RCLC_CUDA(cudaMemcpyAsync(*d_T, Tcl, 7 * 8, cudaMemcpyHostToDevice, ctx->stream));
    return RCLC_OK;
}

// The frames' rows (two buffers of F x kRow doubles) live in the leading CTA's shared memory while they fit beside the rest,
// in global memory (ctx->d_tmp2) beyond.
constexpr size_t kPoseSmemBudget = 200 * 1024;
static int pose_rows(rclc_ctx* ctx, int F, size_t other_bytes, size_t* dyn, double** g_rows)
{
    const size_t rows_bytes = 2 * (size_t)F * kRow * sizeof(double);
    *g_rows = nullptr;
    *dyn = other_bytes + rows_bytes;
    if (*dyn > kPoseSmemBudget) {
        RCLC_TRY(ctx->d_tmp2.reserve(rows_bytes));
        *g_rows = ctx->d_tmp2.as<double>();
        *dyn = other_bytes;
    }
    return RCLC_OK;
}

// The cluster launch: 16 CTAs (non-portable size, opt-in), else 8. RCLC_ERR_UNSUPPORTED when the per-corner factors of a CTA's
// frames do not fit shared memory or no cluster can be launched; the caller then uses the one-CTA kernel.
static int launch_pose_cluster(rclc_ctx* ctx, const double* pc, const double* img, int F, int C, double huber_a, double ftol, int max_iters,
                               const double* T0, PoseOut* out)
{
    const int mode = ctx->pose_debug;                 // test hook (rclc_ctx_set_pose_test_hooks): 0 in production
    for (int n_cta : {16, 8}) {
        const int S = kPoseThreads / n_cta;
        const size_t corner_bytes = (size_t)kCornerVals * corner_stride(S, C) * sizeof(double);
        if (corner_bytes > kPoseSmemBudget) continue;
        size_t dyn; double* g_rows;
        RCLC_TRY(pose_rows(ctx, F, corner_bytes, &dyn, &g_rows));
        auto kern = g_rows ? k_pose_refine_cluster<true> : k_pose_refine_cluster<false>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) { cudaGetLastError(); continue; }
        if (n_cta > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)n_cta);
        lc.blockDim = dim3(kPoseThreads);
        lc.dynamicSmemBytes = dyn;
        lc.stream = ctx->stream;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = (unsigned)n_cta; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        lc.attrs = &at;
        lc.numAttrs = 1;
        if (cudaLaunchKernelEx(&lc, kern, pc, img, F, C, huber_a, ftol, max_iters, T0, out, g_rows, mode) == cudaSuccess) return RCLC_OK;
        cudaGetLastError();
    }
    return RCLC_ERR_UNSUPPORTED;
}

} // namespace rclc

using namespace rclc;

extern "C" {

int rclc_pose_eval(rclc_ctx* ctx, const double* pc_corners, const double* img_norm, int32_t F, int32_t C,
                   const double Tcl[7], double* residuals, double* jac6)
{
    RCLC_CHECK(ctx && pc_corners && img_norm && Tcl && residuals && F > 0 && C > 0, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    double *d_pc, *d_img, *d_T;
    RCLC_TRY(upload_pose_inputs(ctx, pc_corners, img_norm, F, C, Tcl, &d_pc, &d_img, &d_T));
    RCLC_TRY(ctx->d_out.reserve((size_t)F * 7 * 8));
    double* d_res = ctx->d_out.as<double>();
    double* d_jac = d_res + F;
    if (ctx->pose_eval_split) {                       // test hook: the per-corner / per-frame split of the cluster shape, on the plain grid
        RCLC_TRY(ctx->d_tmp2.reserve((size_t)F * kCornerVals * C * 8));
        k_pose_eval_split<<<(F + 127) / 128, 128, 0, ctx->stream>>>(d_pc, d_img, F, C, d_T, d_res, d_jac, ctx->d_tmp2.as<double>());
    } else
    k_pose_eval<<<(F + 127) / 128, 128, 0, ctx->stream>>>(d_pc, d_img, F, C, d_T, d_res, d_jac);
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    RCLC_CUDA(cudaMemcpyAsync(residuals, d_res, (size_t)F * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (jac6) RCLC_CUDA(cudaMemcpyAsync(jac6, d_jac, (size_t)F * 6 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
    return RCLC_OK;
}

int rclc_pose_refine(rclc_ctx* ctx, const rclc_config* cfg, const double* pc_corners, const double* img_norm,
                     int32_t F, int32_t C, double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations)
{
    RCLC_CHECK(ctx && cfg && pc_corners && img_norm && Tcl_io && F > 0 && C > 0, RCLC_ERR_INVALID, "bad argument");
    RCLC_CUDA(cudaSetDevice(ctx->device));
    double *d_pc, *d_img, *d_T;
    RCLC_TRY(upload_pose_inputs(ctx, pc_corners, img_norm, F, C, Tcl_io, &d_pc, &d_img, &d_T));
    return rclc::pose_refine_dev(ctx, cfg, d_pc, d_img, F, C, d_T, Tcl_io, initial_cost, final_cost, iterations);
}

} // extern "C"

// corners, image points and the start (7 doubles) already on the device
int rclc::pose_refine_dev(rclc_ctx* ctx, const rclc_config* cfg, const double* d_pc, const double* d_img, int32_t F, int32_t C, const double* d_T,
                          double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations)
{
    RCLC_TRY(ctx->d_out.reserve(sizeof(PoseOut)));
    // pose_shape (rclc_ctx_set_pose_test_hooks): 1 keeps the one-CTA kernel, 2 takes the cluster at any size — the A/B switch of the
    // bit-identity test; both shapes return the same bits (the oracle's)
    bool launched = false;
    // the cluster from 64 corner observations up: it is three times faster than one CTA already at 5 frames x 35 corners (9.6 vs 29 ms)
    if (ctx->pose_shape != 1 && ((int64_t)F * C >= 64 || ctx->pose_shape == 2))
        launched = launch_pose_cluster(ctx, d_pc, d_img, F, C, cfg->pose_huber_a, cfg->pose_function_tol, cfg->pose_max_lm_iters, d_T,
                                       ctx->d_out.as<PoseOut>()) == RCLC_OK;
    if (!launched) {
        size_t dyn; double* g_rows;
        RCLC_TRY(pose_rows(ctx, F, 0, &dyn, &g_rows));
        auto kern = g_rows ? k_pose_refine<true> : k_pose_refine<false>;
        RCLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<1, kPoseThreads, dyn, ctx->stream>>>(d_pc, d_img, F, C, cfg->pose_huber_a, cfg->pose_function_tol, cfg->pose_max_lm_iters,
                                                            d_T, ctx->d_out.as<PoseOut>(), g_rows);
    }
    ctx->launches += 1;
    RCLC_CUDA(cudaGetLastError());
    PoseOut h;
    RCLC_CUDA(cudaMemcpyAsync(&h, ctx->d_out.p, sizeof(PoseOut), cudaMemcpyDeviceToHost, ctx->stream));
    RCLC_CUDA(cudaStreamSynchronize(ctx->stream));
#ifdef RCLC_POSE_TIMING
    if (launched) {
        long long tk[8], z[8] = {};
        cudaMemcpyFromSymbol(tk, g_pose_ticks, sizeof(tk));
        cudaMemcpyToSymbol(g_pose_ticks, z, sizeof(z));
        fprintf(stderr, "  evaluate, cycles per iteration: corners %.0f, wait %.0f, frames %.0f (corner loop %.0f, epilogue %.0f, row %.0f), wait %.0f, cluster sync %.0f\n",
                (double)tk[1] / h.iterations, (double)tk[2] / h.iterations, (double)tk[3] / h.iterations, (double)tk[6] / h.iterations,
                (double)tk[7] / h.iterations, (double)tk[0] / h.iterations, (double)tk[4] / h.iterations, (double)tk[5] / h.iterations);
    }
    {
        long long tk[8], z[8] = {};
        cudaMemcpyFromSymbol(tk, g_step_ticks, sizeof(tk));
        cudaMemcpyToSymbol(g_step_ticks, z, sizeof(z));
        fprintf(stderr, "  step, cycles per iteration: diagonal %.0f, cholesky %.0f, model change %.0f, plus %.0f; decide: sums %.0f\n",
                (double)tk[0] / h.iterations, (double)tk[1] / h.iterations, (double)tk[2] / h.iterations, (double)tk[3] / h.iterations, (double)tk[4] / h.iterations);
    }
    fprintf(stderr, "pose timing (%s, F=%d): %d iterations; cycles per iteration: step %.0f, publish %.0f, evaluate %.0f, decide %.0f\n",
            launched ? "cluster" : "one CTA", F, h.iterations, (double)h.t_step / h.iterations, (double)h.t_pub / h.iterations,
            (double)h.t_eval / h.iterations, (double)h.t_decide / h.iterations);
#endif
    for (int k = 0; k < 7; ++k) Tcl_io[k] = h.Tcl[k];
    if (initial_cost) *initial_cost = h.initial_cost;
    if (final_cost) *final_cost = h.final_cost;
    if (iterations) *iterations = h.iterations;
    return h.termination == 2 ? RCLC_ERR_NUMERIC : RCLC_OK;
}

extern "C" int rclc_ctx_set_pose_test_hooks(rclc_ctx* ctx, int32_t shape, int32_t cluster_debug, int32_t eval_split)
{
    RCLC_CHECK(ctx && shape >= 0 && shape <= 2 && cluster_debug >= 0 && cluster_debug <= 1, RCLC_ERR_INVALID, "bad pose test hook");
    ctx->pose_shape = shape; ctx->pose_debug = cluster_debug; ctx->pose_eval_split = eval_split ? 1 : 0;
    return RCLC_OK;
}
