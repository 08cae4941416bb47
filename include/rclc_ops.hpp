// rclc_ops.hpp — the reference's C++ operator surface on top of the C ABI (rclc_b200.h).
//
// zhijianglu/RCLC has no plugin layer: its hot path is a set of header-inline free functions and small classes with
// PCL / Eigen / Ceres types in their signatures. This header gives that path the SAME names, argument
// order, in/out conventions and "never throws, prints and carries on" error behaviour, and forwards to
// librclc_b200.so:
//
//   reference symbol (file:line)                                         forwards to
//   ---------------------------------------------------------------------------------------------------------
//   GMMFit::fit_1d                       include/GMMFit.h:31                rclc_gmm_fit_1d
//   pc_plane_prj                         include/pc_process.h:41            rclc_plane_project
//   pc_plane_fitting                     include/opt/opt_plane_param.hpp:45 rclc_plane_fit
//   opt_grid_fitting                     include/opt/opt_grid_fitting.h:25  rclc_gridfit_solve
//   BOARD_FITTING_COST::Create/Evaluate  include/opt/board_fitting_cost.h:64-78   rclc_batch_* (one sweep serves all 82 functors)
//   cv::solvePnPRansac + Rodrigues + Quaterniond(R)   apps/Calibrate.cpp:41-62       rclc_pnp_ransac  (pnp_init)
//   pose_reprj_opt                       include/opt/opt_reprj_calib.h:21   rclc_pose_refine
//   PinholeCam::{imgPt2norm,normPt2img,Pt2img,is_in_FoV}  include/PinholeCam.h:78-110   (host inline, same arithmetic)
//   colourise loop                       apps/Calibrate.cpp:134-148         rclc_project_colorize  (colorize_cloud)
//   pcl::UniformSampling<PoinT>          include/pc_process.h:111-115, apps/BoardSegmentation.cpp:58-61   rclc_uniform_sample  (UniformSampling)
//   pcl::EuclideanClusterExtraction<PoinT>  include/pc_process.h:62-72, 131-139     rclc_euclidean_cluster  (EuclideanClusterExtraction)
//   EulerCluster, eular_cluster_iter     include/pc_process.h:59-79, 104-211  (host logic over the two classes above)
//   get_ref_pt + calc_laser_coor_v1 + line_fitting   include/utils.h:160, pc_process.h:373, opt/opt_line_fitting.h:51   rclc_calc_laser_coor
//   pcl::transformPointCloud(in, out, Matrix4f)      pc_process.h:594,618; Calibrate.cpp:99   rclc_transform_cloud  (transformPointCloud)
//   est_board_corner                     include/pc_process.h:583-646       (the per-frame chain over the operators above)
//   pcl::PassThrough, cut_roi, the per-cloud body of main()   pc_process.h:81-101, apps/BoardSegmentation.cpp:45-165   (PassThrough, cut_roi, segment_board)
//   pcl::SACSegmentation<>::segment as configured at apps/BoardSegmentation.cpp:64-72
//                                                                           rclc_ransac_plane_score + host replay of
//                                                                           PCL's adaptive rule + rclc_ransac_plane_select
//
// Types. With -DRCLC_WITH_PCL_EIGEN the signatures use the real pcl:: / Eigen:: types, so the header drops into
// BoardSegmentation.cpp / Calibrate.cpp unchanged. Neither library exists in the build image of this repository, so by
// default it compiles against the small stand-ins in rclc_b200::compat, which have the SAME memory layout
// (pcl::PointXYZI is 32 bytes: x y z pad | intensity pad pad pad) and the accessors the shims use.
//
// Threading: the reference calls these operators inside `#pragma omp parallel for` over frames; every host thread
// gets its own rclc_ctx (one CUDA stream + scratch) through thread_ctx().
#ifndef RCLC_OPS_HPP
#define RCLC_OPS_HPP

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "rclc_b200.h"

#ifdef RCLC_WITH_PCL_EIGEN
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/ModelCoefficients.h>
#include <pcl/PointIndices.h>
#endif

namespace rclc_b200 {

#ifndef RCLC_WITH_PCL_EIGEN
// ---------------------------------------------------------------------------------------------------------------
// Stand-ins with the layout and the accessors of the PCL / Eigen types in the reference's signatures.
// ---------------------------------------------------------------------------------------------------------------
namespace compat {
struct alignas(16) PointXYZI {          // == pcl::PointXYZI (32 bytes)
    float x = 0, y = 0, z = 0, _w = 1.f;
    float intensity = 0, _p[3] = {0, 0, 0};
};
struct alignas(16) PointXYZRGB {        // == pcl::PointXYZRGB (32 bytes, bgra packed at offset 16)
    float x = 0, y = 0, z = 0, _w = 1.f;
    uint8_t b = 0, g = 0, r = 0, a = 255;
    float _p[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZI) == 32 && sizeof(PointXYZRGB) == 32, "PCL point layout");
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
    size_t size() const { return points.size(); }
};
template <int N> struct Vec {
    double v[N];
    Vec() { for (int i = 0; i < N; ++i) v[i] = 0; }
    Vec(double a, double b) { static_assert(N == 2, ""); v[0] = a; v[1] = b; }
    Vec(double a, double b, double c) { static_assert(N == 3, ""); v[0] = a; v[1] = b; v[2] = c; }
    Vec(double a, double b, double c, double d) { static_assert(N == 4, ""); v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double& x() { return v[0]; } double& y() { return v[1]; } double& z() { return v[2]; } double& w() { return v[3]; }
    double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; } double w() const { return v[3]; }
};
typedef Vec<2> Vector2d;
typedef Vec<3> Vector3d;
typedef Vec<4> Vector4d;
template <class T> struct Mat4 {        // (row, col) access like Eigen; storage order is private to the shim
    T m[16];
    Mat4() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? T(1) : T(0); }
    T& operator()(int r, int c) { return m[r * 4 + c]; }
    T operator()(int r, int c) const { return m[r * 4 + c]; }
};
typedef Mat4<double> Matrix4d;
typedef Mat4<float> Matrix4f;
struct VectorXd {                       // the subset GMMFit::fit_1d uses
    std::vector<double> d;
    VectorXd() {}
    explicit VectorXd(long n) : d((size_t)n, 0.0) {}
    long size() const { return (long)d.size(); }
    long rows() const { return (long)d.size(); }
    double& operator()(long i) { return d[(size_t)i]; }
    double operator()(long i) const { return d[(size_t)i]; }
};
struct PointIndices { std::vector<int> indices; };
struct ModelCoefficients { std::vector<float> values; };
}  // namespace compat
using compat::PointXYZI; using compat::PointXYZRGB; using compat::PointCloud;
using compat::Vector2d; using compat::Vector3d; using compat::Vector4d; using compat::Matrix4d; using compat::Matrix4f;
using compat::VectorXd; using compat::PointIndices; using compat::ModelCoefficients;
#else
typedef pcl::PointXYZI PointXYZI; typedef pcl::PointXYZRGB PointXYZRGB;
template <class P> using PointCloud = pcl::PointCloud<P>;
using Eigen::Vector2d; using Eigen::Vector3d; using Eigen::Vector4d; using Eigen::Matrix4d; using Eigen::Matrix4f; using Eigen::VectorXd;
typedef pcl::PointIndices PointIndices; typedef pcl::ModelCoefficients ModelCoefficients;
#endif
typedef PointXYZI PoinT;                // include/global.h:59 "typedef pcl::PointXYZI PoinT"
typedef PointXYZRGB DisplayType;        // include/global.h:60

// ---------------------------------------------------------------------------------------------------------------
// Session state: the reference's globals CfgParam / Cam (include/global.h:118-120) become one POD config, read-only
// after set_config(); every host thread owns a context.
// ---------------------------------------------------------------------------------------------------------------
inline rclc_config& CfgParam()
{
    static rclc_config cfg = [] { rclc_config c; rclc_config_default(&c); return c; }();
    return cfg;
}
inline int& device_ordinal() { static int dev = 0; return dev; }

struct CtxHolder {
    rclc_ctx* ctx = nullptr;
    ~CtxHolder() { if (ctx) rclc_ctx_destroy(ctx); }
};
// nullptr (after printing why) when there is no CUDA device: there is NO CPU fallback behind these operators.
inline rclc_ctx* thread_ctx()
{
    static thread_local CtxHolder h;
    if (!h.ctx && rclc_ctx_create(device_ordinal(), &h.ctx) != RCLC_OK) {
        std::fprintf(stderr, "rclc_b200: no context (%s)\n", rclc_last_error_string());
        h.ctx = nullptr;
    }
    return h.ctx;
}
inline bool ok(int rc, const char* what)
{
    // the reference's operators do not report errors (failures are printed, Calibrate.cpp carries on): same here
    if (rc != RCLC_OK) std::fprintf(stderr, "rclc_b200: %s failed (%d): %s\n", what, rc, rclc_last_error_string());
    return rc == RCLC_OK;
}
inline void to_row_major(const Matrix4d& T, double out[16]) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[r * 4 + c] = T(r, c); }
inline void from_row_major(const double in[16], Matrix4d& T) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T(r, c) = in[r * 4 + c]; }

// ---------------------------------------------------------------------------------------------------------------
// GMMFit::fit_1d — include/GMMFit.h:31, src/GMMFit.cpp:7-108. mu / sigma in-out. (The reference's function falls off
// its end without a return value; callers ignore it. This one returns whether the device call succeeded.)
// The reference feeds it float intensities widened to double (pc_process.h:600-606); the kernel reads floats.
// ---------------------------------------------------------------------------------------------------------------
class GMMFit {
public:
    static bool fit_1d(const VectorXd& in_vec, const int K, std::vector<double>& mu, std::vector<double>& sigma, bool show_debug_info = false)
    {
        (void)show_debug_info;
        rclc_ctx* ctx = thread_ctx();
        if (!ctx || (int)mu.size() != K || (int)sigma.size() != K) return false;
        std::vector<float> x((size_t)in_vec.size());
        for (long i = 0; i < (long)in_vec.size(); ++i) x[(size_t)i] = (float)in_vec(i);
        double priors[8] = {0};
        return ok(rclc_gmm_fit_1d(ctx, x.data(), 4, (int64_t)x.size(), K, CfgParam().gmm_iters, mu.data(), sigma.data(), priors), "GMMFit::fit_1d");
    }
};

// pc_plane_prj — include/pc_process.h:41-57: slides every point along its ray onto the plane, in place.
inline void pc_plane_prj(typename PointCloud<PoinT>::Ptr& Points, Vector4d& plane_param_l)
{
    rclc_ctx* ctx = thread_ctx();
    if (!ctx || !Points || Points->size() == 0) return;
    const double plane[4] = {plane_param_l.x(), plane_param_l.y(), plane_param_l.z(), plane_param_l.w()};
    ok(rclc_plane_project(ctx, Points->points.data(), (int)sizeof(PoinT), (int64_t)Points->size(), plane), "pc_plane_prj");
}

// pc_plane_fitting — include/opt/opt_plane_param.hpp:45-88: Huber(0.1)-robust plane through the points whose intensity is at
// least reflactive_thd (default 100, as there); param = (a, b, c, d), not normalised. Always returns 0 like the reference.
inline int pc_plane_fitting(typename PointCloud<PoinT>::Ptr& Points, Vector4d& param, double reflactive_thd = 100)
{
    rclc_ctx* ctx = thread_ctx();
    if (!ctx || !Points) return 0;
    rclc_config cfg = CfgParam();
    cfg.plane_reflactive_thd = reflactive_thd;
    cfg.plane_huber_a = 0.1;                                   // hard-coded at opt_plane_param.hpp:63
    double abcd[4] = {0.1, 0.1, 1, 1};
    if (Points->size() > 0)
        ok(rclc_plane_fit(ctx, &cfg, Points->points.data(), (int)sizeof(PoinT), 16, (int64_t)Points->size(), abcd, nullptr, nullptr),
           "pc_plane_fitting");
    param = Vector4d(abcd[0], abcd[1], abcd[2], abcd[3]);
    return 0;
}

// opt_grid_fitting — include/opt/opt_grid_fitting.h:25-132. cloud in the board frame, T_bl in/out, 35 corners appended.
inline void opt_grid_fitting(typename PointCloud<PoinT>::Ptr& cloud_src, Matrix4d& T_bl, std::vector<Vector3d>& est_pc_corners,
                             void* viewer = nullptr, rclc_gridfit_report* report = nullptr)
{
    (void)viewer;
    rclc_ctx* ctx = thread_ctx();
    if (!ctx || !cloud_src || cloud_src->size() == 0) return;
    const rclc_config& cfg = CfgParam();
    const int nc = (cfg.board_width - 1) * (cfg.board_height - 1);
    std::vector<double> corners((size_t)nc * 3);
    double Tb[16];
    to_row_major(T_bl, Tb);
    rclc_gridfit_report rep;
    if (!ok(rclc_gridfit_solve(ctx, &cfg, cloud_src->points.data(), (int)sizeof(PoinT), 16, (int64_t)cloud_src->size(), Tb, corners.data(), &rep),
            "opt_grid_fitting"))
        return;
    from_row_major(Tb, T_bl);
    for (int k = 0; k < nc; ++k) est_pc_corners.push_back(Vector3d(corners[3 * k], corners[3 * k + 1], corners[3 * k + 2]));
    if (report) *report = rep;
    std::printf("cost: %g --> %g\n", rep.first_initial_cost, rep.last_final_cost);      // opt_grid_fitting.h:112
}

// ---------------------------------------------------------------------------------------------------------------
// BOARD_FITTING_COST — include/opt/board_fitting_cost.h:27-280. The reference builds one functor per target (82) and
// lets Ceres call Evaluate on each; here the functors of one cloud share a device-resident, binned copy of it, and
// the first Evaluate at a new (x, y, yaw) runs ONE sweep whose 82 residuals and Jacobian rows serve all of them.
// ---------------------------------------------------------------------------------------------------------------
class BoardCloudOnDevice {
public:
    explicit BoardCloudOnDevice(typename PointCloud<PoinT>::Ptr cloud)
    {
        rclc_ctx* ctx = thread_ctx();
        const rclc_config& cfg = CfgParam();
        nt_ = (cfg.board_width - 1) * cfg.board_height + (cfg.board_height - 1) * cfg.board_width;
        tgt_.resize((size_t)nt_ * 2); is_row_.resize((size_t)nt_);
        int32_t n = 0;
        rclc_generate_targets(&cfg, tgt_.data(), is_row_.data(), &n);
        if (!ctx || !cloud || cloud->size() == 0) return;
        if (!ok(rclc_batch_create(ctx, &cfg, 1, (int64_t)cloud->size(), &batch_), "BOARD_FITTING_COST (batch)")) { batch_ = nullptr; return; }
        valid_ = ok(rclc_batch_upload(batch_, 0, cloud->points.data(), (int)sizeof(PoinT), 16, (int64_t)cloud->size()), "upload") &&
                 ok(rclc_batch_bin(batch_), "bin");
        res_.assign((size_t)nt_, 0.0); jac_.assign((size_t)nt_ * 3, 0.0);
    }
    ~BoardCloudOnDevice() { if (batch_) rclc_batch_destroy(batch_); }
    int find_target(double x, double y, bool is_row) const
    {
        for (int t = 0; t < nt_; ++t)
            if ((is_row_[(size_t)t] != 0) == is_row && std::fabs(tgt_[2 * (size_t)t] - x) < 1e-9 && std::fabs(tgt_[2 * (size_t)t + 1] - y) < 1e-9) return t;
        return -1;
    }
    bool evaluate(const double pose[3])
    {
        if (!valid_) return false;
        if (have_ && pose[0] == pose_[0] && pose[1] == pose_[1] && pose[2] == pose_[2]) return true;
        if (!ok(rclc_batch_sweep(batch_, pose), "sweep") || !ok(rclc_batch_read_eval(batch_, res_.data(), jac_.data(), nullptr), "read")) return false;
        pose_[0] = pose[0]; pose_[1] = pose[1]; pose_[2] = pose[2];
        have_ = true;
        return true;
    }
    double residual(int t) const { return res_[(size_t)t]; }
    const double* jacobian(int t) const { return &jac_[(size_t)t * 3]; }
private:
    rclc_batch* batch_ = nullptr;
    bool valid_ = false, have_ = false;
    int nt_ = 0;
    double pose_[3] = {0, 0, 0};
    std::vector<double> tgt_, res_, jac_;
    std::vector<int32_t> is_row_;
};

class BOARD_FITTING_COST {          // (with Ceres present, derive from ceres::SizedCostFunction<1, 3>)
public:
    virtual ~BOARD_FITTING_COST() {}
    BOARD_FITTING_COST(Vector3d _target_point, bool _is_row, std::shared_ptr<BoardCloudOnDevice> _cloud)
        : cloud_(_cloud), target_(_cloud ? _cloud->find_target(_target_point.x(), _target_point.y(), _is_row) : -1) {}
    // _kdtree of the reference (board_fitting_cost.h:67) has no counterpart: the neighbour test runs on the device.
    static BOARD_FITTING_COST* Create(Vector3d _target_point, bool _is_row, std::shared_ptr<BoardCloudOnDevice> _cloud_src)
    {
        return new BOARD_FITTING_COST(_target_point, _is_row, _cloud_src);
    }
    // parameters[0] = {x, y, yaw in degrees}; jacobians / jacobians[0] may be NULL (board_fitting_cost.h:256)
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const
    {
        if (!cloud_ || target_ < 0 || !cloud_->evaluate(parameters[0])) { residuals[0] = std::nan(""); return true; }
        residuals[0] = cloud_->residual(target_);
        if (jacobians && jacobians[0]) { const double* j = cloud_->jacobian(target_); jacobians[0][0] = j[0]; jacobians[0][1] = j[1]; jacobians[0][2] = j[2]; }
        return true;            // the reference returns true always (board_fitting_cost.h:261)
    }
private:
    std::shared_ptr<BoardCloudOnDevice> cloud_;
    int target_;
};

// pnp_init — apps/Calibrate.cpp:41-62: the starting extrinsic of pose_reprj_opt from all corner correspondences, on the
// normalised image points the reference builds at :36 (e_imagePoints). Tcl = (qx, qy, qz, qw, tx, ty, tz) as at :59-62.
// reprj_thresh_px / focal_px: OpenCV's default reprojectionError of 8 px at the camera's focal length. Host code.
inline bool pnp_init(const std::vector<std::vector<Vector2d>>& img_norm_points, const std::vector<std::vector<Vector3d>>& pc_corner_points, double* Tcl,
                     double reprj_thresh_px = 8.0, double focal_px = 0.0, unsigned seed = 1)
{
    std::vector<double> obj, uv;
    for (size_t f = 0; f < pc_corner_points.size() && f < img_norm_points.size(); ++f)
        for (size_t k = 0; k < pc_corner_points[f].size() && k < img_norm_points[f].size(); ++k) {
            obj.push_back(pc_corner_points[f][k].x()); obj.push_back(pc_corner_points[f][k].y()); obj.push_back(pc_corner_points[f][k].z());
            uv.push_back(img_norm_points[f][k].x()); uv.push_back(img_norm_points[f][k].y());
        }
    if (focal_px <= 0) focal_px = 0.5 * (CfgParam().fx + CfgParam().fy);
    double R[9], t[3];
    if (!ok(rclc_pnp_ransac(obj.data(), uv.data(), (int64_t)(uv.size() / 2), reprj_thresh_px / focal_px, 100, 0.99, seed, R, t, nullptr, nullptr), "pnp_init")) return false;
    // Eigen::Quaterniond(Matrix3d) (Calibrate.cpp:61): Shepperd's branches as in Eigen/src/Geometry/Quaternion.h
    double qw, qx, qy, qz;
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) { double s = std::sqrt(tr + 1.0); qw = 0.5 * s; s = 0.5 / s; qx = (R[7] - R[5]) * s; qy = (R[2] - R[6]) * s; qz = (R[3] - R[1]) * s; }
    else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        double q[3];
        q[i] = 0.5 * s; s = 0.5 / s;
        qw = (R[3 * k + j] - R[3 * j + k]) * s;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    Tcl[0] = qx; Tcl[1] = qy; Tcl[2] = qz; Tcl[3] = qw; Tcl[4] = t[0]; Tcl[5] = t[1]; Tcl[6] = t[2];
    return true;
}

// pose_reprj_opt — include/opt/opt_reprj_calib.h:21-66. Tcl = (qx, qy, qz, qw, tx, ty, tz) in/out (Calibrate.cpp:59-66).
inline void pose_reprj_opt(std::vector<std::vector<Vector2d>>& img_norm_points, std::vector<std::vector<Vector3d>>& pc_corner_points, double* Tcl,
                           bool debug_show = false)
{
    rclc_ctx* ctx = thread_ctx();
    const int F = (int)pc_corner_points.size();
    if (!ctx || F == 0 || img_norm_points.size() != pc_corner_points.size()) return;
    const int C = (int)pc_corner_points[0].size();
    std::vector<double> pc((size_t)F * C * 3), im((size_t)F * C * 2);
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < C; ++k) {
            for (int d = 0; d < 3; ++d) pc[((size_t)f * C + k) * 3 + d] = pc_corner_points[(size_t)f][(size_t)k][d];
            for (int d = 0; d < 2; ++d) im[((size_t)f * C + k) * 2 + d] = img_norm_points[(size_t)f][(size_t)k][d];
        }
    double c0 = 0, c1 = 0;
    int32_t it = 0;
    if (ok(rclc_pose_refine(ctx, &CfgParam(), pc.data(), im.data(), F, C, Tcl, &c0, &c1, &it), "pose_reprj_opt") && debug_show)
        std::printf("pose_reprj_opt: cost %g -> %g in %d iterations\n", c0, c1, it);
}

// ---------------------------------------------------------------------------------------------------------------
// PinholeCam — include/PinholeCam.h:78-110 (projection side only; undistort/remap is image-side OpenCV, out of scope)
// ---------------------------------------------------------------------------------------------------------------
class PinholeCam {
public:
    int cam_width = 0, cam_height = 0;
    double fx = 0, fy = 0, cx = 0, cy = 0;
    PinholeCam() {}
    PinholeCam(int w, int h, double _fx, double _fy, double _cx, double _cy) : cam_width(w), cam_height(h), fx(_fx), fy(_fy), cx(_cx), cy(_cy) {}
    Vector2d imgPt2norm(Vector2d& imgPt) { return Vector2d((imgPt.x() - cx) / fx, (imgPt.y() - cy) / fy); }
    Vector2d normPt2img(Vector3d normPt) { return Vector2d(normPt.x() * fx + cx, normPt.y() * fy + cy); }
    Vector2d Pt2img(Vector3d Pt)
    {   // cvRound == lrint under the default rounding mode: ties to even (PinholeCam.h:93-99)
        return Vector2d((double)std::lrint((Pt.x() / Pt.z()) * fx + cx), (double)std::lrint((Pt.y() / Pt.z()) * fy + cy));
    }
    bool is_in_FoV(Vector2d& pt) { return !(pt.x() < 0 || pt.y() < 0 || pt.x() >= cam_width || pt.y() >= cam_height); }
    void to_config(rclc_config& c) const { c.cam_width = cam_width; c.cam_height = cam_height; c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy; }
};

// The colourise loop of apps/Calibrate.cpp:134-148 as one call: transformPointCloud(T_cl) (float) -> Pt2img -> is_in_FoV ->
// BGR gather. `bgr` is the cv::Mat data pointer of a continuous CV_8UC3 image of cam_width x cam_height.
inline void colorize_cloud(typename PointCloud<DisplayType>::Ptr& pc_show, const Matrix4f& T_cl, const uint8_t* bgr,
                           std::vector<uint8_t>* in_fov = nullptr)
{
    rclc_ctx* ctx = thread_ctx();
    if (!ctx || !pc_show || pc_show->size() == 0) return;
    const int64_t n = (int64_t)pc_show->size();
    float T[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[r * 4 + c] = T_cl(r, c);
    std::vector<uint8_t> rgb((size_t)n * 3), fov((size_t)n);
    std::vector<int32_t> pix((size_t)n * 2);
    if (!ok(rclc_project_colorize(ctx, &CfgParam(), pc_show->points.data(), (int)sizeof(DisplayType), n, T, bgr, rgb.data(), pix.data(), fov.data()),
            "colorize_cloud"))
        return;
    for (int64_t i = 0; i < n; ++i) {
        DisplayType& pt = pc_show->points[(size_t)i];
        // the reference transforms the cloud in place first (Calibrate.cpp:133): x' = m00 x + m01 y + m02 z + m03 in float
        const float x = pt.x, y = pt.y, z = pt.z;
        pt.x = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
        pt.y = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
        pt.z = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
        if (!fov[(size_t)i]) continue;
        pt.r = rgb[3 * (size_t)i]; pt.g = rgb[3 * (size_t)i + 1]; pt.b = rgb[3 * (size_t)i + 2];
    }
    if (in_fov) in_fov->swap(fov);
}

// ---------------------------------------------------------------------------------------------------------------
// pcl::SACSegmentation<PoinT> as configured at apps/BoardSegmentation.cpp:64-72 (SAC_RANSAC, SACMODEL_PLANE,
// maxIterations 800, threshold 0.08, defaults: probability 0.99, optimize_coefficients true). PCL's loop is adaptive and
// sequential; here all hypotheses are scored in one device pass and the loop is replayed on
// the host over the counts, which selects the same model after the same number of iterations.
// ---------------------------------------------------------------------------------------------------------------
class PlaneSegmentation {
public:
    void setInputCloud(typename PointCloud<PoinT>::Ptr c) { cloud_ = c; }
    void setMaxIterations(int n) { max_iterations_ = n; }
    void setDistanceThreshold(double t) { threshold_ = t; }
    void setProbability(double p) { probability_ = p; }
    int iterations() const { return iterations_; }

    // PCL SampleConsensusModel::drawIndexSample: mt19937 seeded 12345 per model object, partial Fisher-Yates on a
    // persistent index list (the hypotheses are an INPUT of the parity definition: any fixed list is legitimate).
    static void draw_triplets(int64_t n, int H, std::vector<int32_t>& tri)
    {
        std::mt19937 rng(12345u);
        std::vector<int32_t> idx((size_t)n);
        for (int64_t i = 0; i < n; ++i) idx[(size_t)i] = (int32_t)i;
        tri.resize((size_t)H * 3);
        for (int h = 0; h < H; ++h) {
            for (int i = 0; i < 3; ++i) { const uint32_t r = rng() >> 1; std::swap(idx[(size_t)i], idx[(size_t)(i + (r % (uint64_t)(n - i)))]); }
            for (int i = 0; i < 3; ++i) tri[(size_t)h * 3 + i] = idx[(size_t)i];
        }
    }
    // pcl::RandomSampleConsensus::computeModel replayed over per-hypothesis inlier counts; returns the winner or -1
    static int replay(const std::vector<int32_t>& counts, const std::vector<int32_t>& valid, int64_t n, int max_iterations, double probability,
                      int* iterations_out)
    {
        const int H = (int)counts.size();
        int best = -INT_MAX, winner = -1, iter = 0, skipped = 0, h = 0;
        double k = 1.0;
        const double log_probability = std::log(1.0 - probability);
        const int max_skip = max_iterations * 10;
        while (iter < k && skipped < max_skip && h < H) {
            const int cur = h++;
            if (!valid[(size_t)cur]) { ++skipped; continue; }
            if (counts[(size_t)cur] > best) {
                best = counts[(size_t)cur]; winner = cur;
                const double w = (double)best / (double)n;
                double p_no_outliers = 1.0 - w * w * w;
                p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
                p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
                k = log_probability / std::log(p_no_outliers);
            }
            ++iter;
            if (iter > max_iterations) break;
        }
        if (iterations_out) *iterations_out = iter;
        return winner;
    }
    void segment(PointIndices& inliers, ModelCoefficients& model_coefficients)
    {
        inliers.indices.clear(); model_coefficients.values.clear();
        rclc_ctx* ctx = thread_ctx();
        if (!ctx || !cloud_ || cloud_->size() < 3) return;
        const int64_t n = (int64_t)cloud_->size();
        const int H = max_iterations_ + 1;
        std::vector<int32_t> tri;
        draw_triplets(n, H, tri);
        std::vector<float> planes((size_t)H * 4);
        std::vector<int32_t> valid((size_t)H), counts((size_t)H);
        if (!ok(rclc_ransac_plane_score(ctx, cloud_->points.data(), (int)sizeof(PoinT), n, tri.data(), H, threshold_, planes.data(), valid.data(), counts.data()),
                "SACSegmentation::segment (score)"))
            return;
        const int win = replay(counts, valid, n, max_iterations_, probability_, &iterations_);
        if (win < 0) return;
        std::vector<uint8_t> mask((size_t)n);
        float refined[4], centroid[3];
        int64_t n_in = 0;
        if (!ok(rclc_ransac_plane_select(ctx, cloud_->points.data(), (int)sizeof(PoinT), n, &planes[(size_t)win * 4], threshold_, mask.data(), refined, centroid, &n_in),
                "SACSegmentation::segment (select)"))
            return;
        // SACSegmentation::segment re-selects the inliers with the optimised coefficients (Appendix A)
        float unused4[4], unused3[3];
        if (!ok(rclc_ransac_plane_select(ctx, cloud_->points.data(), (int)sizeof(PoinT), n, refined, threshold_, mask.data(), unused4, unused3, &n_in),
                "SACSegmentation::segment (re-select)"))
            return;
        for (int64_t i = 0; i < n; ++i) if (mask[(size_t)i]) inliers.indices.push_back((int)i);
        model_coefficients.values.assign(refined, refined + 4);
    }
private:
    typename PointCloud<PoinT>::Ptr cloud_;
    int max_iterations_ = 800, iterations_ = 0;
    double threshold_ = 0.08, probability_ = 0.99;
};

// pcl::UniformSampling<PoinT> drop-in (the calls the reference makes: setInputCloud, setRadiusSearch, filter).
// Output order: ascending voxel key (PCL's is its hash map's iteration order). Occupied voxels and count are PCL's; which point of a
// voxel is kept follows SURVEY.md Appendix A (nearest the metric voxel centre) — see rclc_uniform_sample in rclc_b200.h.
class UniformSampling {
public:
    void setInputCloud(typename PointCloud<PoinT>::Ptr c) { cloud_ = c; }
    void setRadiusSearch(double r) { leaf_ = r; }
    void filter(PointCloud<PoinT>& out)
    {
        std::vector<PoinT> kept;
        rclc_ctx* ctx = thread_ctx();
        if (ctx && cloud_ && cloud_->size() > 0 && leaf_ > 0) {
            std::vector<int32_t> idx(cloud_->size());
            int64_t n_out = 0;
            if (ok(rclc_uniform_sample(ctx, cloud_->points.data(), (int)sizeof(PoinT), 16, (int64_t)cloud_->size(), leaf_, idx.data(), &n_out), "UniformSampling::filter")) {
                kept.reserve((size_t)n_out);
                for (int64_t k = 0; k < n_out; ++k) kept.push_back(cloud_->points[(size_t)idx[(size_t)k]]);
            }
        }
        out.points.assign(kept.begin(), kept.end());      // (out may be the input cloud, as at pc_process.h:115)
    }
private:
    typename PointCloud<PoinT>::Ptr cloud_;
    double leaf_ = 0;
};

// pcl::EuclideanClusterExtraction<PoinT> drop-in: clusters ranked by size descending, indices ascending inside a cluster.
class EuclideanClusterExtraction {
public:
    void setInputCloud(typename PointCloud<PoinT>::Ptr c) { cloud_ = c; }
    void setClusterTolerance(double t) { tol_ = t; }
    void setMinClusterSize(int64_t n) { min_ = n; }
    void setMaxClusterSize(int64_t n) { max_ = n; }
    template <class Tree> void setSearchMethod(const Tree&) {}      // the kd-tree is replaced by a uniform grid on device
    void extract(std::vector<PointIndices>& clusters)
    {
        clusters.clear();
        rclc_ctx* ctx = thread_ctx();
        if (!ctx || !cloud_ || cloud_->size() == 0) return;
        const int64_t n = (int64_t)cloud_->size();
        std::vector<int32_t> lab((size_t)n), sizes((size_t)n);
        int32_t nc = 0;
        if (!ok(rclc_euclidean_cluster(ctx, cloud_->points.data(), (int)sizeof(PoinT), 16, n, tol_, min_, max_, lab.data(), &nc, sizes.data()),
                "EuclideanClusterExtraction::extract"))
            return;
        clusters.resize((size_t)nc);
        for (int32_t k = 0; k < nc; ++k) clusters[(size_t)k].indices.reserve((size_t)sizes[(size_t)k]);
        for (int64_t i = 0; i < n; ++i) if (lab[(size_t)i] >= 0) clusters[(size_t)lab[(size_t)i]].indices.push_back((int)i);
    }
private:
    typename PointCloud<PoinT>::Ptr cloud_;
    double tol_ = 0;
    int64_t min_ = 1, max_ = INT_MAX;
};

// EulerCluster — include/pc_process.h:59-79: the largest cluster (tolerance 3 cm, 500 .. 2.5 M points) of target_clouds.
inline void EulerCluster(typename PointCloud<PoinT>::Ptr& target_clouds, typename PointCloud<PoinT>::Ptr& output)
{
    std::vector<PointIndices> ece_inlier;
    EuclideanClusterExtraction ece;
    ece.setInputCloud(target_clouds);
    ece.setClusterTolerance(0.03);
    ece.setMinClusterSize(500);
    ece.setMaxClusterSize(2500000);
    ece.extract(ece_inlier);
    if (!output) output = std::make_shared<PointCloud<PoinT>>();
    output->points.clear();
    if (ece_inlier.empty()) return;                       // (the reference indexes [0] unconditionally)
    for (int i : ece_inlier[0].indices) output->points.push_back(target_clouds->points[(size_t)i]);
}

// eular_cluster_iter — include/pc_process.h:104-211: centroids of the board's black blocks. Thins the cloud to 3 mm, then
// lowers the reflectivity threshold step by step until the black points fall apart into at least board_w * board_h / 2
// clusters; the first block_num clusters (by size, stopping at a 0.65 size drop) are the blocks, later ones are merged into
// the nearest block within half a square diagonal. max_rounds bounds the reference's unbounded loop.
inline void eular_cluster_iter(const typename PointCloud<PoinT>::Ptr& cloudInRaw, double reflactive_thd, std::vector<Vector3d>& block_centroid,
                               void* viewer = nullptr, int max_rounds = 200)
{
    (void)viewer;
    const rclc_config& cfg = CfgParam();
    auto blocks_cloud = std::make_shared<PointCloud<PoinT>>();
    UniformSampling US;
    US.setInputCloud(cloudInRaw);
    US.setRadiusSearch(cfg.block_us_radius);
    US.filter(*blocks_cloud);
    std::vector<std::vector<PoinT>> block_clouds_inlier;
    std::vector<int> block_centroid_inlier_num;
    const int block_num = cfg.board_width * cfg.board_height / 2;
    const double diag = cfg.gride_side_length * std::sqrt(2.0f);                      // src/global.cpp:47
    double curr = reflactive_thd / cfg.reflactivity_dec_step;
    for (int round = 0; round < max_rounds; ++round) {
        curr *= cfg.reflactivity_dec_step;
        {   // pcl::PassThrough on "intensity", [3, curr], inclusive
            std::vector<PoinT> kept;
            for (const PoinT& p : blocks_cloud->points)
                if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z) && p.intensity >= 3.0f && p.intensity <= (float)curr) kept.push_back(p);
            blocks_cloud->points.swap(kept);
        }
        if (blocks_cloud->size() == 0) return;
        std::vector<PointIndices> ece_inlier;
        EuclideanClusterExtraction ece;
        ece.setInputCloud(blocks_cloud);
        ece.setClusterTolerance(cfg.black_pt_cluster_dst_thd);
        ece.setMinClusterSize(3);
        ece.setMaxClusterSize(25000);
        ece.extract(ece_inlier);
        if ((int)ece_inlier.size() < block_num) continue;
        for (int i_seg = 0; i_seg < (int)ece_inlier.size(); ++i_seg) {
            std::vector<PoinT> ext;
            float cx = 0.f, cy = 0.f, cz = 0.f;                                        // pcl::compute3DCentroid: float accumulation
            for (int i : ece_inlier[(size_t)i_seg].indices) { const PoinT& p = blocks_cloud->points[(size_t)i]; ext.push_back(p); cx += p.x; cy += p.y; cz += p.z; }
            const float inv = 1.0f / (float)ext.size();
            const Vector3d centroid_curr((double)(cx * inv), (double)(cy * inv), (double)(cz * inv));
            if (i_seg < block_num) {
                if (i_seg > 0 && (double(ext.size()) / double(block_clouds_inlier[(size_t)i_seg - 1].size())) < 0.65) break;
                block_clouds_inlier.push_back(ext);
                block_centroid.push_back(centroid_curr);
                block_centroid_inlier_num.push_back((int)ext.size());
            } else {
                int nearest = -1;
                double nearest_dst = 10e3;
                for (int i = 0; i < (int)block_centroid.size(); ++i) {
                    const double dx = block_centroid[(size_t)i].x() - centroid_curr.x(), dy = block_centroid[(size_t)i].y() - centroid_curr.y(),
                                 dz = block_centroid[(size_t)i].z() - centroid_curr.z();
                    const float dst = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
                    if (dst < nearest_dst) { nearest_dst = dst; nearest = i; }
                }
                if (nearest >= 0 && nearest_dst < 0.5 * diag) {
                    Vector3d& old = block_centroid[(size_t)nearest];
                    const double a = block_centroid_inlier_num[(size_t)nearest], b = (double)ext.size();
                    old = Vector3d((old.x() * a + centroid_curr.x() * b) / (a + b), (old.y() * a + centroid_curr.y() * b) / (a + b),
                                   (old.z() * a + centroid_curr.z() * b) / (a + b));
                    block_centroid_inlier_num[(size_t)nearest] += (int)ext.size();
                    block_clouds_inlier[(size_t)nearest].insert(block_clouds_inlier[(size_t)nearest].end(), ext.begin(), ext.end());
                }
            }
        }
        if ((int)block_centroid.size() < block_num) {
            block_clouds_inlier.clear();
            block_centroid.clear();
            block_centroid_inlier_num.clear();
            continue;
        }
        return;
    }
}

// pcl::transformPointCloud(cloud_in, cloud_out, Matrix4f) as the reference calls it (in place: pc_process.h:594, 618).
inline void transformPointCloud(const PointCloud<PoinT>& in, PointCloud<PoinT>& out, const Matrix4f& T)
{
    if (&in != &out) out.points = in.points;
    rclc_ctx* ctx = thread_ctx();
    if (!ctx || out.size() == 0) return;
    float Tr[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) Tr[r * 4 + c] = T(r, c);
    ok(rclc_transform_cloud(ctx, out.points.data(), (int)sizeof(PoinT), (int64_t)out.size(), Tr), "transformPointCloud");
}

// calc_laser_coor_v1 — include/pc_process.h:373-462 (with get_ref_pt, utils.h:160, and line_fitting, opt_line_fitting.h:51):
// T_bl (board <- laser) from the black-block centroids and the board plane. Returns 0, or -1 when the centroids cannot be
// arranged into board_height rows of board_width / 2 blocks (the reference's calc_laser_coor prints and returns -1 there).
inline int calc_laser_coor_v1(std::vector<Vector3d>& v_grid_centroid, Vector4d& board_plane_param, Matrix4d& Tbl, void* viewer = nullptr)
{
    (void)viewer;
    std::vector<double> cen;
    for (const Vector3d& p : v_grid_centroid) { cen.push_back(p.x()); cen.push_back(p.y()); cen.push_back(p.z()); }
    const double plane[4] = {board_plane_param.x(), board_plane_param.y(), board_plane_param.z(), board_plane_param.w()};
    double T[16];
    if (!ok(rclc_calc_laser_coor(&CfgParam(), cen.data(), (int32_t)v_grid_centroid.size(), plane, T), "calc_laser_coor_v1")) return -1;
    from_row_major(T, Tbl);
    return 0;
}

// est_board_corner — include/pc_process.h:583-646: one board-only cloud (laser frame, after T_base) in, the (W-1) x (H-1)
// board corners in the laser frame out. The cloud is modified like in the reference (projected onto its plane, then moved into
// the board frame). Returns false when a stage gives up (the reference carries on and crashes or produces garbage there).
inline bool est_board_corner(typename PointCloud<PoinT>::Ptr& chess_board_processe, std::vector<Vector3d>& board_corner, void* viewer = nullptr,
                             rclc_gridfit_report* report = nullptr)
{
    (void)viewer;
    const rclc_config& cfg = CfgParam();
    if (!chess_board_processe || chess_board_processe->size() == 0) return false;
    if (cfg.board_order == 1) {
        Matrix4f T_grid_shift;
        T_grid_shift(0, 0) = -1.0f;
        transformPointCloud(*chess_board_processe, *chess_board_processe, T_grid_shift);
    }
    Vector4d board_plane_param;
    pc_plane_fitting(chess_board_processe, board_plane_param);
    pc_plane_prj(chess_board_processe, board_plane_param);
    VectorXd board_intensity((long)chess_board_processe->size());
    for (long i = 0; i < (long)chess_board_processe->size(); ++i) board_intensity(i) = chess_board_processe->points[(size_t)i].intensity;
    std::vector<double> mu{15, 100}, sigma{15, 15};
    GMMFit::fit_1d(board_intensity, 2, mu, sigma, false);
    const double reflactive_thd = mu[0] + 2.0 * sigma[0];
    std::vector<Vector3d> block_centroid;
    eular_cluster_iter(chess_board_processe, reflactive_thd, block_centroid);
    if ((int)block_centroid.size() < cfg.board_width * cfg.board_height / 2) return false;
    Matrix4d T_bl;
    if (calc_laser_coor_v1(block_centroid, board_plane_param, T_bl) != 0) return false;
    Matrix4f T_blf;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T_blf(r, c) = (float)T_bl(r, c);
    transformPointCloud(*chess_board_processe, *chess_board_processe, T_blf);
    const size_t before = board_corner.size();
    opt_grid_fitting(chess_board_processe, T_bl, board_corner, nullptr, report);
    if (board_corner.size() == before) return false;
    if (cfg.board_order == 1) {
        const int n_pt_per_row = cfg.board_width - 1;
        std::vector<Vector3d> tmp(board_corner.size());
        for (int row = 0; row < cfg.board_height - 1; ++row)
            for (int col = 0; col < n_pt_per_row; ++col) {
                const Vector3d& pt = board_corner[(size_t)(row * n_pt_per_row + col)];
                tmp[(size_t)(row * n_pt_per_row + n_pt_per_row - col - 1)] = Vector3d(-pt.x(), pt.y(), pt.z());
            }
        board_corner = tmp;
    }
    return true;
}

// pcl::PassThrough<PoinT> as the reference uses it (field "x" / "y" / "z" / "intensity", inclusive limits, optional negative):
// host code — a compaction of a few hundred thousand points is not worth a round trip.
class PassThrough {
public:
    void setInputCloud(typename PointCloud<PoinT>::Ptr c) { cloud_ = c; }
    void setFilterFieldName(const std::string& f) { field_ = f; }
    void setFilterLimits(double lo, double hi) { lo_ = (float)lo; hi_ = (float)hi; }
    void setFilterLimitsNegative(bool n) { negative_ = n; }
    void filter(PointCloud<PoinT>& out)
    {
        std::vector<PoinT> kept;
        if (cloud_)
            for (const PoinT& p : cloud_->points) {
                if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
                const float v = field_ == "x" ? p.x : field_ == "y" ? p.y : field_ == "z" ? p.z : p.intensity;
                const bool in = std::isfinite(v) && v >= lo_ && v <= hi_;
                if (in != negative_) kept.push_back(p);
            }
        out.points.swap(kept);
    }
private:
    typename PointCloud<PoinT>::Ptr cloud_;
    std::string field_ = "x";
    float lo_ = -3.4e38f, hi_ = 3.4e38f;
    bool negative_ = false;
};

// cut_roi — include/pc_process.h:81-101: keeps the points of raw_board inside the axis-aligned box [min_p, max_p].
inline void cut_roi(typename PointCloud<PoinT>::Ptr& raw_board, PoinT min_p, PoinT max_p)
{
    PassThrough pass;
    const char* names[3] = {"x", "y", "z"};
    const float lo[3] = {min_p.x, min_p.y, min_p.z}, hi[3] = {max_p.x, max_p.y, max_p.z};
    for (int d = 0; d < 3; ++d) {
        pass.setInputCloud(raw_board);
        pass.setFilterFieldName(names[d]);
        pass.setFilterLimits(lo[d], hi[d]);
        pass.setFilterLimitsNegative(false);
        pass.filter(*raw_board);
    }
}

// segment_board — the per-cloud body of apps/BoardSegmentation.cpp:45-165: the chessboard's points out of one raw scan (lidar
// frame, x forward). ROI x in [3, 6] m, 8 mm uniform sampling, planes peeled off by RANSAC until 5 % of the points are left,
// the nearest large plane that faces the lidar is the board candidate; its largest cluster gives the ROI box of the raw scan;
// a GMM on the candidate's reflectance splits black from white, each side keeps its largest cluster. Returns false when no
// plane qualifies (the reference goes on with an empty cloud).
inline bool segment_board(const typename PointCloud<PoinT>::Ptr& raw_cloud_read, typename PointCloud<PoinT>::Ptr& final_plane,
                          double x_min = 3.0, double x_max = 6.0)
{
    const rclc_config& cfg = CfgParam();
    auto raw_cloud = std::make_shared<PointCloud<PoinT>>();
    {
        PassThrough pass;
        pass.setInputCloud(raw_cloud_read);
        pass.setFilterFieldName("x");
        pass.setFilterLimits(x_min, x_max);                                       // BoardSegmentation.cpp:48
        pass.filter(*raw_cloud);
    }
    auto vox_cloud = std::make_shared<PointCloud<PoinT>>();
    UniformSampling US0;
    US0.setInputCloud(raw_cloud);
    US0.setRadiusSearch(0.008f);                                                  // :60
    US0.filter(*vox_cloud);
    PlaneSegmentation sac;
    sac.setMaxIterations(800);                                                    // :71-72
    sac.setDistanceThreshold(0.08);
    auto target_plane = std::make_shared<PointCloud<PoinT>>();
    const size_t pt_size = vox_cloud->size();
    double closest_plane_dist = 1000;
    while (vox_cloud->size() > pt_size * 0.05) {
        PointIndices inliner;
        ModelCoefficients coefficients;
        sac.setInputCloud(vox_cloud);
        sac.segment(inliner, coefficients);
        if (coefficients.values.size() != 4 || inliner.indices.empty()) break;   // (PCL prints an error and the reference would spin)
        std::vector<uint8_t> is_in(vox_cloud->size(), 0);
        for (int i : inliner.indices) is_in[(size_t)i] = 1;
        std::vector<PoinT> ext, rest;
        for (size_t i = 0; i < vox_cloud->size(); ++i) (is_in[i] ? ext : rest).push_back(vox_cloud->points[i]);
        vox_cloud->points.swap(rest);
        const double a = coefficients.values[0], b = coefficients.values[1], c = coefficients.values[2];
        const double nx = a / std::sqrt(a * a + b * b + c * c);
        if (ext.size() < pt_size * 0.02 || std::fabs(nx) <= cfg.angle_thd) continue;          // :101
        float sx = 0.f, sy = 0.f, sz = 0.f;                                        // pcl::compute3DCentroid
        for (const PoinT& p : ext) { sx += p.x; sy += p.y; sz += p.z; }
        const float inv = 1.0f / (float)ext.size();
        const double plane_distance = std::sqrt((double)(sx * inv) * (sx * inv) + (double)(sy * inv) * (sy * inv) + (double)(sz * inv) * (sz * inv));
        if ((closest_plane_dist - plane_distance) > 0.3) {                         // :107
            closest_plane_dist = plane_distance;
            target_plane->points = ext;
        }
    }
    if (target_plane->size() == 0) return false;
    EulerCluster(target_plane, target_plane);
    if (target_plane->size() == 0) return false;
    PoinT min_p = target_plane->points[0], max_p = target_plane->points[0];        // pcl::getMinMax3D
    for (const PoinT& p : target_plane->points) {
        min_p.x = std::min(min_p.x, p.x); min_p.y = std::min(min_p.y, p.y); min_p.z = std::min(min_p.z, p.z);
        max_p.x = std::max(max_p.x, p.x); max_p.y = std::max(max_p.y, p.y); max_p.z = std::max(max_p.z, p.z);
    }
    cut_roi(raw_cloud, min_p, max_p);
    VectorXd board_intensity((long)target_plane->size());
    for (long i = 0; i < (long)target_plane->size(); ++i) board_intensity(i) = target_plane->points[(size_t)i].intensity;
    std::vector<double> mu{15, 100}, sigma{15, 15};
    GMMFit::fit_1d(board_intensity, 2, mu, sigma, false);
    const double reflactive_thd = mu[0] + 1.7 * sigma[0];                          // :134
    auto black = std::make_shared<PointCloud<PoinT>>(), white = std::make_shared<PointCloud<PoinT>>();
    PassThrough pass_intensity;
    pass_intensity.setInputCloud(raw_cloud);
    pass_intensity.setFilterFieldName("intensity");
    pass_intensity.setFilterLimits(cfg.noise_reflactivity_thd, reflactive_thd);
    pass_intensity.filter(*black);
    EulerCluster(black, black);
    pass_intensity.setInputCloud(raw_cloud);
    pass_intensity.setFilterLimits(reflactive_thd, 150);
    pass_intensity.filter(*white);
    EulerCluster(white, white);
    if (!final_plane) final_plane = std::make_shared<PointCloud<PoinT>>();
    final_plane->points = white->points;                                           // *final_plane = *white + *black (:160)
    final_plane->points.insert(final_plane->points.end(), black->points.begin(), black->points.end());
    return final_plane->size() > 0;
}

}  // namespace rclc_b200
#endif  // RCLC_OPS_HPP
