// pnp.cu — PnP initialisation without OpenCV (SURVEY.md §8f "next #3"): host code, a few hundred correspondences.
//
// Replaces cv::solvePnPRansac(objectPoints, imagePoints, K, dist, rvec, tvec, false, SOLVEPNP_UPNP) at apps/Calibrate.cpp:41-50
// (OpenCV >= 3 runs EPnP for that flag) for the one thing the reference uses it for: a starting pose for pose_reprj_opt
// (apps/Calibrate.cpp:57-68). It works on NORMALISED image points (Cam.imgPt2norm, which the reference computes anyway at :36),
// so no intrinsics enter here. RANSAC as in OpenCV (100 iterations by default, adaptive stop at the given confidence), model from
// a linear solve — DLT for general samples, a plane-induced homography when the sample is coplanar (one board) — polished by
// Gauss-Newton on the reprojection error; the winner is re-fitted on its inliers. The random sequence is this library's own
// (mt19937 with the caller's seed): parity with OpenCV is on the resulting pose, which tests/ compare against cv2.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <random>

namespace rclc {
namespace {

// cyclic Jacobi for a symmetric n x n matrix (n <= 12): A = V diag(w) V^T, columns of V are eigenvectors
void jacobi_eig(int n, double* A, double* w, double* V)
{
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) { const double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
                for (int k = 0; k < n; ++k) { const double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
                for (int k = 0; k < n; ++k) { const double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

int argmin(const double* w, int n) { int k = 0; for (int i = 1; i < n; ++i) if (w[i] < w[k]) k = i; return k; }

double det3(const double* M) { return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]); }

// nearest rotation to M (polar factor): R = M (M^T M)^(-1/2)
bool nearest_rotation(const double* M, double* R)
{
    double S[9], w[3], V[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += M[k * 3 + i] * M[k * 3 + j]; S[i * 3 + j] = s; }
    jacobi_eig(3, S, w, V);
    for (int i = 0; i < 3; ++i) if (!(w[i] > 1e-300)) return false;
    double Q[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += V[i * 3 + k] * V[j * 3 + k] / std::sqrt(w[k]); Q[i * 3 + j] = s; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += M[i * 3 + k] * Q[k * 3 + j]; R[i * 3 + j] = s; }
    if (det3(R) < 0) return false;
    return true;
}

struct Pose { double R[9], t[3]; };

void project(const Pose& P, const double* X, double& u, double& v, double& z)
{
    const double x = P.R[0] * X[0] + P.R[1] * X[1] + P.R[2] * X[2] + P.t[0], y = P.R[3] * X[0] + P.R[4] * X[1] + P.R[5] * X[2] + P.t[1];
    z = P.R[6] * X[0] + P.R[7] * X[1] + P.R[8] * X[2] + P.t[2];
    u = x / z; v = y / z;
}

// Gauss-Newton (lightly damped) on the normalised reprojection error over the points idx[0..m)
void polish(const double* obj, const double* uv, const int* idx, int m, Pose& P, int iters)
{
    for (int it = 0; it < iters; ++it) {
        double H[36] = {0}, g[6] = {0};
        for (int k = 0; k < m; ++k) {
            const double* X = obj + 3 * idx[k];
            const double* q = uv + 2 * idx[k];
            const double x = P.R[0] * X[0] + P.R[1] * X[1] + P.R[2] * X[2] + P.t[0], y = P.R[3] * X[0] + P.R[4] * X[1] + P.R[5] * X[2] + P.t[1],
                         z = P.R[6] * X[0] + P.R[7] * X[1] + P.R[8] * X[2] + P.t[2];
            if (!(z > 1e-9)) continue;
            const double iz = 1.0 / z, ru = x * iz - q[0], rv = y * iz - q[1];
            // d(x,y,z)/d(omega) = -[p]_x with p = R X + t - t (rotation about the camera origin acts on R X), d/d(tau) = I
            const double px = x - P.t[0], py = y - P.t[1], pz = z - P.t[2];
            const double Ju[6] = {iz * (0) - x * iz * iz * (py), iz * (pz) - x * iz * iz * (-px), iz * (-py) - x * iz * iz * 0, iz, 0, -x * iz * iz};
            const double Jv[6] = {iz * (-pz) - y * iz * iz * (py), iz * 0 - y * iz * iz * (-px), iz * (px) - y * iz * iz * 0, 0, iz, -y * iz * iz};
            for (int a = 0; a < 6; ++a) { g[a] += Ju[a] * ru + Jv[a] * rv; for (int b = 0; b < 6; ++b) H[a * 6 + b] += Ju[a] * Ju[b] + Jv[a] * Jv[b]; }
        }
        for (int a = 0; a < 6; ++a) H[a * 6 + a] += 1e-9 + 1e-6 * H[a * 6 + a];
        // Cholesky solve H d = -g
        double L[36] = {0}, y6[6], d[6];
        bool ok = true;
        for (int i = 0; i < 6 && ok; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = H[i * 6 + j];
                for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
                if (i == j) { if (!(s > 0)) { ok = false; break; } L[i * 6 + i] = std::sqrt(s); } else L[i * 6 + j] = s / L[j * 6 + j];
            }
        if (!ok) return;
        for (int i = 0; i < 6; ++i) { double s = -g[i]; for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y6[k]; y6[i] = s / L[i * 6 + i]; }
        for (int i = 5; i >= 0; --i) { double s = y6[i]; for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * d[k]; d[i] = s / L[i * 6 + i]; }
        // R <- exp([omega]_x) R, t <- t + tau
        const double th = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        double E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (th > 1e-14) {
            const double kx = d[0] / th, ky = d[1] / th, kz = d[2] / th, c = std::cos(th), s = std::sin(th), C = 1 - c;
            const double Ex[9] = {c + kx * kx * C, kx * ky * C - kz * s, kx * kz * C + ky * s, ky * kx * C + kz * s, c + ky * ky * C, ky * kz * C - kx * s,
                                  kz * kx * C - ky * s, kz * ky * C + kx * s, c + kz * kz * C};
            std::copy(Ex, Ex + 9, E);
        }
        double Rn[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += E[i * 3 + k] * P.R[k * 3 + j]; Rn[i * 3 + j] = s; }
        std::copy(Rn, Rn + 9, P.R);
        for (int k = 0; k < 3; ++k) P.t[k] += d[3 + k];
        if (th < 1e-12 && std::fabs(d[3]) + std::fabs(d[4]) + std::fabs(d[5]) < 1e-12) break;
    }
}

// centroid, principal axes and spread of the sample: the frame in which the linear solves are conditioned
struct Frame3 { double c[3], ax[9], ev[3]; };      // ax rows: e1, e2, n (ascending planarity: ev[2] smallest)
Frame3 sample_frame(const double* obj, const int* idx, int m)
{
    Frame3 F;
    for (int d = 0; d < 3; ++d) { double s = 0; for (int k = 0; k < m; ++k) s += obj[3 * idx[k] + d]; F.c[d] = s / m; }
    double C[9] = {0}, w[3], V[9];
    for (int k = 0; k < m; ++k) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i * 3 + j] += (obj[3 * idx[k] + i] - F.c[i]) * (obj[3 * idx[k] + j] - F.c[j]) / m;
    jacobi_eig(3, C, w, V);
    int o[3] = {0, 1, 2};
    std::sort(o, o + 3, [&](int a, int b) { return w[a] > w[b]; });
    for (int r = 0; r < 3; ++r) { F.ev[r] = w[o[r]]; for (int d = 0; d < 3; ++d) F.ax[r * 3 + d] = V[d * 3 + o[r]]; }
    // right-handed
    const double n[3] = {F.ax[1] * F.ax[5] - F.ax[2] * F.ax[4], F.ax[2] * F.ax[3] - F.ax[0] * F.ax[5], F.ax[0] * F.ax[4] - F.ax[1] * F.ax[3]};
    for (int d = 0; d < 3; ++d) F.ax[6 + d] = n[d];
    return F;
}

// pose from m >= 6 points in general position: DLT on [R | t] in the sample's own frame
bool pose_dlt(const double* obj, const double* uv, const int* idx, int m, const Frame3& F, Pose& P)
{
    const double scale = 1.0 / std::sqrt(std::max(F.ev[0], 1e-300));
    double A[144] = {0}, w[12], V[144];
    for (int k = 0; k < m; ++k) {
        double Y[3];
        for (int r = 0; r < 3; ++r) { double s = 0; for (int d = 0; d < 3; ++d) s += F.ax[r * 3 + d] * (obj[3 * idx[k] + d] - F.c[d]); Y[r] = s * scale; }
        const double u = uv[2 * idx[k]], v = uv[2 * idx[k] + 1];
        const double r1[12] = {Y[0], Y[1], Y[2], 1, 0, 0, 0, 0, -u * Y[0], -u * Y[1], -u * Y[2], -u};
        const double r2[12] = {0, 0, 0, 0, Y[0], Y[1], Y[2], 1, -v * Y[0], -v * Y[1], -v * Y[2], -v};
        for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) A[i * 12 + j] += r1[i] * r1[j] + r2[i] * r2[j];
    }
    jacobi_eig(12, A, w, V);
    const int k0 = argmin(w, 12);
    double p[12];
    for (int i = 0; i < 12; ++i) p[i] = V[i * 12 + k0];
    double M[9] = {p[0], p[1], p[2], p[4], p[5], p[6], p[8], p[9], p[10]};
    double dt = det3(M);
    if (std::fabs(dt) < 1e-300) return false;
    const double s = (dt > 0 ? 1.0 : -1.0) / std::cbrt(std::fabs(dt));
    for (int i = 0; i < 9; ++i) M[i] *= s;
    double Rl[9];
    if (!nearest_rotation(M, Rl)) return false;
    const double tl[3] = {p[3] * s, p[7] * s, p[11] * s};
    // back to the caller's frame: Y = scale * ax (X - c)  =>  R = Rl * scale * ax ... the scale belongs to t: x = Rl Y + tl = (Rl ax) scale (X - c) + tl
    // dividing the homogeneous point by scale: R = Rl ax, t = tl / scale - R c
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += Rl[i * 3 + k] * F.ax[k * 3 + j]; P.R[i * 3 + j] = a; }
    for (int i = 0; i < 3; ++i) { double a = tl[i] / scale; for (int j = 0; j < 3; ++j) a -= P.R[i * 3 + j] * F.c[j]; P.t[i] = a; }
    return true;
}

// pose from m >= 4 coplanar points: homography of the plane (a, b, 1) -> (u, v, 1)
bool pose_planar(const double* obj, const double* uv, const int* idx, int m, const Frame3& F, Pose& P)
{
    const double scale = 1.0 / std::sqrt(std::max(F.ev[0], 1e-300));
    double A[81] = {0}, w[9], V[81];
    for (int k = 0; k < m; ++k) {
        double a = 0, b = 0;
        for (int d = 0; d < 3; ++d) { a += F.ax[d] * (obj[3 * idx[k] + d] - F.c[d]); b += F.ax[3 + d] * (obj[3 * idx[k] + d] - F.c[d]); }
        a *= scale; b *= scale;
        const double u = uv[2 * idx[k]], v = uv[2 * idx[k] + 1];
        const double r1[9] = {a, b, 1, 0, 0, 0, -u * a, -u * b, -u}, r2[9] = {0, 0, 0, a, b, 1, -v * a, -v * b, -v};
        for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) A[i * 9 + j] += r1[i] * r1[j] + r2[i] * r2[j];
    }
    jacobi_eig(9, A, w, V);
    const int k0 = argmin(w, 9);
    double h[9];
    for (int i = 0; i < 9; ++i) h[i] = V[i * 9 + k0];
    // H = [h1 h2 h3] ~ [r1 r2 t]
    const double n1 = std::sqrt(h[0] * h[0] + h[3] * h[3] + h[6] * h[6]), n2 = std::sqrt(h[1] * h[1] + h[4] * h[4] + h[7] * h[7]);
    if (!(n1 > 1e-300) || !(n2 > 1e-300)) return false;
    double lam = 2.0 / (n1 + n2);
    if (h[8] * lam < 0) lam = -lam;                                   // the plane is in front of the camera
    const double r1[3] = {h[0] * lam, h[3] * lam, h[6] * lam}, r2[3] = {h[1] * lam, h[4] * lam, h[7] * lam};
    const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double M[9] = {r1[0], r2[0], r3[0], r1[1], r2[1], r3[1], r1[2], r2[2], r3[2]};
    double Rl[9];
    if (!nearest_rotation(M, Rl)) return false;
    const double tl[3] = {h[2] * lam, h[5] * lam, h[8] * lam};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += Rl[i * 3 + k] * F.ax[k * 3 + j]; P.R[i * 3 + j] = a; }
    for (int i = 0; i < 3; ++i) { double a = tl[i] / scale; for (int j = 0; j < 3; ++j) a -= P.R[i * 3 + j] * F.c[j]; P.t[i] = a; }
    return true;
}

bool fit(const double* obj, const double* uv, const int* idx, int m, Pose& P)
{
    const Frame3 F = sample_frame(obj, idx, m);
    const bool planar = F.ev[2] < 1e-6 * F.ev[0] || m < 6;
    bool ok = planar ? pose_planar(obj, uv, idx, m, F, P) : pose_dlt(obj, uv, idx, m, F, P);
    if (!ok && !planar) ok = pose_planar(obj, uv, idx, m, F, P);
    if (!ok) return false;
    polish(obj, uv, idx, m, P, 10);
    for (int k = 0; k < 12; ++k) if (!std::isfinite(k < 9 ? P.R[k] : P.t[k - 9])) return false;
    return true;
}

} // namespace
} // namespace rclc

using namespace rclc;

extern "C" int rclc_pnp_ransac(const double* obj, const double* norm_pts, int64_t n, double thresh, int32_t max_iters, double confidence, uint32_t seed,
                               double R_out[9], double t_out[3], uint8_t* inlier_mask, int32_t* n_inliers)
{
    RCLC_CHECK(obj && norm_pts && R_out && t_out && n >= 4 && n < (1 << 30) && thresh > 0, RCLC_ERR_INVALID, "bad argument");
    if (max_iters <= 0) max_iters = 100;                              // cv::solvePnPRansac defaults: 100 iterations, confidence 0.99
    if (!(confidence > 0 && confidence < 1)) confidence = 0.99;
    const int N = (int)n, ms = N >= 6 ? 6 : 4;
    std::vector<int> all((size_t)N);
    for (int i = 0; i < N; ++i) all[(size_t)i] = i;
    std::mt19937 rng(seed);
    auto count = [&](const Pose& P, std::vector<uint8_t>* mask) {
        int c = 0;
        for (int i = 0; i < N; ++i) {
            double u, v, z;
            project(P, obj + 3 * i, u, v, z);
            const double du = u - norm_pts[2 * i], dv = v - norm_pts[2 * i + 1];
            const bool in = z > 0 && du * du + dv * dv <= thresh * thresh;
            if (mask) (*mask)[(size_t)i] = in;
            c += in;
        }
        return c;
    };
    Pose best;
    int best_cnt = -1;
    int iters = max_iters;
    std::vector<int> sample((size_t)ms);
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < ms; ++k) {                                // partial Fisher-Yates
            std::uniform_int_distribution<int> d(k, N - 1);
            std::swap(all[(size_t)k], all[(size_t)d(rng)]);
            sample[(size_t)k] = all[(size_t)k];
        }
        Pose P;
        if (!fit(obj, norm_pts, sample.data(), ms, P)) continue;
        const int c = count(P, nullptr);
        if (c > best_cnt) {
            best_cnt = c; best = P;
            const double w = (double)c / N, denom = std::log(std::max(1e-300, 1.0 - std::pow(w, ms)));
            if (denom < 0) iters = std::min(iters, std::max(it + 1, (int)std::ceil(std::log(1.0 - confidence) / denom)));
        }
    }
    if (best_cnt < ms) {                                              // no sample worked: all points at once
        if (!fit(obj, norm_pts, all.data(), N, best)) { set_last_error("PnP: no pose found"); return RCLC_ERR_NUMERIC; }
        best_cnt = count(best, nullptr);
    }
    // re-fit on the inliers of the winner (twice: the second pass picks up the points the better pose explains)
    std::vector<uint8_t> mask((size_t)N);
    for (int pass = 0; pass < 2; ++pass) {
        count(best, &mask);
        std::vector<int> in;
        for (int i = 0; i < N; ++i) if (mask[(size_t)i]) in.push_back(i);
        if ((int)in.size() < ms) break;
        Pose P = best;
        if (fit(obj, norm_pts, in.data(), (int)in.size(), P) && count(P, nullptr) >= best_cnt) { best = P; best_cnt = count(P, nullptr); }
        else { polish(obj, norm_pts, in.data(), (int)in.size(), best, 10); best_cnt = count(best, nullptr); }
    }
    best_cnt = count(best, &mask);
    std::copy(best.R, best.R + 9, R_out);
    std::copy(best.t, best.t + 3, t_out);
    if (inlier_mask) std::copy(mask.begin(), mask.end(), inlier_mask);
    if (n_inliers) *n_inliers = best_cnt;
    return RCLC_OK;
}

// Eigen::Quaterniond(Matrix3d) (apps/Calibrate.cpp:61): Shepperd's branches as in Eigen/src/Geometry/Quaternion.h. R row-major,
// q = (x, y, z, w), the storage order of the reference's Tcl (Calibrate.cpp:59-62).
extern "C" int rclc_quat_from_R(const double R[9], double q_xyzw[4])
{
    RCLC_CHECK(R && q_xyzw, RCLC_ERR_INVALID, "null argument");
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = std::sqrt(tr + 1.0);
        q_xyzw[3] = 0.5 * s;
        s = 0.5 / s;
        q_xyzw[0] = (R[7] - R[5]) * s; q_xyzw[1] = (R[2] - R[6]) * s; q_xyzw[2] = (R[3] - R[1]) * s;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q_xyzw[i] = 0.5 * s;
        s = 0.5 / s;
        q_xyzw[3] = (R[3 * k + j] - R[3 * j + k]) * s;
        q_xyzw[j] = (R[3 * j + i] + R[3 * i + j]) * s;
        q_xyzw[k] = (R[3 * k + i] + R[3 * i + k]) * s;
    }
    return RCLC_OK;
}
