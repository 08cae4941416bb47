// BASELINE configs[0] in miniature, through the reference-shaped operators only: raw lidar scans of a scene with a chessboard
//   -> segment_board (apps/BoardSegmentation.cpp) -> T_base -> est_board_corner (apps/Calibrate.cpp:90-103)
//   -> pnp_init + pose_reprj_opt (calib_opt, Calibrate.cpp:14-68) -> extrinsic, compared with the ground truth of the synthetic rig.
// Exit 0 = all checks passed, 77 = no CUDA device (nothing was computed).
#include <cmath>
#include <cstdio>
#include <random>

#include "rclc_pcd.hpp"
#include "rclc_oracle.h"      // the CPU oracle: the checker of every stage below (tests may link it, the product may not)

using namespace rclc_b200;
static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } } while (0)

static void rot_xyz(double ax, double ay, double az, double* R)
{
    const double cx = std::cos(ax), sx = std::sin(ax), cy = std::cos(ay), sy = std::sin(ay), cz = std::cos(az), sz = std::sin(az);
    const double M[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, -sy, cy * sx, cy * cx};
    for (int i = 0; i < 9; ++i) R[i] = M[i];
}

int main()
{
    if (!thread_ctx()) { std::printf("no CUDA device: nothing computed (there is no CPU fallback)\n"); return 77; }
    const int F = 6;
    std::mt19937 rng(2024);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> N01(0.0, 1.0);
    // ground-truth extrinsic camera <- (T_base'd) laser frame
    double Rcl[9];
    rot_xyz(0.02, -0.03, 0.015, Rcl);
    const double tcl[3] = {0.07, -0.05, 0.04};
    std::vector<std::vector<Vector3d>> pcs, pcs_oracle;
    std::vector<std::vector<Vector2d>> ims;
    const Matrix4f Tb = io::T_base();
    for (int f = 0; f < F; ++f) {
        // the board in the camera-like frame (u right, v down, w forward), top-right square black
        double Rb[9];
        rot_xyz(0.12 * (U(rng) - 0.5), 0.3 * (U(rng) - 0.5), 0.16 * (U(rng) - 0.5), Rb);
        const double tb[3] = {0.8 * (U(rng) - 0.5), 0.4 * (U(rng) - 0.5), 4.0 + 1.0 * U(rng)};
        auto board_to_cam = [&](double a, double b, double c, double* o) { for (int r = 0; r < 3; ++r) o[r] = Rb[3 * r] * a + Rb[3 * r + 1] * b + Rb[3 * r + 2] * c + tb[r]; };
        auto raw = std::make_shared<PointCloud<PoinT>>();
        auto push_cam = [&](const double* o, double I) { PoinT p; p.x = (float)o[2]; p.y = (float)-o[0]; p.z = (float)-o[1]; p.intensity = (float)I; raw->points.push_back(p); };   // inverse of T_base
        const size_t n_board = 100000;
        for (size_t i = 0; i < n_board; ++i) {
            const double a = -0.4 + 0.8 * U(rng), b = -0.3 + 0.6 * U(rng);
            const double fu = (a + 0.4) / 0.1, fv = (b + 0.3) / 0.1;
            const int ci = std::min(7, (int)std::floor(fu)), ri = std::min(5, (int)std::floor(fv));
            const double e = std::min(std::min(fu - ci, ci + 1 - fu), std::min(fv - ri, ri + 1 - fv)) * 0.1;
            const bool black = ((ci + ri) % 2) == 1;
            double o[3];
            board_to_cam(a, b, 0.002 * N01(rng), o);
            push_cam(o, std::max(4.0, (black ? 10.0 + 60.0 * std::max(0.0, 1.0 - e / 0.012) : 110.0) + 2.0 * N01(rng)));
        }
        // ground (z_l = -1.2), back wall (x_l = 5.9) and scattered clutter, all inside the x in [3, 6] ROI
        for (int i = 0; i < 40000; ++i) { PoinT p; p.x = (float)(3.0 + 3.0 * U(rng)); p.y = (float)(-2.5 + 5.0 * U(rng)); p.z = (float)(-1.2 + 0.01 * N01(rng)); p.intensity = 40.f; raw->points.push_back(p); }
        for (int i = 0; i < 40000; ++i) { PoinT p; p.x = (float)(5.9 + 0.01 * N01(rng)); p.y = (float)(-2.5 + 5.0 * U(rng)); p.z = (float)(-1.2 + 2.7 * U(rng)); p.intensity = 60.f; raw->points.push_back(p); }
        for (int i = 0; i < 5000; ++i) { PoinT p; p.x = (float)(1.0 + 8.0 * U(rng)); p.y = (float)(-3.0 + 6.0 * U(rng)); p.z = (float)(-1.2 + 3.0 * U(rng)); p.intensity = (float)(150.0 * U(rng)); raw->points.push_back(p); }

        // ---- stage 1 ----
        PointCloud<PoinT>::Ptr board;
        const bool found = segment_board(raw, board);
        CHECK(found && board && board->size() > 0.85 * n_board && board->size() < 1.02 * n_board, "frame %d: board cloud %zu of %zu", f, board ? board->size() : 0, n_board);
        if (!found || !board) continue;
        {   // oracle: the same chain on the CPU must return the same points in the same order
            std::vector<float> o4(raw->size() * 4);
            int64_t n_o = 0;
            const int rc = orc_segment_board(&CfgParam(), raw->points.data(), (int)sizeof(PoinT), 16, (int64_t)raw->size(), 3.0, 6.0, o4.data(), &n_o);
            bool same = rc == 0 && (size_t)n_o == board->size();
            for (size_t i = 0; same && i < board->size(); ++i)
                same = board->points[i].x == o4[4 * i] && board->points[i].y == o4[4 * i + 1] && board->points[i].z == o4[4 * i + 2] && board->points[i].intensity == o4[4 * i + 3];
            CHECK(same, "frame %d: segment_board differs from the oracle (rc %d, %lld vs %zu points)", f, rc, (long long)n_o, board->size());
        }
        // the hand-off: PCD round trip, then the axis permutation of Calibrate.cpp:98-99
        CHECK(io::savePCDFileBinary("/tmp/rclc_pipeline_board.pcd", *board) == 0, "save");
        auto chess_board = std::make_shared<PointCloud<PoinT>>();
        CHECK(io::loadPCDFile("/tmp/rclc_pipeline_board.pcd", *chess_board) == 0 && chess_board->size() == board->size(), "load");
        transformPointCloud(*chess_board, *chess_board, Tb);

        // ---- stage 2, per frame ----
        std::vector<float> cb4(chess_board->size() * 4);            // the oracle's copy of the input (est_board_corner works in place)
        for (size_t i = 0; i < chess_board->size(); ++i) { const PoinT& p = chess_board->points[i]; cb4[4 * i] = p.x; cb4[4 * i + 1] = p.y; cb4[4 * i + 2] = p.z; cb4[4 * i + 3] = p.intensity; }
        std::vector<Vector3d> corners;
        const bool ok_c = est_board_corner(chess_board, corners);
        {
            std::vector<double> oc(35 * 3);
            orc_gridfit_report orep;
            const int rc = orc_est_board_corner(&CfgParam(), cb4.data(), (int64_t)(cb4.size() / 4), oc.data(), &orep);
            double dmax = 0;
            if (rc == 0 && corners.size() == 35)
                for (int k = 0; k < 35; ++k)
                    dmax = std::max(dmax, std::max(std::fabs(corners[(size_t)k].x() - oc[3 * k]), std::max(std::fabs(corners[(size_t)k].y() - oc[3 * k + 1]), std::fabs(corners[(size_t)k].z() - oc[3 * k + 2]))));
            // The plane fit returns the oracle's plane bit for bit (exact sums), so the projected floats, the block centroids and the
            // grid fit's whole LM path are the oracle's: the corners agree to the last digits of the final double-precision transform.
            // (An upstream agreement of "1e-9" is not enough: tests/test_chain_sensitivity.py shows the oracle itself forking by
            // 0.1 - 0.5 mm under such a perturbation.)
            CHECK(rc == 0 && corners.size() == 35 && dmax < 2e-6, "frame %d: est_board_corner vs oracle: rc %d, max |d| %g m", f, rc, dmax);
            std::vector<Vector3d> oc_v;
            for (int k = 0; k < 35; ++k) oc_v.push_back(Vector3d(oc[3 * k], oc[3 * k + 1], oc[3 * k + 2]));
            pcs_oracle.push_back(oc_v);
        }
        CHECK(ok_c && corners.size() == 35, "frame %d: est_board_corner %d, %zu corners", f, (int)ok_c, corners.size());
        if (!ok_c || corners.size() != 35) continue;
        std::vector<Vector2d> im;
        std::vector<int> hit(35, 0);
        double worst = 0;
        for (auto& c : corners) {
            double best = 1e9, bo[3] = {0, 0, 0}; int bi = -1;
            for (int k = 0; k < 35; ++k) {
                double o[3];
                board_to_cam(-0.3 + 0.1 * (k % 7), -0.2 + 0.1 * (k / 7), 0.0, o);
                const double d = std::sqrt((c.x() - o[0]) * (c.x() - o[0]) + (c.y() - o[1]) * (c.y() - o[1]) + (c.z() - o[2]) * (c.z() - o[2]));
                if (d < best) { best = d; bi = k; bo[0] = o[0]; bo[1] = o[1]; bo[2] = o[2]; }
            }
            worst = std::max(worst, best);
            if (bi >= 0) hit[(size_t)bi]++;
            // what the camera sees of that corner (the image-side detector is out of scope: OpenCV in the reference)
            double P[3];
            for (int r = 0; r < 3; ++r) P[r] = Rcl[3 * r] * bo[0] + Rcl[3 * r + 1] * bo[1] + Rcl[3 * r + 2] * bo[2] + tcl[r];
            im.push_back(Vector2d(P[0] / P[2] + 1e-4 * N01(rng), P[1] / P[2] + 1e-4 * N01(rng)));
        }
        bool once = true;
        for (int h : hit) once = once && h == 1;
        CHECK(once && worst < 6e-3, "frame %d: corner error %g m, each once %d", f, worst, (int)once);
        pcs.push_back(corners);
        ims.push_back(im);
    }
    // ---- stage 2, extrinsic ----
    CHECK((int)pcs.size() == F, "%zu of %d frames produced corners", pcs.size(), F);
    if (pcs.size() >= 2) {
        double Tcl[7] = {0, 0, 0, 1, 0, 0, 0};
        CHECK(pnp_init(ims, pcs, Tcl), "pnp_init");
        pose_reprj_opt(ims, pcs, Tcl, false);
        const double x = Tcl[0], y = Tcl[1], z = Tcl[2], w = Tcl[3], nq = std::sqrt(x * x + y * y + z * z + w * w);
        const double q[4] = {x / nq, y / nq, z / nq, w / nq};
        const double R[9] = {1 - 2 * (q[1] * q[1] + q[2] * q[2]), 2 * (q[0] * q[1] - q[2] * q[3]), 2 * (q[0] * q[2] + q[1] * q[3]),
                             2 * (q[0] * q[1] + q[2] * q[3]), 1 - 2 * (q[0] * q[0] + q[2] * q[2]), 2 * (q[1] * q[2] - q[0] * q[3]),
                             2 * (q[0] * q[2] - q[1] * q[3]), 2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[0] * q[0] + q[1] * q[1])};
        double tr = 0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tr += R[3 * i + j] * Rcl[3 * i + j];
        const double ang = std::acos(std::min(1.0, std::max(-1.0, (tr - 1) / 2))) * 180.0 / M_PI;
        const double dt = std::sqrt((Tcl[4] - tcl[0]) * (Tcl[4] - tcl[0]) + (Tcl[5] - tcl[1]) * (Tcl[5] - tcl[1]) + (Tcl[6] - tcl[2]) * (Tcl[6] - tcl[2]));
        std::printf("extrinsic: rotation error %.4f deg, translation error %.4f m\n", ang, dt);
        CHECK(ang < 0.2 && dt < 0.02, "extrinsic error %g deg, %g m", ang, dt);
        // the same solve from the ORACLE's corners: the two extrinsics must agree to the north star's bar
        if (pcs_oracle.size() == pcs.size()) {
            double To[7] = {0, 0, 0, 1, 0, 0, 0};
            CHECK(pnp_init(ims, pcs_oracle, To), "pnp_init (oracle corners)");
            pose_reprj_opt(ims, pcs_oracle, To, false);
            const double no = std::sqrt(To[0] * To[0] + To[1] * To[1] + To[2] * To[2] + To[3] * To[3]);
            double dq = 0, dtt = 0;
            const double sgn = (To[3] * Tcl[3] < 0) ? -1.0 : 1.0;
            for (int k = 0; k < 4; ++k) dq = std::max(dq, std::fabs(To[k] / no * sgn - Tcl[k] / nq));
            for (int k = 4; k < 7; ++k) dtt = std::max(dtt, std::fabs(To[k] - Tcl[k]));
            std::printf("extrinsic from the oracle's corners: |dq| %.2e (about %.2e rad), |dt| %.2e m\n", dq, 2 * dq, dtt);
            // the north star's bar: 1e-4 rad, 0.1 mm
            CHECK(2 * dq < 1e-4 && dtt < 1e-4, "extrinsic differs from the oracle-fed solve: %g rad, %g m", 2 * dq, dtt);
        }
    }
    if (g_fail) { std::printf("%d check(s) failed\n", g_fail); return 1; }
    std::printf("pipeline checks passed\n");
    return 0;
}
