// test_ops.cpp — exercises the reference-shaped C++ operator surface (include/rclc_ops.hpp) end to end on the GPU.
// Built and run by tests/test_cpp_ops.py. Exit codes: 0 all checks passed, 77 no CUDA device (nothing computed: the
// product has no CPU fallback), anything else = failure.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "rclc_ops.hpp"

using namespace rclc_b200;

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++fails; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

// board-only cloud in the (approximate) board frame: stored p satisfy R(yaw) (p + (dx, dy)) = truth  (board_fitting_cost.h:132)
static PointCloud<PoinT>::Ptr board_cloud(int n, unsigned seed, double dx, double dy, double yaw_deg)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> ux(-0.4 + 1e-6, 0.4 - 1e-6), uy(-0.3 + 1e-6, 0.3 - 1e-6);
    std::normal_distribution<double> nb(15.0, 4.0), nw(100.0, 12.0), nz(0.0, 0.002);
    auto cloud = std::make_shared<PointCloud<PoinT>>();
    cloud->points.resize((size_t)n);
    const double th = yaw_deg * M_PI / 180.0, c = std::cos(th), s = std::sin(th);
    for (int i = 0; i < n; ++i) {
        const double u = ux(rng), v = uy(rng);
        const int ci = (int)std::floor((u + 0.4) / 0.1), ri = (int)std::floor((v + 0.3) / 0.1);
        const bool black = ((ci + ri) % 2) == 0;
        double I = black ? nb(rng) : nw(rng);
        I = std::min(149.0, std::max(4.0, I));
        PoinT& p = cloud->points[(size_t)i];
        p.x = (float)(c * u + s * v - dx);
        p.y = (float)(-s * u + c * v - dy);
        p.z = (float)nz(rng);
        p.intensity = (float)I;
    }
    return cloud;
}

int main()
{
    if (!thread_ctx()) { std::printf("no CUDA device: nothing computed (there is no CPU fallback)\n"); return 77; }

    // ---- GMMFit::fit_1d (BoardSegmentation.cpp:134-137 initial guesses) ----
    auto cloud = board_cloud(60000, 7, 0.006, -0.004, 0.4);
    {
        VectorXd I((long)cloud->size());
        for (size_t i = 0; i < cloud->size(); ++i) I((long)i) = cloud->points[i].intensity;
        std::vector<double> mu = {15, 100}, sigma = {15, 15};
        CHECK(GMMFit::fit_1d(I, 2, mu, sigma, false), "fit_1d returned false");
        CHECK(std::fabs(mu[0] - 15) < 1.0 && std::fabs(mu[1] - 100) < 2.0, "gmm means %g %g", mu[0], mu[1]);
        CHECK(sigma[0] > 2 && sigma[0] < 6 && sigma[1] > 8 && sigma[1] < 16, "gmm sigmas %g %g", sigma[0], sigma[1]);
    }
    // ---- BOARD_FITTING_COST: 82 functors share one device cloud; Evaluate at the true misalignment has small residuals ----
    {
        auto dev = std::make_shared<BoardCloudOnDevice>(cloud);
        const rclc_config& cfg = CfgParam();
        const int nt = (cfg.board_width - 1) * cfg.board_height + (cfg.board_height - 1) * cfg.board_width;
        std::vector<double> txy((size_t)nt * 2); std::vector<int32_t> isr((size_t)nt); int32_t n = 0;
        rclc_generate_targets(&cfg, txy.data(), isr.data(), &n);
        CHECK(n == 82, "targets %d", n);
        double at_truth[3] = {0.006, -0.004, 0.4}, at_zero[3] = {0, 0, 0};
        const double* params[1];
        double cost_truth = 0, cost_zero = 0;
        for (int t = 0; t < nt; ++t) {
            std::unique_ptr<BOARD_FITTING_COST> f(BOARD_FITTING_COST::Create(Vector3d(txy[2 * t], txy[2 * t + 1], 0), isr[t] != 0, dev));
            double r = 0, j[3] = {0, 0, 0}; double* jj[1] = {j};
            params[0] = at_truth; CHECK(f->Evaluate(params, &r, jj), "Evaluate"); cost_truth += r * r;
            CHECK(std::isfinite(r) && std::isfinite(j[0]) && std::isfinite(j[2]), "non-finite residual at target %d", t);
            params[0] = at_zero; f->Evaluate(params, &r, nullptr); cost_zero += r * r;
        }
        CHECK(cost_truth < cost_zero, "cost at truth %g should be below cost at zero %g", cost_truth, cost_zero);
    }
    // ---- opt_grid_fitting: recovers the board: corners within 5 mm of the truth lattice ----
    {
        Matrix4d T_bl;                      // identity: the cloud is already in the board frame
        std::vector<Vector3d> corners;
        rclc_gridfit_report rep;
        opt_grid_fitting(cloud, T_bl, corners, nullptr, &rep);
        CHECK(corners.size() == 35, "corners %zu", corners.size());
        // est corners live in the laser frame = stored frame here; truth corner (u, v) is stored at R^T (u, v) - t
        const double th = 0.4 * M_PI / 180.0, c = std::cos(th), s = std::sin(th);
        double worst = 0;
        int k = 0;
        for (int row = 0; row < 5 && corners.size() == 35; ++row)
            for (int col = 0; col < 7; ++col, ++k) {
                const double u = 0.3 - col * 0.1, v = 0.2 - row * 0.1;          // utils.h:47-66
                const double px = c * u + s * v - 0.006, py = -s * u + c * v + 0.004;
                worst = std::max(worst, std::hypot(corners[(size_t)k].x() - px, corners[(size_t)k].y() - py));
            }
        CHECK(worst < 5e-3, "corner error %g m", worst);   // accuracy of the reference's own heuristic-Jacobian fit at 60k points
        CHECK(rep.outer_iterations >= 1 && rep.pose_evals > 0 && rep.last_final_cost <= rep.first_initial_cost, "report");
    }
    // ---- pc_plane_prj: points end up on the plane ----
    {
        auto pc = std::make_shared<PointCloud<PoinT>>();
        std::mt19937 rng(3); std::uniform_real_distribution<double> u(-0.3, 0.3), z(3.8, 4.2);
        for (int i = 0; i < 5000; ++i) { PoinT p; p.z = (float)z(rng); p.x = (float)(u(rng) * p.z); p.y = (float)(u(rng) * p.z); p.intensity = 50; pc->points.push_back(p); }
        Vector4d plane(0.1, -0.05, 0.99, -3.96);
        pc_plane_prj(pc, plane);
        double worst = 0;
        for (auto& p : pc->points) worst = std::max(worst, std::fabs(0.1 * p.x - 0.05 * p.y + 0.99 * p.z - 3.96));
        CHECK(worst < 1e-5, "distance to plane %g", worst);
    }
    // ---- pc_plane_fitting: the board plane from the bright points only (the dark ones are displaced on purpose) ----
    {
        auto pc = std::make_shared<PointCloud<PoinT>>();
        std::mt19937 rng(9); std::uniform_real_distribution<double> u(-0.4, 0.4); std::normal_distribution<double> nz(0, 0.003);
        for (int i = 0; i < 30000; ++i) {
            PoinT p; p.x = (float)u(rng); p.y = (float)u(rng);
            const bool bright = (i % 2) == 0;
            p.z = (float)(4.0 - 0.1 * p.x + 0.2 * p.y + nz(rng) + (bright ? 0.0 : 0.05));
            p.intensity = bright ? 120.f : 20.f;
            pc->points.push_back(p);
        }
        Vector4d plane;
        CHECK(pc_plane_fitting(pc, plane) == 0, "return value");
        const double nrm = std::sqrt(plane.x() * plane.x() + plane.y() * plane.y() + plane.z() * plane.z());
        const double s = plane.z() < 0 ? -1.0 : 1.0, ref = std::sqrt(0.01 + 0.04 + 1.0);
        CHECK(std::fabs(s * plane.x() / nrm - 0.1 / ref) < 1e-3 && std::fabs(s * plane.y() / nrm + 0.2 / ref) < 1e-3 &&
                  std::fabs(s * plane.w() / nrm + 4.0 / ref) < 1e-3,
              "plane %g %g %g %g", plane.x() / nrm, plane.y() / nrm, plane.z() / nrm, plane.w() / nrm);
    }
    // ---- UniformSampling / EuclideanClusterExtraction / EulerCluster / eular_cluster_iter ----
    {
        // a board whose black squares read darker towards their middle (mixed pixels along the edges), plus far clutter
        std::mt19937 rng(11);
        std::uniform_real_distribution<double> ux(-0.4 + 1e-6, 0.4 - 1e-6), uy(-0.3 + 1e-6, 0.3 - 1e-6), far(-2.0, 2.0);
        std::normal_distribution<double> nI(0.0, 2.0), nz(0.0, 0.001);
        auto cloud = std::make_shared<PointCloud<PoinT>>();
        for (int i = 0; i < 150000; ++i) {
            PoinT p;
            const double u = ux(rng), v = uy(rng);
            const double fu = (u + 0.4) / 0.1, fv = (v + 0.3) / 0.1;
            const int ci = (int)std::floor(fu), ri = (int)std::floor(fv);
            const double eu = std::min(fu - ci, ci + 1 - fu) * 0.1, ev = std::min(fv - ri, ri + 1 - fv) * 0.1, e = std::min(eu, ev);
            const bool black = ((ci + ri) % 2) == 0;
            p.x = (float)u; p.y = (float)v; p.z = (float)(4.004 + nz(rng));      // mid-voxel for the 8 mm grid below
            p.intensity = (float)std::max(4.0, (black ? 10.0 + 60.0 * std::max(0.0, 1.0 - e / 0.012) : 110.0) + nI(rng));
            cloud->points.push_back(p);
        }
        const size_t n_board = cloud->size();
        for (int i = 0; i < 300; ++i) { PoinT p; p.x = (float)(3.0 + far(rng)); p.y = (float)far(rng); p.z = (float)(8.0 + far(rng)); p.intensity = 50.f; cloud->points.push_back(p); }

        auto thin = std::make_shared<PointCloud<PoinT>>();
        UniformSampling US; US.setInputCloud(cloud); US.setRadiusSearch(0.008); US.filter(*thin);
        CHECK(thin->size() > 5000 && thin->size() < 9000, "uniform sampling kept %zu", thin->size());        // ~ 0.48 m^2 / (8 mm)^2 = 7500 voxels
        PointCloud<PoinT>::Ptr board;
        EulerCluster(thin, board);
        CHECK(board && board->size() > 5000 && board->size() + 250 < thin->size() + 1 && board->size() <= n_board, "largest cluster %zu of %zu", board ? board->size() : 0, thin->size());
        bool all_on_board = true;
        if (board) for (auto& p : board->points) all_on_board = all_on_board && p.z < 5.f;
        CHECK(all_on_board, "largest cluster is the board");

        std::vector<Vector3d> centroids;
        eular_cluster_iter(board, 50.0, centroids);
        CHECK(centroids.size() == 24, "black block centroids: %zu", centroids.size());
        double worst = 0;
        for (auto& c : centroids) {
            // nearest black square centre
            const double fu = (c.x() + 0.4) / 0.1, fv = (c.y() + 0.3) / 0.1;
            const int ci = (int)std::floor(fu), ri = (int)std::floor(fv);
            CHECK(((ci + ri) % 2) == 0, "centroid (%g, %g) is on a white square", c.x(), c.y());
            worst = std::max(worst, std::hypot(c.x() - (-0.4 + 0.1 * ci + 0.05), c.y() - (-0.3 + 0.1 * ri + 0.05)));
        }
        CHECK(worst < 3e-3, "block centroid error %g m", worst);
    }
    // ---- est_board_corner: the whole per-frame chain (plane fit, projection, GMM, block clustering, board frame, grid fit) ----
    {
        // board in the laser frame (x right, y down, z forward, i.e. after T_base), top-right square black as the reference requires
        std::mt19937 rng(21);
        std::uniform_real_distribution<double> ux(-0.4 + 1e-6, 0.4 - 1e-6), uy(-0.3 + 1e-6, 0.3 - 1e-6);
        std::normal_distribution<double> nI(0.0, 2.0), nz(0.0, 0.002);
        const double ax = 0.06, ay = -0.09, az = 0.05;      // small tilt
        const double cxr = std::cos(ax), sxr = std::sin(ax), cyr = std::cos(ay), syr = std::sin(ay), czr = std::cos(az), szr = std::sin(az);
        const double R[9] = {czr * cyr, czr * syr * sxr - szr * cxr, czr * syr * cxr + szr * sxr, szr * cyr, szr * syr * sxr + czr * cxr, szr * syr * cxr - czr * sxr,
                             -syr, cyr * sxr, cyr * cxr};
        const double t[3] = {0.1, -0.05, 4.0};
        auto to_laser = [&](double u, double v, double w, double* o) { for (int r = 0; r < 3; ++r) o[r] = R[3 * r] * u + R[3 * r + 1] * v + R[3 * r + 2] * w + t[r]; };
        auto cloud = std::make_shared<PointCloud<PoinT>>();
        for (int i = 0; i < 120000; ++i) {
            const double u = ux(rng), v = uy(rng);
            const double fu = (u + 0.4) / 0.1, fv = (v + 0.3) / 0.1;
            const int ci = (int)std::floor(fu), ri = (int)std::floor(fv);
            const double e = std::min(std::min(fu - ci, ci + 1 - fu), std::min(fv - ri, ri + 1 - fv)) * 0.1;
            const bool black = ((ci + ri) % 2) == 1;         // (7, 0) = max x, min y = top right: black
            double o[3];
            to_laser(u, v, nz(rng), o);
            PoinT p; p.x = (float)o[0]; p.y = (float)o[1]; p.z = (float)o[2];
            p.intensity = (float)std::max(4.0, (black ? 10.0 + 60.0 * std::max(0.0, 1.0 - e / 0.012) : 110.0) + nI(rng));
            cloud->points.push_back(p);
        }
        std::vector<Vector3d> corners;
        rclc_gridfit_report rep;
        const bool done = est_board_corner(cloud, corners, nullptr, &rep);
        CHECK(done && corners.size() == 35, "est_board_corner: %d, %zu corners", (int)done, corners.size());
        // every interior corner of the board is found once, within 5 mm
        std::vector<int> hit(35, 0);
        double worst = 0;
        for (auto& c : corners) {
            double best = 1e9; int bi = -1;
            for (int k = 0; k < 35; ++k) {
                double o[3];
                to_laser(-0.3 + 0.1 * (k % 7), -0.2 + 0.1 * (k / 7), 0.0, o);
                const double d = std::sqrt((c.x() - o[0]) * (c.x() - o[0]) + (c.y() - o[1]) * (c.y() - o[1]) + (c.z() - o[2]) * (c.z() - o[2]));
                if (d < best) { best = d; bi = k; }
            }
            worst = std::max(worst, best);
            if (bi >= 0) hit[(size_t)bi]++;
        }
        bool once = true;
        for (int h : hit) once = once && h == 1;
        CHECK(once && worst < 5e-3, "corner error %g m, each corner once: %d", worst, (int)once);
    }
    // ---- PlaneSegmentation (pcl::SACSegmentation drop-in): plane + 30 % outliers ----
    {
        auto pc = std::make_shared<PointCloud<PoinT>>();
        std::mt19937 rng(5); std::uniform_real_distribution<double> u(-1, 1), o(-1, 1); std::normal_distribution<double> nz(0, 0.01);
        for (int i = 0; i < 20000; ++i) {
            PoinT p; p.y = (float)u(rng); p.z = (float)u(rng);
            p.x = (i % 10 < 7) ? (float)(4.0 + 0.2 * p.y - 0.1 * p.z + nz(rng)) : (float)(4.0 + 2.0 * o(rng));
            pc->points.push_back(p);
        }
        PlaneSegmentation sac; sac.setInputCloud(pc); sac.setMaxIterations(800); sac.setDistanceThreshold(0.08);
        PointIndices inl; ModelCoefficients coef;
        sac.segment(inl, coef);
        CHECK(coef.values.size() == 4, "coefficients");
        CHECK(inl.indices.size() > 14000 && inl.indices.size() < 16500, "inliers %zu", inl.indices.size());
        if (coef.values.size() == 4) {
            const double nrm = std::sqrt(1 + 0.04 + 0.01), s = coef.values[0] < 0 ? -1 : 1;
            CHECK(std::fabs(s * coef.values[0] - 1 / nrm) < 5e-3 && std::fabs(s * coef.values[1] + 0.2 / nrm) < 5e-3, "normal %g %g %g", coef.values[0], coef.values[1], coef.values[2]);
        }
        CHECK(sac.iterations() >= 1 && sac.iterations() <= 800, "iterations %d", sac.iterations());
    }
    // ---- pnp_init + pose_reprj_opt: the extrinsic comes back from the correspondences alone ----
    {
        std::mt19937 rng(11); std::uniform_real_distribution<double> u(-0.4, 0.4), d(3.5, 5.5); std::normal_distribution<double> px(0, 1e-4);
        const double q[4] = {0.02, -0.01, 0.015, 0.9996}; const double t[3] = {0.05, -0.03, 0.02};
        auto rot = [&](const double* qq, const double* p, double* out) {
            const double x = qq[0], y = qq[1], z = qq[2], w = qq[3];
            const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                                 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
            for (int r = 0; r < 3; ++r) out[r] = R[3 * r] * p[0] + R[3 * r + 1] * p[1] + R[3 * r + 2] * p[2];
        };
        std::vector<std::vector<Vector3d>> pcs; std::vector<std::vector<Vector2d>> ims;
        for (int f = 0; f < 12; ++f) {
            std::vector<Vector3d> pc; std::vector<Vector2d> im;
            const double zc = d(rng), xc = u(rng), yc = u(rng);
            for (int k = 0; k < 35; ++k) {
                const double p[3] = {xc + 0.1 * (k % 7), yc + 0.1 * (k / 7), zc + 0.02 * (k % 5)};
                double P[3]; rot(q, p, P); for (int r = 0; r < 3; ++r) P[r] += t[r];
                pc.push_back(Vector3d(p[0], p[1], p[2])); im.push_back(Vector2d(P[0] / P[2] + px(rng), P[1] / P[2] + px(rng)));
            }
            pcs.push_back(pc); ims.push_back(im);
        }
        // start from the PnP pose, as apps/Calibrate.cpp:41-63 does
        double Tpnp[7] = {0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0};
        CHECK(pnp_init(ims, pcs, Tpnp), "pnp_init");
        double dq0 = 0; for (int k = 0; k < 4; ++k) dq0 = std::max(dq0, std::fabs(Tpnp[k] - q[k] / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])));
        CHECK(dq0 < 2e-3 && std::fabs(Tpnp[4] - t[0]) < 5e-3 && std::fabs(Tpnp[5] - t[1]) < 5e-3 && std::fabs(Tpnp[6] - t[2]) < 2e-2, "PnP start %g %g %g %g | %g %g %g", Tpnp[0],
              Tpnp[1], Tpnp[2], Tpnp[3], Tpnp[4], Tpnp[5], Tpnp[6]);
        double Tcl[7];
        for (int k = 0; k < 7; ++k) Tcl[k] = Tpnp[k];
        pose_reprj_opt(ims, pcs, Tcl, false);
        const double nq = std::sqrt(Tcl[0] * Tcl[0] + Tcl[1] * Tcl[1] + Tcl[2] * Tcl[2] + Tcl[3] * Tcl[3]);
        double dq = 0; for (int k = 0; k < 4; ++k) dq = std::max(dq, std::fabs(Tcl[k] / nq - q[k] / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])));
        CHECK(dq < 2e-3 && std::fabs(Tcl[4] - t[0]) < 5e-3 && std::fabs(Tcl[5] - t[1]) < 5e-3 && std::fabs(Tcl[6] - t[2]) < 2e-2, "Tcl %g %g %g %g | %g %g %g", Tcl[0], Tcl[1],
              Tcl[2], Tcl[3], Tcl[4], Tcl[5], Tcl[6]);
    }
    // ---- PinholeCam + colorize_cloud agree point by point ----
    {
        rclc_config& cfg = CfgParam();
        PinholeCam cam(cfg.cam_width, cfg.cam_height, cfg.fx, cfg.fy, cfg.cx, cfg.cy);
        std::vector<uint8_t> img((size_t)cfg.cam_width * cfg.cam_height * 3);
        for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);
        auto pc = std::make_shared<PointCloud<DisplayType>>();
        std::mt19937 rng(9); std::uniform_real_distribution<double> z(1, 50), a(-1.1, 1.1);
        for (int i = 0; i < 20000; ++i) { DisplayType p; p.z = (float)z(rng); p.x = (float)(a(rng) * p.z); p.y = (float)(a(rng) * 0.65 * p.z); pc->points.push_back(p); }
        auto before = pc->points;
        Matrix4f T; T(0, 3) = 0.05f; T(1, 3) = -0.02f; T(0, 1) = 0.01f; T(1, 0) = -0.01f;
        std::vector<uint8_t> fov;
        colorize_cloud(pc, T, img.data(), &fov);
        int bad = 0, inside = 0;
        for (size_t i = 0; i < pc->size() && fov.size() == pc->size(); ++i) {
            const DisplayType& p = pc->points[i];
            Vector2d pix = cam.Pt2img(Vector3d(p.x, p.y, p.z));
            const bool in = cam.is_in_FoV(pix);
            if (in != (fov[i] != 0)) { ++bad; continue; }
            if (!in) continue;
            ++inside;
            const size_t o = ((size_t)pix.y() * cfg.cam_width + (size_t)pix.x()) * 3;
            if (p.r != img[o + 2] || p.g != img[o + 1] || p.b != img[o]) ++bad;
        }
        CHECK(fov.size() == pc->size() && bad == 0 && inside > 1000, "colourise mismatches %d (inside %d)", bad, inside);
        (void)before;
    }
    std::printf(fails ? "%d check(s) failed\n" : "all operator checks passed%.0d\n", fails);
    return fails ? 1 : 0;
}
