// rclc_pipeline.cpp — CPU ORACLE (test infrastructure, not the product): the reference's host-side chains restated over the
// oracle's own primitives, so that the product's reference-shaped operators (include/rclc_ops.hpp) can be compared stage by stage.
//
//   orc_segment_board       apps/BoardSegmentation.cpp:45-165   (PassThrough, UniformSampling, the RANSAC peeling loop,
//                                                                 EulerCluster :59-79, cut_roi :81-101, GMM split)
//   orc_eular_cluster_iter  include/pc_process.h:104-211
//   orc_est_board_corner    include/pc_process.h:583-646
//
// Parity unpinned: PCL / Ceres are absent here and the reference ships no tests for this path; third-party
// semantics follow SURVEY.md Appendix A / B, the same definitions the primitives in rclc_oracle.cpp use.
#include "rclc_oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

typedef std::array<float, 4> P4;        // x, y, z, intensity
typedef std::vector<P4> Cloud;

Cloud pass_through(const Cloud& in, int field, float lo, float hi)       // pcl::PassThrough, inclusive, finite points only
{
    Cloud out;
    for (const P4& p : in) {
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        const float v = p[(size_t)field];
        if (std::isfinite(v) && v >= lo && v <= hi) out.push_back(p);
    }
    return out;
}

Cloud uniform_sample(const Cloud& in, double leaf)
{
    Cloud out;
    if (in.empty()) return out;
    std::vector<int32_t> idx(in.size());
    int64_t n = 0;
    orc_uniform_sample(in.data(), 16, (int64_t)in.size(), leaf, idx.data(), &n);
    for (int64_t k = 0; k < n; ++k) out.push_back(in[(size_t)idx[(size_t)k]]);
    return out;
}

std::vector<std::vector<int>> clusters(const Cloud& in, double tol, int64_t mn, int64_t mx)
{
    std::vector<std::vector<int>> out;
    if (in.empty()) return out;
    std::vector<int32_t> lab(in.size()), sizes(in.size());
    int32_t nc = 0;
    orc_euclidean_cluster(in.data(), 16, (int64_t)in.size(), tol, mn, mx, lab.data(), &nc, sizes.data());
    out.resize((size_t)nc);
    for (size_t i = 0; i < in.size(); ++i) if (lab[i] >= 0) out[(size_t)lab[i]].push_back((int)i);
    return out;
}

Cloud euler_cluster(const Cloud& in)                                      // pc_process.h:59-79
{
    Cloud out;
    const auto cl = clusters(in, 0.03, 500, 2500000);
    if (cl.empty()) return out;
    for (int i : cl[0]) out.push_back(in[(size_t)i]);
    return out;
}

// pcl::SACSegmentation<>::segment as configured at BoardSegmentation.cpp:64-72
bool sac_segment(const Cloud& in, int max_iterations, double thr, std::vector<uint8_t>& mask, float plane[4])
{
    const int64_t n = (int64_t)in.size();
    if (n < 3) return false;
    const int H = max_iterations + 1;
    std::vector<int32_t> tri((size_t)H * 3), valid((size_t)H), counts((size_t)H);
    std::vector<float> planes((size_t)H * 4);
    orc_ransac_draw_triplets(n, H, tri.data());
    orc_ransac_score(in.data(), 16, n, tri.data(), H, thr, planes.data(), valid.data(), counts.data());
    int32_t best = -1, iters = 0;
    orc_ransac_replay(counts.data(), valid.data(), H, n, max_iterations, 0.99, &best, &iters);
    if (best < 0) return false;
    mask.assign((size_t)n, 0);
    orc_plane_select(in.data(), 16, n, &planes[(size_t)best * 4], thr, mask.data());
    float centroid[3];
    orc_plane_refine(in.data(), 16, n, mask.data(), &planes[(size_t)best * 4], plane, centroid);
    orc_plane_select(in.data(), 16, n, plane, thr, mask.data());            // re-selection with the optimised coefficients
    return true;
}

void gmm_threshold(const rclc_config* cfg, const Cloud& c, double k_sigma, double& thd)
{
    std::vector<float> I(c.size());
    for (size_t i = 0; i < c.size(); ++i) I[i] = c[i][3];
    double mu[2] = {15, 100}, sigma[2] = {15, 15}, pri[8];
    orc_gmm_fit_1d(I.data(), 4, (int64_t)I.size(), 2, cfg->gmm_iters, mu, sigma, pri);
    thd = mu[0] + k_sigma * sigma[0];
}

Cloud load(const void* pts, int stride, int ioff, int64_t n)
{
    Cloud c((size_t)n);
    for (int64_t k = 0; k < n; ++k) {
        const char* r = static_cast<const char*>(pts) + (size_t)k * (size_t)stride;
        std::memcpy(c[(size_t)k].data(), r, 12);
        std::memcpy(&c[(size_t)k][3], r + ioff, 4);
    }
    return c;
}

int eular_cluster_iter(const rclc_config* cfg, const Cloud& raw, double reflactive_thd, std::vector<std::array<double, 3>>& block_centroid, int max_rounds)
{
    Cloud blocks = uniform_sample(raw, cfg->block_us_radius);
    std::vector<size_t> block_size;                     // points per accepted block (the merged clouds themselves are not needed)
    std::vector<int> block_num_pts;
    const int block_num = cfg->board_width * cfg->board_height / 2;
    const double diag = cfg->gride_side_length * std::sqrt(2.0f);
    double curr = reflactive_thd / cfg->reflactivity_dec_step;
    for (int round = 0; round < max_rounds; ++round) {
        curr *= cfg->reflactivity_dec_step;
        blocks = pass_through(blocks, 3, 3.0f, (float)curr);
        if (blocks.empty()) return 1;
        const auto cl = clusters(blocks, cfg->black_pt_cluster_dst_thd, 3, 25000);
        if ((int)cl.size() < block_num) continue;
        for (int s = 0; s < (int)cl.size(); ++s) {
            float cx = 0.f, cy = 0.f, cz = 0.f;          // pcl::compute3DCentroid, float accumulation
            for (int i : cl[(size_t)s]) { cx += blocks[(size_t)i][0]; cy += blocks[(size_t)i][1]; cz += blocks[(size_t)i][2]; }
            const float inv = 1.0f / (float)cl[(size_t)s].size();
            const std::array<double, 3> c{(double)(cx * inv), (double)(cy * inv), (double)(cz * inv)};
            if (s < block_num) {
                if (s > 0 && (double(cl[(size_t)s].size()) / double(block_size[(size_t)s - 1])) < 0.65) break;
                block_size.push_back(cl[(size_t)s].size());
                block_centroid.push_back(c);
                block_num_pts.push_back((int)cl[(size_t)s].size());
            } else {
                int nearest = -1;
                double nearest_dst = 10e3;
                for (int i = 0; i < (int)block_centroid.size(); ++i) {
                    const double dx = block_centroid[(size_t)i][0] - c[0], dy = block_centroid[(size_t)i][1] - c[1], dz = block_centroid[(size_t)i][2] - c[2];
                    const float dst = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
                    if (dst < nearest_dst) { nearest_dst = dst; nearest = i; }
                }
                if (nearest >= 0 && nearest_dst < 0.5 * diag) {
                    auto& old = block_centroid[(size_t)nearest];
                    const double a = block_num_pts[(size_t)nearest], b = (double)cl[(size_t)s].size();
                    for (int d = 0; d < 3; ++d) old[(size_t)d] = (old[(size_t)d] * a + c[(size_t)d] * b) / (a + b);
                    block_num_pts[(size_t)nearest] += (int)cl[(size_t)s].size();
                    block_size[(size_t)nearest] += cl[(size_t)s].size();
                }
            }
        }
        if ((int)block_centroid.size() < block_num) { block_size.clear(); block_centroid.clear(); block_num_pts.clear(); continue; }
        return 0;
    }
    return 1;
}

} // namespace

extern "C" {

int orc_eular_cluster_iter(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double reflactive_thd, double* centroids_out,
                           int32_t* n_out)
{
    std::vector<std::array<double, 3>> c;
    const int rc = eular_cluster_iter(cfg, load(pts, stride, ioff, n), reflactive_thd, c, 200);
    *n_out = (int32_t)c.size();
    for (size_t k = 0; k < c.size(); ++k) for (int d = 0; d < 3; ++d) centroids_out[3 * k + (size_t)d] = c[k][(size_t)d];
    return rc;
}

// pts_io [n] float4 {x, y, z, I} (stride 16): projected onto the fitted plane and moved into the board frame, like the reference's
// chess_board_processe. corners_out (W-1)*(H-1)*3.
int orc_est_board_corner(const rclc_config* cfg, void* pts_io, int64_t n, double* corners_out, orc_gridfit_report* rep)
{
    if (cfg->board_order == 1) {
        const float T[16] = {-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        orc_transform_cloud(pts_io, 16, n, T);
    }
    double plane[4];
    int32_t its = 0;
    int64_t used = 0;
    if (orc_plane_fit(cfg, pts_io, 16, 12, n, plane, &its, &used) != 0) return 1;
    orc_plane_project(pts_io, 16, n, plane);
    Cloud c = load(pts_io, 16, 12, n);
    double thd;
    gmm_threshold(cfg, c, 2.0, thd);
    std::vector<std::array<double, 3>> cen;
    if (eular_cluster_iter(cfg, c, thd, cen, 200) != 0) return 2;
    std::vector<double> flat;
    for (auto& p : cen) { flat.push_back(p[0]); flat.push_back(p[1]); flat.push_back(p[2]); }
    double Tbl[16];
    if (orc_calc_laser_coor(cfg, flat.data(), (int32_t)cen.size(), plane, Tbl) != 0) return 3;
    float Tf[16];
    for (int k = 0; k < 16; ++k) Tf[k] = (float)Tbl[k];
    orc_transform_cloud(pts_io, 16, n, Tf);
    orc_gridfit_report r;
    if (orc_gridfit_solve(cfg, pts_io, 16, 12, n, Tbl, corners_out, &r) != 0) return 4;
    if (rep) *rep = r;
    if (cfg->board_order == 1) {
        const int per_row = cfg->board_width - 1, rows = cfg->board_height - 1;
        std::vector<double> tmp((size_t)per_row * rows * 3);
        for (int row = 0; row < rows; ++row)
            for (int col = 0; col < per_row; ++col) {
                const double* p = corners_out + 3 * (row * per_row + col);
                double* q = tmp.data() + 3 * (row * per_row + per_row - col - 1);
                q[0] = -p[0]; q[1] = p[1]; q[2] = p[2];
            }
        std::memcpy(corners_out, tmp.data(), tmp.size() * 8);
    }
    return 0;
}

// out [capacity n] float4: white cluster followed by black cluster (BoardSegmentation.cpp:160). Returns 0, or 1 when no plane qualifies.
int orc_segment_board(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double x_min, double x_max, void* out, int64_t* n_out)
{
    *n_out = 0;
    Cloud raw = pass_through(load(pts, stride, ioff, n), 0, (float)x_min, (float)x_max);
    Cloud vox = uniform_sample(raw, 0.008f);
    const size_t pt_size = vox.size();
    Cloud target;
    double closest = 1000;
    while (vox.size() > pt_size * 0.05) {
        std::vector<uint8_t> mask;
        float plane[4];
        if (!sac_segment(vox, 800, 0.08, mask, plane)) break;
        Cloud ext, rest;
        for (size_t i = 0; i < vox.size(); ++i) (mask[i] ? ext : rest).push_back(vox[i]);
        if (ext.empty()) break;
        vox.swap(rest);
        const double a = plane[0], b = plane[1], c = plane[2];
        const double nx = a / std::sqrt(a * a + b * b + c * c);
        if (ext.size() < pt_size * 0.02 || std::fabs(nx) <= cfg->angle_thd) continue;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (const P4& p : ext) { sx += p[0]; sy += p[1]; sz += p[2]; }
        const float inv = 1.0f / (float)ext.size();
        const double dist = std::sqrt((double)(sx * inv) * (sx * inv) + (double)(sy * inv) * (sy * inv) + (double)(sz * inv) * (sz * inv));
        if ((closest - dist) > 0.3) { closest = dist; target = ext; }
    }
    if (target.empty()) return 1;
    target = euler_cluster(target);
    if (target.empty()) return 1;
    float mn[3] = {target[0][0], target[0][1], target[0][2]}, mx[3] = {target[0][0], target[0][1], target[0][2]};
    for (const P4& p : target) for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], p[(size_t)d]); mx[d] = std::max(mx[d], p[(size_t)d]); }
    for (int d = 0; d < 3; ++d) raw = pass_through(raw, d, mn[d], mx[d]);        // cut_roi
    double thd;
    gmm_threshold(cfg, target, 1.7, thd);
    const Cloud black = euler_cluster(pass_through(raw, 3, (float)cfg->noise_reflactivity_thd, (float)thd));
    const Cloud white = euler_cluster(pass_through(raw, 3, (float)thd, 150.f));
    P4* o = static_cast<P4*>(out);
    int64_t k = 0;
    for (const P4& p : white) o[k++] = p;
    for (const P4& p : black) o[k++] = p;
    *n_out = k;
    return k > 0 ? 0 : 1;
}

} // extern "C"
