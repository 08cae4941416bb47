// ops_dev.cuh — the operators of the per-frame chain on clouds that are RESIDENT in HBM.
//
// The C ABI of the single operators (rclc_uniform_sample, rclc_euclidean_cluster, rclc_ransac_plane_*, rclc_plane_fit,
// rclc_gmm_fit_1d, rclc_plane_project, rclc_transform_cloud) takes host clouds, the way the reference's call sites hold them; each
// of them is a thin wrapper — upload, the function below, download. The batched calibration chain (calib.cu) calls the functions
// below directly, so a scan is uploaded once and no cloud crosses the bus again before the corners come back.
//
// Conventions: clouds are float4 {x, y, z, I}; everything is enqueued on ctx->stream and uses ctx's scratch buffers (d_tmp, d_out,
// h_pin2), so one call at a time per context, and a result that lives in scratch is valid until the next call on that context.
// Control scalars (counts, the winning hypothesis, the state of a small solver) cross to the host behind a stream synchronise.
#pragma once
#include "common.cuh"

#include <vector>

namespace rclc {

// pcl::UniformSampling: *d_idx (ctx scratch) = indices of the kept points in ascending voxel-key order.
int dev_uniform_sample(rclc_ctx* ctx, const float4* d_pts, int64_t n, double leaf, const int** d_idx, int64_t* n_out);

// d_dst[i] = d_src[d_idx[i]]
int dev_gather(rclc_ctx* ctx, const float4* d_src, const int* d_idx, int64_t n, float4* d_dst);

// pcl::EuclideanClusterExtraction: d_label[i] (ctx scratch) = rank of the point's cluster (size descending, ties: smallest member
// index first) or -1; sizes of the ranked clusters on the host.
struct ClusterSet {
    const int* d_label = nullptr;
    std::vector<int> sizes;
    // in: a box that holds every finite point, when the caller knows one (saves the bounding-box pass and its read-back);
    // out: the box that was used. Any such box gives the same partition: it only anchors the cell grid.
    bool have_bbox = false;
    float bbox[6] = {0, 0, 0, 0, 0, 0};
};
int dev_euclidean_cluster(rclc_ctx* ctx, const float4* d_pts, int64_t n, double tol, int64_t min_size, int64_t max_size, ClusterSet* out);

// PCL SACSegmentation scoring: planes / validity / inlier counts of H hypotheses (host arrays; planes_out / valid_out may be null).
int dev_ransac_score(rclc_ctx* ctx, const float4* d_pts, int64_t n, const int32_t* h_triplets, int32_t H, double thr, float* planes_out,
                     int32_t* valid_out, int32_t* counts_out);
// selectWithinDistance into *d_mask (ctx scratch, one byte per point) + optimizeModelCoefficients.
int dev_ransac_select(rclc_ctx* ctx, const float4* d_pts, int64_t n, const float plane[4], double thr, const uint8_t** d_mask, float refined_out[4],
                      float centroid_out[3], int64_t* n_inliers);

// pc_plane_fitting
int dev_plane_fit(rclc_ctx* ctx, const rclc_config* cfg, const float4* d_pts, int64_t n, double plane_out[4], int32_t* iterations, int64_t* n_used);

// GMMFit::fit_1d on n floats `elem_stride` floats apart
int dev_gmm_fit(rclc_ctx* ctx, const float* d_x, int elem_stride, int64_t n, int iters, double* mu_io, double* sigma_io, double* priors_out);

// pose_reprj_opt with corners [F][C][3], image points [F][C][2] and the start (7 doubles) on the device
int pose_refine_dev(rclc_ctx* ctx, const rclc_config* cfg, const double* d_pc, const double* d_img, int32_t F, int32_t C, const double* d_T, double* Tcl_io,
                    double* initial_cost, double* final_cost, int32_t* iterations);

int dev_plane_project(rclc_ctx* ctx, float4* d_pts, int64_t n, const double plane[4]);
int dev_transform(rclc_ctx* ctx, float4* d_pts, int64_t n, const float T[16]);

} // namespace rclc
