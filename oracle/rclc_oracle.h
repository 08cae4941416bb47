/*
 * oracle/rclc_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * C entry points of the CPU oracle: a restatement of zhijianglu/RCLC's calibration hot path
 * (reference files cited per function in rclc_oracle.cpp). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library. The product
 * (librclc_b200.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures on this path, and
 * its arithmetic lives in un-vendored, un-pinned PCL / Ceres / Eigen; it
 * cannot be compiled in this image. The only reference-held known answers are the literals in
 * apps/calc_rmse.cpp, which pin quaternion->R, Euler extraction and rigid composition
 * (tests/test_oracle_golden.py).
 *
 * Point layout: `pts` is a byte pointer to records of `stride` bytes with float x,y,z at offset
 * 0,4,8 and float intensity at byte offset `ioff` (16-byte float4 {x,y,z,I}: stride=16, ioff=12;
 * pcl::PointXYZI: stride=32, ioff=16).
 */
#ifndef RCLC_ORACLE_H
#define RCLC_ORACLE_H
#include <stdint.h>
#include "../include/rclc_config.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_gridfit_report {
    int32_t outer_iterations;       /* executions of the while loop body, opt_grid_fitting.h:68 */
    int32_t status;                 /* 0 ok, 1 = some Ceres solve ended in FAILURE (NaN residual) */
    int64_t pose_evals;             /* distinct (x,y,yaw) poses at which all residuals were evaluated */
    int64_t evaluate_calls;         /* BOARD_FITTING_COST::Evaluate invocations (Ceres re-evaluates accepted points) */
    int64_t lm_iterations;          /* sum of Ceres iterations over outer iterations */
    double  first_initial_cost, last_final_cost;
    double  se2_first[3];           /* (x,y,yaw) solved in the first outer iteration */
    float   T_base_laser[16];       /* row-major accumulated float transform */
} orc_gridfit_report;

/* utils.h:286-316. targets_out: n_col col targets then n_row row targets, (x,y) doubles. */
int orc_generate_targets(const rclc_config* cfg, double* targets_xy, int32_t* is_row, int32_t* n_targets);

/* board_fitting_cost.h:74-262 for every target; neighbour sets per :264-280 (FLANN radius search
 * on the cloud as given). residuals[nt], jac[nt*3]; order = col targets then row targets
 * (opt_grid_fitting.h:78-94). acc (optional, nt*16): cost sums d,u,l,r; cost counts d,u,l,r;
 * interval sums d,u,l,r; interval counts d,u,l,r. */
int orc_gridfit_eval(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                     const double pose[3], double* residuals, double* jac, double* acc);

/* Robustified cost 0.5*sum rho(r^2) at `pose` exactly as Ceres would report it. */
int orc_gridfit_cost(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                     const double pose[3], double* cost);

/* opt_grid_fitting.h:25-132. pts_io is moved like pc_iter is NOT (cloud_src only changes under
 * apply_noise_test); T_bl row-major 4x4 double in/out; corners_out (W-1)*(H-1)*3. */
int orc_gridfit_solve(const rclc_config* cfg, void* pts_io, int stride, int ioff, int64_t n,
                      double* T_bl_io, double* corners_out, orc_gridfit_report* rep);

/* GMMFit.cpp:7-108 (K components, fixed iteration count). */
int orc_gmm_fit_1d(const void* intensity, int stride, int64_t n, int K, int iters,
                   double* mu_io, double* sigma_io, double* priors_out);

/* pc_process.h:41-57 */
int orc_plane_project(void* pts_io, int stride, int64_t n, const double plane[4]);

/* pcl::transformPointCloud(Matrix4f), float, row-major T. */
int orc_transform_cloud(void* pts_io, int stride, int64_t n, const float T[16]);

/* PCL SampleConsensusModelPlane::computeModelCoefficients + countWithinDistance for H index triplets.
 * planes_out[H*4] floats, valid_out[H], counts_out[H]. */
int orc_ransac_score(const void* pts, int stride, int64_t n, const int32_t* triplets, int H,
                     double thr, float* planes_out, int32_t* valid_out, int32_t* counts_out);
/* selectWithinDistance mask for one plane. */
int orc_plane_select(const void* pts, int stride, int64_t n, const float plane[4], double thr, uint8_t* mask_out);
/* RandomSampleConsensus::computeModel replay over precomputed counts/valid flags: returns index of
 * the winning hypothesis (or -1) and the number of iterations consumed. */
int orc_ransac_replay(const int32_t* counts, const int32_t* valid, int H, int64_t n_indices,
                      int max_iterations, double probability, int32_t* best_out, int32_t* iters_out);
/* optimizeModelCoefficients: centroid + covariance (double accumulation) + smallest eigenvector. */
int orc_plane_refine(const void* pts, int stride, int64_t n, const uint8_t* mask, const float plane_in[4],
                     float plane_out[4], float centroid_out[3]);
/* PCL sample stream: boost::mt19937(12345) partial Fisher-Yates. Writes H triplets. */
int orc_ransac_draw_triplets(int64_t n_indices, int H, int32_t* triplets_out);

/* pose_reprj_cost.hpp:28-55 : residual and 1x6 local Jacobian per frame (before loss). */
int orc_pose_eval(const double* pc_corners, const double* img_norm, int F, int C,
                  const double Tcl[7], double* residuals, double* jac6);
/* opt_reprj_calib.h:21-66 */
int orc_pose_refine(const rclc_config* cfg, const double* pc_corners, const double* img_norm, int F, int C,
                    double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations);
/* The same solve with a choice of libm (0: pow / sin / cos correctly rounded, the definition orc_pose_refine uses; 1: this machine's
 * libm), of the linear solver (0: DENSE_NORMAL_CHOLESKY, the reference's; 1: DENSE_QR) and the iteration log's cost / trust-region-radius columns (cap entries each, may be null). */
int orc_pose_refine_ex(const rclc_config* cfg, const double* pc_corners, const double* img_norm, int F, int C,
                       double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations, int libm_mode, int linear_solver,
                       double* cost_trace, double* radius_trace, int32_t cap, int32_t* n_trace);

/* Calibrate.cpp:134-148 + PinholeCam.h:93-110 : transform (float), project, FoV, gather BGR->RGB.
 * xyz stride as above; rgb_out[n*3] (untouched when not in FoV), pix_out[n*2] int32 (cvRound), infov_out[n]. */
int orc_project_colorize(const rclc_config* cfg, const void* pts, int stride, int64_t n, const float T_cl[16],
                         const uint8_t* bgr, uint8_t* rgb_out, int32_t* pix_out, uint8_t* infov_out);

/* utils.h:422-442 / calc_rmse.cpp:47-66 and helpers for the KAT. q = (w,x,y,z). */
int orc_quat_to_R(const double q_wxyz[4], double R[9]);
int orc_R_to_euler(const double R[9], double rpy[3]);
int orc_inverse4(const double T[16], double Tinv[16]);

/* opt_plane_param.hpp:45-88 */
int orc_plane_fit(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double plane_out[4],
                  int32_t* iterations, int64_t* n_used);

/* pcl::UniformSampling as used at pc_process.h:111-115 / BoardSegmentation.cpp:58-61: voxel grid of leaf
 * `leaf` anchored at the cloud's min corner, per voxel the input point nearest the voxel centre (float arithmetic, earliest index
 * on ties); kept indices in ascending voxel-key order (PCL's own order is its hash map's). idx_out has room for n entries. */
int orc_uniform_sample(const void* pts, int stride, int64_t n, double leaf, int32_t* idx_out, int64_t* n_out);

/* pcl::EuclideanClusterExtraction as used at pc_process.h:62-72, 131-139: connected components of "FLANN L2_Simple distance^2 <
 * (float)(tol*tol)", components with min_size <= size <= max_size kept, ranked by size descending (ties: smallest member index
 * first). labels_out[i] = rank or -1; sizes_out has room for n entries (only n_clusters are written). */
int orc_euclidean_cluster(const void* pts, int stride, int64_t n, double tol, int64_t min_size, int64_t max_size, int32_t* labels_out,
                          int32_t* n_clusters, int32_t* sizes_out);

/* get_ref_pt (utils.h:160-202) + calc_laser_coor_v1 (pc_process.h:373-462) + line_fitting (opt/opt_line_fitting.h:51-108): the
 * board frame T_bl (row-major 4x4, board <- laser) from the centroids of the black blocks and the board plane. Returns 0. */
int orc_calc_laser_coor(const rclc_config* cfg, const double* centroids, int32_t n, const double plane[4], double Tbl_out[16]);

/* ceres_lm.hpp's correctly rounded pow(q, 3), sin and cos (what the extrinsic solve uses in place of libm's). */
double orc_rn_cube(double q);
int orc_rn_sincos(double x, double* sin_out, double* cos_out);

/* The two Ceres tutorial problems whose solver logs the Ceres documentation prints (0: f(x) = 10 - x, 1: Powell's function), run
 * through ceres_lm.hpp (linear_solver 0 = DENSE_QR as published, 1 = DENSE_NORMAL_CHOLESKY): the `cost` and `tr_radius` column of every iteration, iteration 0 included. The pin of
 * the restated minimiser against published output of the dependency itself. Returns 0. */
int orc_lm_tutorial(int which, int linear_solver, double* x_io, int max_iterations, double* cost_trace, double* radius_trace, int32_t cap,
                    int32_t* n_trace, int32_t* termination);

/* The reference's host-side chains over the primitives above (rclc_pipeline.cpp): eular_cluster_iter (pc_process.h:104-211;
 * centroids_out has room for board_width * board_height doubles x 3), est_board_corner (pc_process.h:583-646; pts_io float4
 * {x, y, z, I}, stride 16, modified like the reference's cloud), the per-cloud body of apps/BoardSegmentation.cpp:45-165 (out:
 * float4 [n], white cluster then black cluster). */
int orc_eular_cluster_iter(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double reflactive_thd, double* centroids_out,
                           int32_t* n_out);
int orc_est_board_corner(const rclc_config* cfg, void* pts_io, int64_t n, double* corners_out, orc_gridfit_report* rep);
int orc_segment_board(const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double x_min, double x_max, void* out, int64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
