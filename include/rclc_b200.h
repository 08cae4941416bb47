/*
 * rclc_b200.h — C ABI of the B200-native calibration hot path (librclc_b200.so).
 *
 * The reference (zhijianglu/RCLC) has no plugin/FFI layer: its operators are header-inline C++
 * with PCL/Eigen/Ceres types in the signatures. Each entry point below replaces
 * one of those operators with plain pointers and sizes; the C++ operator surface in include/rclc_ops.hpp gives them
 * back the reference's names and argument meaning (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - every function returns an int status (RCLC_OK == 0), never throws, never exits;
 *  - `rclc_ctx` owns one CUDA stream + scratch, so concurrent host threads (the reference calls its
 *    operators inside `#pragma omp parallel for` over frames) use one ctx each;
 *  - point clouds are byte pointers to records of `stride` bytes, float x,y,z at offsets 0,4,8 and
 *    float intensity at byte offset `ioff` (float4 {x,y,z,I}: 16/12; pcl::PointXYZI: 32/16);
 *  - functions without a `_dev` suffix take HOST pointers and do their own H2D/D2H copies;
 *    `rclc_batch_*` keeps frames resident in HBM across calls;
 *  - there is NO CPU fallback: without a CUDA device every compute call returns RCLC_ERR_CUDA.
 */
#ifndef RCLC_B200_H
#define RCLC_B200_H

#include <stdint.h>
#include "rclc_config.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RCLC_OK = 0,
    RCLC_ERR_CUDA = 1,          /* CUDA runtime error (rclc_last_error_string has details)      */
    RCLC_ERR_INVALID = 2,       /* bad argument                                                 */
    RCLC_ERR_UNSUPPORTED = 3,   /* configuration outside what the kernels implement             */
    RCLC_ERR_NUMERIC = 4        /* NaN residual etc. (the reference would print and carry on)   */
};

typedef struct rclc_ctx rclc_ctx;
typedef struct rclc_batch rclc_batch;

/* Mirrors what opt_grid_fitting prints/returns (opt_grid_fitting.h:112-131) plus solver counters. */
typedef struct rclc_gridfit_report {
    int32_t outer_iterations;     /* executions of the while body, opt_grid_fitting.h:68        */
    int32_t status;               /* 0 ok; 1 = a Ceres solve ended in FAILURE (NaN residual; the reference goes on, and so do the
                                     corners); 2 = a frame without points; 3 = the composed frame has no inverse */
    int64_t pose_evals;           /* distinct (x,y,yaw) poses at which all residuals were evaluated */
    int64_t lm_iterations;        /* Ceres iterations summed over outer iterations                */
    double  first_initial_cost;   /* Summary::initial_cost of the first solve                     */
    double  last_final_cost;      /* Summary::final_cost of the last solve                        */
    double  se2_first[3];         /* (x, y, yaw deg) of the first solve                           */
    float   T_base_laser[16];     /* accumulated float transform, row-major                       */
    int64_t poses_swept;          /* poses the device evaluated, including speculative candidates the LM replay then discarded
                                     (pose_evals counts only the evaluations Ceres would have made)                        */
    int32_t rounds;               /* sweep rounds this frame took part in (pose_evals would be the count without speculation) */
    int32_t spec_mismatch;        /* speculated candidates that differed from the replayed ones: always 0 (self-check)    */
} rclc_gridfit_report;

const char* rclc_last_error_string(void);
int rclc_version(void);
uint64_t rclc_config_sizeof(void);

/* ---- context ---- */
int rclc_ctx_create(int device, rclc_ctx** out);
int rclc_ctx_destroy(rclc_ctx* ctx);
int rclc_ctx_synchronize(rclc_ctx* ctx);
void* rclc_ctx_stream(rclc_ctx* ctx);               /* cudaStream_t the ctx launches on          */
int rclc_ctx_device_info(rclc_ctx* ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* l2_bytes);
/* CUDA-event stopwatch on the ctx stream (bench.py times kernels on the stream that launches them). */
int rclc_ctx_timer_start(rclc_ctx* ctx);
int rclc_ctx_timer_stop_ms(rclc_ctx* ctx, float* ms_out);   /* records stop, synchronises, returns ms */
int rclc_ctx_flush_l2(rclc_ctx* ctx);               /* writes a buffer larger than L2            */
int64_t rclc_ctx_kernel_launches(rclc_ctx* ctx);    /* kernels launched by this ctx so far       */
/* the batched grid-fit solve runs its frames as `groups` independent groups on as many streams (1..32; default: chosen per batch,
 * 20 for frames of 200 k points and more, 10 for board-sized frames) */
int rclc_ctx_set_solve_groups(rclc_ctx* ctx, int32_t groups);
/* device time of the k_sweep kernel launched by the last rclc_batch_sweep (CUDA events recorded on the ctx stream
 * immediately before and after that launch; synchronises on the second one) */
int rclc_ctx_last_sweep_ms(rclc_ctx* ctx, float* ms_out);

/* ---- utils.h:286-316 generate_std_corners: col targets then row targets (opt_grid_fitting.h:78-94) ---- */
int rclc_generate_targets(const rclc_config* cfg, double* targets_xy, int32_t* is_row, int32_t* n_targets);

/* ---- BOARD_FITTING_COST::Evaluate x all targets (board_fitting_cost.h:74-280) at one pose.
 * residuals[nt], jac[nt*3] (may be NULL), acc[nt*16] (may be NULL): cost sum d,u,l,r | cost count
 * d,u,l,r | interval sum d,u,l,r | interval count d,u,l,r. ---- */
int rclc_gridfit_eval(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                      const double pose[3], double* residuals, double* jac, double* acc);

/* ---- opt_grid_fitting (opt_grid_fitting.h:25-132). T_bl row-major 4x4 in/out; corners (W-1)(H-1) x 3. ---- */
int rclc_gridfit_solve(rclc_ctx* ctx, const rclc_config* cfg, void* pts_io, int stride, int ioff, int64_t n,
                       double* T_bl_io, double* corners_out, rclc_gridfit_report* report);

/* ---- frames resident in HBM ---- */
int rclc_batch_create(rclc_ctx* ctx, const rclc_config* cfg, int32_t n_frames, int64_t max_pts_per_frame, rclc_batch** out);
int rclc_batch_destroy(rclc_batch* b);
/* host -> device; frame f gets n points. The copy is asynchronous. Pageable or strided sources are packed into pinned staging
 * first, so the caller's buffer is free on return; {x, y, z, I} records (stride 16, ioff 12) that already sit in page-locked
 * host memory are read in place by the DMA: keep such a buffer unchanged until the next call that returns results
 * (rclc_batch_gridfit_solve, rclc_batch_gmm_fit, rclc_batch_read_eval, rclc_ctx_synchronize). Uploads go to the stream of the
 * frame's solve group: a solve that follows them starts on the first frames while the later ones are still on the bus. */
int rclc_batch_upload(rclc_batch* b, int32_t frame, const void* pts_host, int stride, int ioff, int64_t n);
/* device float4 {x,y,z,I} -> batch storage */
int rclc_batch_set_frame_dev(rclc_batch* b, int32_t frame, const void* pts_dev_float4, int64_t n);
int64_t rclc_batch_total_points(rclc_batch* b);
/* bins every frame's cloud by lattice cell (once per cloud state); then single-pose passes */
int rclc_batch_bin(rclc_batch* b);
/* on != 0: every later sweep / solve of this batch runs the all-fp64 body (each predicate evaluated exactly like the reference,
 * board_fitting_cost.h:132-151) instead of the fp32 body with exact re-evaluation near decision boundaries. Same results bit for
 * bit, several times slower: the A/B reference of the fast body. */
int rclc_batch_set_exact_sweep(rclc_batch* b, int32_t on);
/* on != 0: plain sweeps and the solve of this batch run the streaming kernel (k_sweep_stream: one CTA per range of rows, fed by a
 * TMA bulk-copy ring, every row read from L2 once) instead of the gather kernel (k_sweep: one warp per target). Same accumulator
 * tables and LM trajectory bit for bit — the two kernels are each other's A/B check (tests/test_gpu_parity.py) — but the gather
 * kernel is the faster one on B200 (profiles/r02_stream_history.md) and stays the default. The all-fp64 body and the multi-start
 * search always use the gather kernel. */
int rclc_batch_set_stream_sweep(rclc_batch* b, int32_t on);
int rclc_batch_sweep(rclc_batch* b, const double* poses /* F*3 */);           /* accumulate only (async) */
int rclc_batch_read_eval(rclc_batch* b, double* residuals, double* jac, double* acc); /* F*nt[, *3, *16] */
/* full solve for all frames (async LM state machine on device); outputs on host */
int rclc_batch_gridfit_solve(rclc_batch* b, const double* T_bl /* F*16 or NULL */, double* T_bl_out /* F*16 or NULL */,
                             double* corners_out /* F*nc*3 */, rclc_gridfit_report* reports /* F */);
/* K5 on a resident, binned frame: P poses -> robustified cost per pose */
int rclc_batch_multistart(rclc_batch* b, int32_t frame, const double* poses, int32_t P, double* cost_out);
/* GMMFit::fit_1d on every frame's intensities; mu/sigma in-out F*2, priors F*2 */
int rclc_batch_gmm_fit(rclc_batch* b, double* mu_io, double* sigma_io, double* priors_out);

/* ---- GMMFit::fit_1d (src/GMMFit.cpp:7-108), K == 2 ---- */
int rclc_gmm_fit_1d(rclc_ctx* ctx, const void* intensity, int stride, int64_t n, int K, int iters,
                    double* mu_io, double* sigma_io, double* priors_out);

/* ---- pc_plane_prj (pc_process.h:41-57), in place ---- */
int rclc_plane_project(rclc_ctx* ctx, void* pts_io, int stride, int64_t n, const double plane[4]);
/* ---- pc_plane_fitting (opt_plane_param.hpp:45-88): Huber-robust plane (a,b,c,d) through the points with intensity >=
 * cfg->plane_reflactive_thd, Ceres LM from (0.1, 0.1, 1, 1). iterations / n_used may be NULL. ---- */
int rclc_plane_fit(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double plane_out[4],
                   int32_t* iterations, int64_t* n_used);
/* ---- pcl::UniformSampling (pc_process.h:111-115, apps/BoardSegmentation.cpp:58-61): per voxel of side `leaf` (grid anchored at
 * the cloud's min corner) the input point nearest the voxel centre (SURVEY.md Appendix A's definition of the filter, which the
 * oracle restates). idx_out (room for n) receives the kept point indices in ascending voxel-key order (PCL's order is its hash
 * map's). One voxel keeps one point in every PCL release, so the occupied voxels and the count are PCL's; WHICH point of a
 * voxel survives is pinned by the survey's definition, not by PCL's sources (not vendored: some releases measure the distance
 * against the voxel's integer index vector rather than its metric centre — a change of one expression in us_key /
 * orc_uniform_sample if a given PCL build has to be matched point for point). ---- */
int rclc_uniform_sample(rclc_ctx* ctx, const void* pts, int stride, int ioff, int64_t n, double leaf, int32_t* idx_out, int64_t* n_out);
/* ---- pcl::EuclideanClusterExtraction (pc_process.h:62-72, 131-139): connected components of "distance < tol", components with
 * min_size <= size <= max_size ranked by size descending (ties: smallest member index first). labels_out[i] = rank or -1;
 * sizes_out (room for n, may be NULL) receives the n_clusters sizes. ---- */
int rclc_euclidean_cluster(rclc_ctx* ctx, const void* pts, int stride, int ioff, int64_t n, double tol, int64_t min_size, int64_t max_size,
                           int32_t* labels_out, int32_t* n_clusters, int32_t* sizes_out);
/* ---- get_ref_pt (include/utils.h:160-202) + calc_laser_coor_v1 (include/pc_process.h:373-462) + line_fitting
 * (include/opt/opt_line_fitting.h:51-108): T_bl (row-major 4x4, board <- laser) from the n black-block centroids [n][3] of
 * eular_cluster_iter and the board plane. Host code: needs no context and no device. ---- */
int rclc_calc_laser_coor(const rclc_config* cfg, const double* centroids, int32_t n, const double plane[4], double Tbl_out[16]);
/* ---- cv::solvePnPRansac(objectPoints, imagePoints, K, dist, rvec, tvec, false, SOLVEPNP_UPNP) at apps/Calibrate.cpp:41-50, on
 * NORMALISED image points (Cam.imgPt2norm, Calibrate.cpp:36): the starting pose of pose_reprj_opt. obj [n][3], norm_pts [n][2],
 * thresh = reprojection threshold in normalised units (8 px / focal length for OpenCV's default), max_iters / confidence <= 0 pick
 * OpenCV's defaults (100, 0.99). R_out row-major camera <- laser, t_out; inlier_mask [n] and n_inliers may be NULL. Host code:
 * needs no context and no device. ---- */
int rclc_pnp_ransac(const double* obj, const double* norm_pts, int64_t n, double thresh, int32_t max_iters, double confidence, uint32_t seed,
                    double R_out[9], double t_out[3], uint8_t* inlier_mask, int32_t* n_inliers);
/* ---- pcl::transformPointCloud(Matrix4f) in place, row-major T ---- */
int rclc_transform_cloud(rclc_ctx* ctx, void* pts_io, int stride, int64_t n, const float T[16]);

/* ---- PCL SACSegmentation scoring (BoardSegmentation.cpp:64-72,90): H index triplets -> planes, validity, inlier counts ---- */
int rclc_ransac_plane_score(rclc_ctx* ctx, const void* pts, int stride, int64_t n, const int32_t* triplets, int32_t H,
                            double thr, float* planes_out, int32_t* valid_out, int32_t* counts_out);
/* selectWithinDistance mask + optimizeModelCoefficients (centroid, covariance, smallest eigenvector) */
int rclc_ransac_plane_select(rclc_ctx* ctx, const void* pts, int stride, int64_t n, const float plane[4], double thr,
                             uint8_t* mask_out, float refined_out[4], float centroid_out[3], int64_t* n_inliers);

/* ---- pose_reprj_opt (opt_reprj_calib.h:21-66). Tcl = (qx,qy,qz,qw,tx,ty,tz) in/out. The solve runs Ceres' Levenberg-Marquardt on the
 * device in the arithmetic of the CPU oracle, operation by operation (no fused multiply-add, sums in frame order, correctly rounded
 * pow / sin / cos): iteration count, costs and extrinsic are the oracle's bit for bit. ---- */
int rclc_pose_eval(rclc_ctx* ctx, const double* pc_corners, const double* img_norm, int32_t F, int32_t C,
                   const double Tcl[7], double* residuals, double* jac6);
int rclc_pose_refine(rclc_ctx* ctx, const rclc_config* cfg, const double* pc_corners, const double* img_norm,
                     int32_t F, int32_t C, double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations);

/* Test hooks of the pose refine (set per context, never read from the environment): shape 0 = automatic (a thread-block cluster from
 * 64 corner observations up), 1 = one CTA, 2 = the cluster at any size — both shapes return the same bits; cluster_debug 1 = whole-frame
 * evaluation on the cluster (one thread per frame instead of the per-corner / per-frame split); eval_split 1 = rclc_pose_eval through
 * the cluster's per-corner / per-frame split. tests/test_gpu_parity.py uses them to prove the bit-identity of the shapes. */
int rclc_ctx_set_pose_test_hooks(rclc_ctx* ctx, int32_t shape, int32_t cluster_debug, int32_t eval_split);

/* ---- colourise loop (Calibrate.cpp:134-148, PinholeCam.h:93-110) ---- */
int rclc_project_colorize(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int64_t n,
                          const float T_cl[16], const uint8_t* bgr, uint8_t* rgb_out, int32_t* pix_out, uint8_t* infov_out);

/* ---- multi-start pose search: P poses x one frame -> robustified cost per pose (K5) ---- */
int rclc_gridfit_multistart(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n,
                            const double* poses, int32_t P, double* cost_out);

/* ---- the per-frame chains on a cloud that is uploaded once and stays in HBM between the operators (csrc/calib.cu) ---- */
/* the per-cloud body of apps/BoardSegmentation.cpp:45-165: the chessboard's points out of one raw scan (lidar frame, x forward):
 * PassThrough x in [x_min, x_max] (:48), 8 mm UniformSampling (:58-61), planes peeled off by RANSAC (:64-113), EulerCluster of the
 * candidate (:116), cut_roi of the raw scan (:118-128), GMM split (:130-137), largest black / white cluster (:140-158).
 * out_xyzi: room for n float4 {x, y, z, I}; receives the white cluster followed by the black one (:160). *n_out == 0: no plane
 * qualified. */
int rclc_segment_board(rclc_ctx* ctx, const rclc_config* cfg, const void* pts, int stride, int ioff, int64_t n, double x_min, double x_max,
                       void* out_xyzi, int64_t* n_out);
/* est_board_corner (include/pc_process.h:583-646): one board-only cloud (float4 {x, y, z, I}, the T_base'd laser frame) in, the
 * (W-1) x (H-1) corners in that frame out [nc][3]. The cloud comes back as the reference leaves it (projected onto its plane, moved
 * into the board frame). T_bl_out (16, may be NULL): the refined board <- laser transform; report may be NULL.
 * *stage_failed (may be NULL): 0, or the stage that gave up: 1 plane fit, 2 block centroids, 3 board frame, 4 grid fit (no points or
 * no inverse; a Ceres FAILURE inside the fit is report->status 1 and, as in opt_grid_fitting.h:96-131, not a failed stage). */
int rclc_est_board_corner(rclc_ctx* ctx, const rclc_config* cfg, void* pts_io_xyzi, int64_t n, double* corners_out, double* T_bl_out,
                          rclc_gridfit_report* report, int32_t* stage_failed);
/* Stage 1 + stage 2 for F raw scans (apps/BoardSegmentation.cpp:31-165 -> apps/Calibrate.cpp:90-107 -> calib_opt :12-67): every scan
 * is uploaded once; the frames' chains run concurrently, one host worker (own stream, own scratch) per frame in flight like the
 * reference's `omp parallel for` over frames; the grid fits of all frames run together (rclc_batch_gridfit_solve); then, when
 * img_norm and Tcl_io are given, the extrinsic: the PnP start (unless RCLC_CALIB_KEEP_TCL: Tcl_io is the start) and pose_reprj_opt
 * over the frames that produced corners.
 *   scans[f]: n_pts[f] records of `stride` bytes, x y z at +0, intensity at +ioff (page-locked {x,y,z,I} records are read in place)
 *   img_norm [F][nc][2] normalised image corners (Cam.imgPt2norm) in est_board_corner's corner order, or NULL
 *   corners_out [F][nc][3]; frame_status [F]: 0 ok, 10 no board plane, 21 / 22 / 23 / 24 plane fit / block centroids / board frame /
 *   grid fit gave up; n_board [F] (may be NULL): points of the segmented board; reports [F] (may be NULL)
 *   Tcl_io (qx, qy, qz, qw, tx, ty, tz); pose_costs [2] initial / final (may be NULL); pose_iterations (may be NULL)
 *   stage_ms [4] (may be NULL): host wall time of per-frame chains | grid fit | extrinsic | total */
#define RCLC_CALIB_KEEP_TCL 1
int rclc_calibrate_frames(rclc_ctx* ctx, const rclc_config* cfg, const void* const* scans, const int64_t* n_pts, int32_t F, int stride, int ioff,
                          double x_min, double x_max, const double* img_norm, int32_t flags, double* corners_out, int32_t* frame_status,
                          int64_t* n_board, rclc_gridfit_report* reports, double* Tcl_io, double* pose_costs, int32_t* pose_iterations,
                          double* stage_ms);
/* host workers of rclc_calibrate_frames: 0 = one per host core (divided by LOCAL_WORLD_SIZE), at most one per frame */
int rclc_ctx_set_calib_workers(rclc_ctx* ctx, int32_t workers);
/* Eigen::Quaterniond(Matrix3d) as at apps/Calibrate.cpp:61; R row-major, q = (x, y, z, w). Host code. */
int rclc_quat_from_R(const double R[9], double q_xyzw[4]);

/* ---- multi-GPU exchanges (csrc/comm.cu): NCCL called from C++ on the context's stream; one process (one rclc_ctx) per GPU ----
 * NCCL is loaded at run time (libnccl.so.2, or the copy the process already holds); RCLC_ERR_UNSUPPORTED if there is none. */
typedef struct rclc_comm rclc_comm;
#define RCLC_COMM_ID_BYTES 128
/* rank 0 creates the id and hands it to the other ranks by any host means (MPI, a file, torch.distributed's store) */
int rclc_comm_get_unique_id(void* id_out_128);
int rclc_comm_create(rclc_ctx* ctx, const void* id_128, int32_t rank, int32_t world, rclc_comm** out);
int rclc_comm_destroy(rclc_comm* comm);
int64_t rclc_comm_collectives(rclc_comm* comm);         /* NCCL collectives issued so far */
/* calib_opt with the frames sharded over the ranks (apps/Calibrate.cpp:90-107 -> :12-67, include/opt/opt_reprj_calib.h:21-66):
 * every rank passes ITS frames' corners [F_local][C][3] and normalised image points [F_local][C][2] (F_local may differ between
 * ranks, and may be 0); ONE all-gather of C * 5 doubles per frame, then every rank runs the same deterministic solve over all
 * frames in rank-major order: the same bits on every rank, and the bits of rclc_pose_refine on the concatenated input.
 * Collective: every rank of the communicator must call it. */
int rclc_pose_refine_sharded(rclc_comm* comm, const rclc_config* cfg, const double* pc_local, const double* img_local, int32_t F_local, int32_t C,
                             double* Tcl_io, double* initial_cost, double* final_cost, int32_t* iterations, int32_t* F_total_out);
/* ONE frame's points split across ranks (include/opt/board_fitting_cost.h:211-235): every rank holds a share of the points of
 * each frame of its batch (same frames, same poses on all ranks); between rclc_batch_sweep and rclc_batch_read_eval this adds the
 * ranks' accumulator tables (82 x 16 doubles per frame). The accumulators are exact sums, so the result equals the unsplit
 * sweep bit for bit. Collective. */
int rclc_batch_allreduce_acc(rclc_batch* b, rclc_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* RCLC_B200_H */
