/*
 * rclc_config.h — POD mirror of the reference's `param CfgParam` (include/global.h:62-116) plus
 * the magic literals the reference hard-codes outside its YAML, so that host code,
 * CUDA kernels and the CPU oracle share one source of truth.
 *
 * Every field cites where the reference reads it. Defaults (rclc_config_default) are the shipped
 * cfg/config_real.yaml values, except board_height which BASELINE.json fixes to 6 (8x6 board).
 */
#ifndef RCLC_CONFIG_H
#define RCLC_CONFIG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rclc_config {
    /* ---- board (global.h:64-68, config_real.yaml:30-33) ---- */
    int32_t board_width;              /* board_size.x, squares along x (must be even)            */
    int32_t board_height;             /* board_size.y                                             */
    int32_t board_order;              /* 0: top-right square black                                */
    int32_t max_iter_time;            /* opt_grid_fitting.h:116, yaml:62 (20)                     */
    double  gride_side_length;        /* 0.1 m                                                    */

    /* ---- grid fitting (board_fitting_cost.h:41-59, opt_grid_fitting.h:71,96-101) ---- */
    double  neighbour_radius_rate;    /* 0.8  -> kd-tree radius 0.08                              */
    double  cost_radius_rate;         /* 0.35 -> cost disc 0.035, gradient disc 0.035*sqrt(2)     */
    double  uniform_sampling_radius;  /* 0 (skipped)                                              */
    double  gridfit_huber_a;          /* HuberLoss(0.1)                                           */
    double  gridfit_function_tol;     /* 1e-15                                                    */
    int32_t gridfit_max_lm_iters;     /* 1000                                                     */
    int32_t apply_noise_test;         /* yaml:68                                                  */
    double  noise_x, noise_y, noise_yaml; /* yaml:69-71 (yaw in degrees)                          */

    /* ---- GMM (pc_process.h:608-612, BoardSegmentation.cpp:134-137, GMMFit.cpp:37) ---- */
    double  gmm_mu0[2];               /* {15, 100}                                                */
    double  gmm_sigma0[2];            /* {15, 15}                                                 */
    int32_t gmm_iters;                /* 10                                                       */
    int32_t noise_reflactivity_thd;   /* 3 (yaml:36)                                              */
    double  seg_sigma_factor;         /* 1.7 (BoardSegmentation.cpp:137)                          */
    double  board_sigma_factor;       /* 2.0 (pc_process.h:612)                                   */
    double  white_cap;                /* 150 (BoardSegmentation.cpp:155)                          */

    /* ---- stage 1 segmentation literals (BoardSegmentation.cpp:45-113) ---- */
    double  crop_x_min, crop_x_max;   /* 3, 6                                                     */
    double  seg_uniform_radius;       /* 0.008                                                    */
    double  ransac_distance_thd;      /* 0.08                                                     */
    double  ransac_probability;       /* PCL default 0.99                                         */
    int32_t ransac_max_iterations;    /* 800                                                      */
    int32_t seg_cluster_min;          /* 500 (pc_process.h:69)                                    */
    int32_t seg_cluster_max;          /* 2500000                                                  */
    int32_t _pad0;
    double  peel_min_remaining;       /* 0.05                                                     */
    double  plane_min_fraction;       /* 0.02                                                     */
    double  plane_closer_margin;      /* 0.3                                                      */
    double  angle_thd;                /* 0.5 (yaml:37)                                            */
    double  seg_cluster_tolerance;    /* 0.03 (pc_process.h:68)                                   */

    /* ---- black block extraction (pc_process.h:104-211, yaml:49-51) ---- */
    double  reflactivity_dec_step;    /* 0.95                                                     */
    double  block_us_radius;          /* 0.003                                                    */
    double  black_pt_cluster_dst_thd; /* 0.0085                                                   */

    /* ---- plane fit (opt_plane_param.hpp:46-78) ---- */
    double  plane_reflactive_thd;     /* 100                                                      */
    double  plane_huber_a;            /* 0.1                                                      */

    /* ---- final pose refinement (opt_reprj_calib.h:29,51-57) ---- */
    double  pose_huber_a;             /* 0.2                                                      */
    double  pose_function_tol;        /* 1e-15                                                    */
    int32_t pose_max_lm_iters;        /* 1000                                                     */

    /* ---- camera (PinholeCam.h, yaml:2-9); distortion is image-side and out of scope ---- */
    int32_t cam_width, cam_height;
    int32_t _pad1;
    double  fx, fy, cx, cy;
} rclc_config;

/* Fills `cfg` with the defaults described above. */
void rclc_config_default(rclc_config* cfg);

#ifdef __cplusplus
}
#endif

#ifdef RCLC_CONFIG_IMPLEMENTATION
#include <string.h>
void rclc_config_default(rclc_config* c)
{
    memset(c, 0, sizeof(*c));
    c->board_width = 8;
    c->board_height = 6;
    c->board_order = 0;
    c->max_iter_time = 20;
    c->gride_side_length = 0.1;
    c->neighbour_radius_rate = 0.8;
    c->cost_radius_rate = 0.35;
    c->uniform_sampling_radius = 0.0;
    c->gridfit_huber_a = 0.1;
    c->gridfit_function_tol = 1e-15;
    c->gridfit_max_lm_iters = 1000;
    c->apply_noise_test = 0;
    c->noise_x = 0.01;
    c->noise_y = 0.01;
    c->noise_yaml = 0.0;
    c->gmm_mu0[0] = 15.0;
    c->gmm_mu0[1] = 100.0;
    c->gmm_sigma0[0] = 15.0;
    c->gmm_sigma0[1] = 15.0;
    c->gmm_iters = 10;
    c->noise_reflactivity_thd = 3;
    c->seg_sigma_factor = 1.7;
    c->board_sigma_factor = 2.0;
    c->white_cap = 150.0;
    c->crop_x_min = 3.0;
    c->crop_x_max = 6.0;
    c->seg_uniform_radius = 0.008;
    c->ransac_distance_thd = 0.08;
    c->ransac_probability = 0.99;
    c->ransac_max_iterations = 800;
    c->seg_cluster_min = 500;
    c->seg_cluster_max = 2500000;
    c->peel_min_remaining = 0.05;
    c->plane_min_fraction = 0.02;
    c->plane_closer_margin = 0.3;
    c->angle_thd = 0.5;
    c->seg_cluster_tolerance = 0.03;
    c->reflactivity_dec_step = 0.95;
    c->block_us_radius = 0.003;
    c->black_pt_cluster_dst_thd = 0.0085;
    c->plane_reflactive_thd = 100.0;
    c->plane_huber_a = 0.1;
    c->pose_huber_a = 0.2;
    c->pose_function_tol = 1e-15;
    c->pose_max_lm_iters = 1000;
    c->cam_width = 1920;
    c->cam_height = 1080;
    c->fx = 1000.0;
    c->fy = 1000.0;
    c->cx = 960.0;
    c->cy = 540.0;
}
#endif /* RCLC_CONFIG_IMPLEMENTATION */

#endif /* RCLC_CONFIG_H */
