"""ctypes mirror of include/rclc_config.h (`rclc_config`), the POD copy of the reference's
`param CfgParam` (include/global.h:62-116) plus the literals it hard-codes."""
import ctypes as C


class RclcConfig(C.Structure):
    _fields_ = [
        ("board_width", C.c_int32), ("board_height", C.c_int32), ("board_order", C.c_int32),
        ("max_iter_time", C.c_int32), ("gride_side_length", C.c_double),
        ("neighbour_radius_rate", C.c_double), ("cost_radius_rate", C.c_double),
        ("uniform_sampling_radius", C.c_double), ("gridfit_huber_a", C.c_double),
        ("gridfit_function_tol", C.c_double), ("gridfit_max_lm_iters", C.c_int32),
        ("apply_noise_test", C.c_int32), ("noise_x", C.c_double), ("noise_y", C.c_double),
        ("noise_yaml", C.c_double),
        ("gmm_mu0", C.c_double * 2), ("gmm_sigma0", C.c_double * 2), ("gmm_iters", C.c_int32),
        ("noise_reflactivity_thd", C.c_int32), ("seg_sigma_factor", C.c_double),
        ("board_sigma_factor", C.c_double), ("white_cap", C.c_double),
        ("crop_x_min", C.c_double), ("crop_x_max", C.c_double), ("seg_uniform_radius", C.c_double),
        ("ransac_distance_thd", C.c_double), ("ransac_probability", C.c_double),
        ("ransac_max_iterations", C.c_int32), ("seg_cluster_min", C.c_int32),
        ("seg_cluster_max", C.c_int32), ("_pad0", C.c_int32),
        ("peel_min_remaining", C.c_double), ("plane_min_fraction", C.c_double),
        ("plane_closer_margin", C.c_double), ("angle_thd", C.c_double),
        ("seg_cluster_tolerance", C.c_double),
        ("reflactivity_dec_step", C.c_double), ("block_us_radius", C.c_double),
        ("black_pt_cluster_dst_thd", C.c_double),
        ("plane_reflactive_thd", C.c_double), ("plane_huber_a", C.c_double),
        ("pose_huber_a", C.c_double), ("pose_function_tol", C.c_double),
        ("pose_max_lm_iters", C.c_int32),
        ("cam_width", C.c_int32), ("cam_height", C.c_int32), ("_pad1", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
    ]

    def n_targets(self):
        w, h = self.board_width, self.board_height
        return (w - 1) * h + w * (h - 1)

    def n_corners(self):
        return (self.board_width - 1) * (self.board_height - 1)


def default_config(lib) -> RclcConfig:
    """Defaults come from the C side (rclc_config_default) so there is one source of truth."""
    cfg = RclcConfig()
    lib.rclc_config_default.argtypes = [C.POINTER(RclcConfig)]
    lib.rclc_config_default.restype = None
    lib.rclc_config_default(C.byref(cfg))
    return cfg
