"""ctypes binding of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from rclc_b200.config import RclcConfig, default_config

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librclc_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def config() -> RclcConfig:
    return default_config(lib())


class GridfitReport(C.Structure):
    _fields_ = [("outer_iterations", C.c_int32), ("status", C.c_int32), ("pose_evals", C.c_int64),
                ("evaluate_calls", C.c_int64), ("lm_iterations", C.c_int64),
                ("first_initial_cost", C.c_double), ("last_final_cost", C.c_double),
                ("se2_first", C.c_double * 3), ("T_base_laser", C.c_float * 16)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cloud(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] in (4, 8)
    stride = pts.shape[1] * 4
    ioff = 12 if pts.shape[1] == 4 else 16
    return pts, stride, ioff


def generate_targets(cfg):
    nt = cfg.n_targets()
    xy = np.zeros((nt, 2)); is_row = np.zeros(nt, np.int32); n = C.c_int32()
    lib().orc_generate_targets(C.byref(cfg), _p(xy), _p(is_row), C.byref(n))
    assert n.value == nt
    return xy, is_row


def gridfit_eval(cfg, pts, pose, want_acc=False):
    pts, stride, ioff = _cloud(pts)
    nt = cfg.n_targets()
    res = np.zeros(nt); jac = np.zeros((nt, 3)); acc = np.zeros((nt, 16))
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    lib().orc_gridfit_eval(C.byref(cfg), _p(pts), stride, ioff, C.c_int64(pts.shape[0]), _p(pose), _p(res), _p(jac),
                           _p(acc))
    return (res, jac, acc) if want_acc else (res, jac)


def gridfit_cost(cfg, pts, pose):
    pts, stride, ioff = _cloud(pts)
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    cost = C.c_double()
    rc = lib().orc_gridfit_cost(C.byref(cfg), _p(pts), stride, ioff, C.c_int64(pts.shape[0]), _p(pose), C.byref(cost))
    return cost.value if rc == 0 else float("nan")


def gridfit_solve(cfg, pts, T_bl=None):
    pts, stride, ioff = _cloud(pts)
    pts = pts.copy()
    T = np.eye(4) if T_bl is None else np.ascontiguousarray(T_bl, dtype=np.float64).copy()
    corners = np.zeros((cfg.n_corners(), 3))
    rep = GridfitReport()
    rc = lib().orc_gridfit_solve(C.byref(cfg), _p(pts), stride, ioff, C.c_int64(pts.shape[0]), _p(T), _p(corners),
                                 C.byref(rep))
    return rc, corners, rep, T


def gmm_fit_1d(intensity, mu, sigma, iters=10):
    x = np.ascontiguousarray(intensity, dtype=np.float32)
    mu = np.array(mu, dtype=np.float64); sigma = np.array(sigma, dtype=np.float64)
    K = mu.shape[0]; pri = np.zeros(K)
    lib().orc_gmm_fit_1d(_p(x), 4, C.c_int64(x.shape[0]), K, iters, _p(mu), _p(sigma), _p(pri))
    return mu, sigma, pri


def plane_project(pts, plane):
    pts, stride, _ = _cloud(pts); pts = pts.copy()
    plane = np.ascontiguousarray(plane, dtype=np.float64)
    lib().orc_plane_project(_p(pts), stride, C.c_int64(pts.shape[0]), _p(plane))
    return pts


def transform_cloud(pts, T):
    pts, stride, _ = _cloud(pts); pts = pts.copy()
    T = np.ascontiguousarray(T, dtype=np.float32)
    lib().orc_transform_cloud(_p(pts), stride, C.c_int64(pts.shape[0]), _p(T))
    return pts


def ransac_score(pts, triplets, thr):
    pts, stride, _ = _cloud(pts)
    tr = np.ascontiguousarray(triplets, dtype=np.int32); H = tr.shape[0]
    planes = np.zeros((H, 4), np.float32); valid = np.zeros(H, np.int32); counts = np.zeros(H, np.int32)
    lib().orc_ransac_score(_p(pts), stride, C.c_int64(pts.shape[0]), _p(tr), H, C.c_double(thr), _p(planes), _p(valid),
                           _p(counts))
    return planes, valid, counts


def plane_select(pts, plane, thr):
    pts, stride, _ = _cloud(pts)
    plane = np.ascontiguousarray(plane, dtype=np.float32); mask = np.zeros(pts.shape[0], np.uint8)
    lib().orc_plane_select(_p(pts), stride, C.c_int64(pts.shape[0]), _p(plane), C.c_double(thr), _p(mask))
    return mask


def ransac_replay(counts, valid, n_indices, max_iterations=800, probability=0.99):
    counts = np.ascontiguousarray(counts, np.int32); valid = np.ascontiguousarray(valid, np.int32)
    best = C.c_int32(); iters = C.c_int32()
    lib().orc_ransac_replay(_p(counts), _p(valid), counts.shape[0], C.c_int64(n_indices), max_iterations,
                            C.c_double(probability), C.byref(best), C.byref(iters))
    return best.value, iters.value


def plane_refine(pts, mask, plane_in):
    pts, stride, _ = _cloud(pts)
    mask = np.ascontiguousarray(mask, np.uint8); pin = np.ascontiguousarray(plane_in, np.float32)
    pout = np.zeros(4, np.float32); cen = np.zeros(3, np.float32)
    lib().orc_plane_refine(_p(pts), stride, C.c_int64(pts.shape[0]), _p(mask), _p(pin), _p(pout), _p(cen))
    return pout, cen


def ransac_draw_triplets(n_indices, H):
    tr = np.zeros((H, 3), np.int32)
    lib().orc_ransac_draw_triplets(C.c_int64(n_indices), H, _p(tr))
    return tr


def pose_eval(pc_corners, img_norm, Tcl):
    pc = np.ascontiguousarray(pc_corners, np.float64); im = np.ascontiguousarray(img_norm, np.float64)
    F, Cn = pc.shape[0], pc.shape[1]
    Tcl = np.ascontiguousarray(Tcl, np.float64); r = np.zeros(F); j = np.zeros((F, 6))
    lib().orc_pose_eval(_p(pc), _p(im), F, Cn, _p(Tcl), _p(r), _p(j))
    return r, j


def pose_refine(cfg, pc_corners, img_norm, Tcl):
    pc = np.ascontiguousarray(pc_corners, np.float64); im = np.ascontiguousarray(img_norm, np.float64)
    F, Cn = pc.shape[0], pc.shape[1]
    Tcl = np.array(Tcl, np.float64)
    c0 = C.c_double(); c1 = C.c_double(); it = C.c_int32()
    rc = lib().orc_pose_refine(C.byref(cfg), _p(pc), _p(im), F, Cn, _p(Tcl), C.byref(c0), C.byref(c1), C.byref(it))
    return rc, Tcl, c0.value, c1.value, it.value


def pose_refine_ex(cfg, pc_corners, img_norm, Tcl, libm_mode=0, linear_solver=0):
    """pose_refine with a choice of libm (0: correctly rounded pow / sin / cos, pose_refine's definition; 1: this machine's), of the linear solver (0: DENSE_NORMAL_CHOLESKY, 1: DENSE_QR)
    and the iteration log: (rc, Tcl, initial cost, final cost, iterations, cost per iteration, trust-region radius per iteration)."""
    pc = np.ascontiguousarray(pc_corners, np.float64); im = np.ascontiguousarray(img_norm, np.float64)
    F, Cn = pc.shape[0], pc.shape[1]
    Tcl = np.array(Tcl, np.float64)
    c0 = C.c_double(); c1 = C.c_double(); it = C.c_int32(); n = C.c_int32()
    cost = np.zeros(1100); rad = np.zeros(1100)
    rc = lib().orc_pose_refine_ex(C.byref(cfg), _p(pc), _p(im), F, Cn, _p(Tcl), C.byref(c0), C.byref(c1), C.byref(it), C.c_int(libm_mode), C.c_int(linear_solver),
                                  _p(cost), _p(rad), C.c_int32(1100), C.byref(n))
    m = min(n.value, 1100)
    return rc, Tcl, c0.value, c1.value, it.value, cost[:m].copy(), rad[:m].copy()


def project_colorize(cfg, pts, T_cl, bgr):
    pts, stride, _ = _cloud(pts); n = pts.shape[0]
    T = np.ascontiguousarray(T_cl, np.float32); bgr = np.ascontiguousarray(bgr, np.uint8)
    rgb = np.zeros((n, 3), np.uint8); pix = np.zeros((n, 2), np.int32); infov = np.zeros(n, np.uint8)
    lib().orc_project_colorize(C.byref(cfg), _p(pts), stride, C.c_int64(n), _p(T), _p(bgr), _p(rgb), _p(pix), _p(infov))
    return rgb, pix, infov


def quat_to_R(q_wxyz):
    q = np.ascontiguousarray(q_wxyz, np.float64); R = np.zeros(9)
    lib().orc_quat_to_R(_p(q), _p(R)); return R.reshape(3, 3)


def R_to_euler(R):
    R = np.ascontiguousarray(R, np.float64).reshape(9); e = np.zeros(3)
    lib().orc_R_to_euler(_p(R), _p(e)); return e


def inverse4(T):
    T = np.ascontiguousarray(T, np.float64).reshape(16); o = np.zeros(16)
    lib().orc_inverse4(_p(T), _p(o)); return o.reshape(4, 4)


def plane_fit(cfg, pts):
    pts, stride, ioff = _cloud(pts)
    out = np.zeros(4); it = C.c_int32(); nu = C.c_int64()
    rc = lib().orc_plane_fit(C.byref(cfg), _p(pts), stride, ioff, C.c_int64(pts.shape[0]), _p(out), C.byref(it),
                             C.byref(nu))
    return rc, out, it.value, nu.value


def uniform_sample(pts, leaf):
    pts, stride, _ = _cloud(pts)
    idx = np.zeros(len(pts), np.int32); n = C.c_int64()
    lib().orc_uniform_sample(_p(pts), stride, C.c_int64(len(pts)), C.c_double(leaf), _p(idx), C.byref(n))
    return idx[:n.value].copy()


def euclidean_cluster(pts, tol, min_size, max_size):
    pts, stride, _ = _cloud(pts)
    lab = np.zeros(len(pts), np.int32); sizes = np.zeros(max(len(pts), 1), np.int32); nc = C.c_int32()
    lib().orc_euclidean_cluster(_p(pts), stride, C.c_int64(len(pts)), C.c_double(tol), C.c_int64(min_size), C.c_int64(max_size),
                                _p(lab), C.byref(nc), _p(sizes))
    return lab, sizes[:nc.value].copy()


def calc_laser_coor(cfg, centroids, plane):
    cen = np.ascontiguousarray(centroids, np.float64); pl = np.ascontiguousarray(plane, np.float64); T = np.zeros(16)
    rc = lib().orc_calc_laser_coor(C.byref(cfg), _p(cen), C.c_int32(len(cen)), _p(pl), _p(T))
    return rc, T.reshape(4, 4)


def segment_board(cfg, raw, x_min=3.0, x_max=6.0):
    """apps/BoardSegmentation.cpp:45-165 on one raw scan: (rc, board cloud [m, 4] = white cluster then black cluster)."""
    raw, stride, ioff = _cloud(raw)
    out = np.zeros((len(raw), 4), np.float32); n = C.c_int64()
    rc = lib().orc_segment_board(C.byref(cfg), _p(raw), stride, ioff, C.c_int64(len(raw)), C.c_double(x_min), C.c_double(x_max), _p(out), C.byref(n))
    return rc, out[:n.value].copy()


def est_board_corner(cfg, board):
    """include/pc_process.h:583-646 on one board-only cloud in the T_base'd laser frame: (rc, corners [(W-1)*(H-1), 3], report,
    the cloud as the reference leaves it: projected and moved into the board frame)."""
    pts = np.ascontiguousarray(board, np.float32).copy()
    assert pts.ndim == 2 and pts.shape[1] == 4
    corners = np.zeros(((cfg.board_width - 1) * (cfg.board_height - 1), 3)); rep = GridfitReport()
    rc = lib().orc_est_board_corner(C.byref(cfg), _p(pts), C.c_int64(len(pts)), _p(corners), C.byref(rep))
    return rc, corners, rep, pts


def eular_cluster_iter(cfg, pts, thd):
    pts, stride, ioff = _cloud(pts)
    cen = np.zeros((cfg.board_width * cfg.board_height, 3)); n = C.c_int32()
    rc = lib().orc_eular_cluster_iter(C.byref(cfg), _p(pts), stride, ioff, C.c_int64(len(pts)), C.c_double(thd), _p(cen), C.byref(n))
    return rc, cen[:n.value].copy()


def lm_tutorial(which, x0, linear_solver=0, max_iterations=100):
    """ceres_lm.hpp on a Ceres tutorial problem (0: f(x) = 10 - x, 1: Powell's function): (x, cost per iteration, trust-region
    radius per iteration, termination) — the columns a Ceres log prints, iteration 0 first."""
    x = np.array(x0, np.float64); cost = np.zeros(256); rad = np.zeros(256); n = C.c_int32(); term = C.c_int32()
    rc = lib().orc_lm_tutorial(C.c_int(which), C.c_int(linear_solver), _p(x), C.c_int(max_iterations), _p(cost), _p(rad), C.c_int32(256),
                               C.byref(n), C.byref(term))
    assert rc == 0 and n.value <= 256
    return x, cost[:n.value].copy(), rad[:n.value].copy(), term.value


def rn_cube(q):
    lib().orc_rn_cube.restype = C.c_double
    return lib().orc_rn_cube(C.c_double(q))


def rn_sincos(x):
    s = C.c_double(); c = C.c_double()
    lib().orc_rn_sincos(C.c_double(x), C.byref(s), C.byref(c))
    return s.value, c.value
