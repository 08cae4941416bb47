"""ctypes binding of librclc_b200.so (the C ABI declared in include/rclc_b200.h).

This is the thinnest possible host-language layer over the C ABI: tests and bench.py call the kernels
through exactly the entry points a C++ / cgo / JNI caller would bind. There is NO CPU fallback: if the
library is missing or no CUDA device is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

from .config import RclcConfig, default_config

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCLC_B200_LIB") or os.path.join(_HERE, "librclc_b200.so")      # (the override: A/B of experimental builds, scripts/ only)
_lib = None


class RclcError(RuntimeError):
    pass


class GridfitReport(C.Structure):
    _fields_ = [("outer_iterations", C.c_int32), ("status", C.c_int32), ("pose_evals", C.c_int64),
                ("lm_iterations", C.c_int64), ("first_initial_cost", C.c_double), ("last_final_cost", C.c_double),
                ("se2_first", C.c_double * 3), ("T_base_laser", C.c_float * 16), ("poses_swept", C.c_int64),
                ("rounds", C.c_int32), ("spec_mismatch", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RclcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.rclc_last_error_string.restype = C.c_char_p
        _lib.rclc_ctx_stream.restype = C.c_void_p
        _lib.rclc_ctx_kernel_launches.restype = C.c_int64
        _lib.rclc_batch_total_points.restype = C.c_int64
        _lib.rclc_config_sizeof.restype = C.c_uint64
        for name in ("rclc_ctx_destroy", "rclc_ctx_synchronize", "rclc_ctx_stream", "rclc_ctx_timer_start",
                     "rclc_ctx_flush_l2", "rclc_ctx_kernel_launches", "rclc_batch_destroy", "rclc_batch_bin",
                     "rclc_batch_total_points"):
            getattr(_lib, name).argtypes = [C.c_void_p]
    return _lib


def _chk(rc):
    if rc != 0:
        raise RclcError(f"rclc status {rc}: {lib().rclc_last_error_string().decode()}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _cloud(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] in (4, 8), pts.shape
    return pts, pts.shape[1] * 4, (12 if pts.shape[1] == 4 else 16)


def config() -> RclcConfig:
    cfg = default_config(lib())
    assert lib().rclc_config_sizeof() == C.sizeof(RclcConfig), "rclc_config layout mismatch"
    return cfg


def pnp_ransac(obj, norm_pts, thresh, max_iters=100, confidence=0.99, seed=1):
    """cv::solvePnPRansac stand-in on normalised image points (apps/Calibrate.cpp:41-50). Host code, no context."""
    obj = np.ascontiguousarray(obj, np.float64); uv = np.ascontiguousarray(norm_pts, np.float64)
    n = len(obj); R = np.zeros(9); t = np.zeros(3); mask = np.zeros(n, np.uint8); ni = C.c_int32()
    _chk(lib().rclc_pnp_ransac(_p(obj), _p(uv), C.c_int64(n), C.c_double(thresh), C.c_int32(max_iters), C.c_double(confidence), C.c_uint32(seed),
                               _p(R), _p(t), _p(mask), C.byref(ni)))
    return R.reshape(3, 3), t, mask.astype(bool)


def calc_laser_coor(cfg, centroids, plane):
    """calc_laser_coor_v1 (pc_process.h:373-462): board frame from the black-block centroids. Host code, no context."""
    cen = np.ascontiguousarray(centroids, np.float64); pl = np.ascontiguousarray(plane, np.float64); T = np.zeros(16)
    _chk(lib().rclc_calc_laser_coor(C.byref(cfg), _p(cen), C.c_int32(len(cen)), _p(pl), _p(T)))
    return T.reshape(4, 4)


def generate_targets(cfg):
    nt = cfg.n_targets()
    xy = np.zeros((nt, 2)); is_row = np.zeros(nt, np.int32); n = C.c_int32()
    _chk(lib().rclc_generate_targets(C.byref(cfg), _p(xy), _p(is_row), C.byref(n)))
    return xy[:n.value], is_row[:n.value]


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        _chk(lib().rclc_ctx_create(int(device), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().rclc_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _chk(lib().rclc_ctx_synchronize(self.h))

    @property
    def stream(self):
        return lib().rclc_ctx_stream(self.h)

    def device_info(self):
        sm = C.c_int32(); ma = C.c_int32(); mi = C.c_int32(); l2 = C.c_int64()
        _chk(lib().rclc_ctx_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), l2_bytes=l2.value)

    def timer_start(self):
        _chk(lib().rclc_ctx_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float()
        _chk(lib().rclc_ctx_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        _chk(lib().rclc_ctx_flush_l2(self.h))

    def set_solve_groups(self, n):
        _chk(lib().rclc_ctx_set_solve_groups(self.h, int(n)))

    def kernel_launches(self):
        return lib().rclc_ctx_kernel_launches(self.h)

    def last_sweep_ms(self):
        ms = C.c_float()
        _chk(lib().rclc_ctx_last_sweep_ms(self.h, C.byref(ms)))
        return ms.value

    # ---- single-frame operators (host pointers in, host pointers out) ----
    def gridfit_eval(self, cfg, pts, pose, want_acc=False):
        pts, stride, ioff = _cloud(pts)
        nt = cfg.n_targets()
        res = np.zeros(nt); jac = np.zeros((nt, 3)); acc = np.zeros((nt, 16))
        pose = np.ascontiguousarray(pose, np.float64)
        _chk(lib().rclc_gridfit_eval(self.h, C.byref(cfg), _p(pts), stride, ioff, C.c_int64(len(pts)), _p(pose), _p(res),
                                     _p(jac), _p(acc)))
        return (res, jac, acc) if want_acc else (res, jac)

    def gridfit_solve(self, cfg, pts, T_bl=None):
        pts, stride, ioff = _cloud(pts)
        pts = pts.copy()
        T = np.eye(4) if T_bl is None else np.ascontiguousarray(T_bl, np.float64).copy()
        corners = np.zeros((cfg.n_corners(), 3)); rep = GridfitReport()
        _chk(lib().rclc_gridfit_solve(self.h, C.byref(cfg), _p(pts), stride, ioff, C.c_int64(len(pts)), _p(T), _p(corners),
                                      C.byref(rep)))
        return corners, rep

    def gridfit_multistart(self, cfg, pts, poses):
        pts, stride, ioff = _cloud(pts)
        poses = np.ascontiguousarray(poses, np.float64); P = poses.shape[0]; cost = np.zeros(P)
        _chk(lib().rclc_gridfit_multistart(self.h, C.byref(cfg), _p(pts), stride, ioff, C.c_int64(len(pts)), _p(poses), P,
                                           _p(cost)))
        return cost

    def gmm_fit_1d(self, intensity, mu, sigma, iters=10):
        x = np.ascontiguousarray(intensity, np.float32)
        mu = np.array(mu, np.float64); sigma = np.array(sigma, np.float64); pri = np.zeros(2)
        _chk(lib().rclc_gmm_fit_1d(self.h, _p(x), 4, C.c_int64(len(x)), len(mu), iters, _p(mu), _p(sigma), _p(pri)))
        return mu, sigma, pri

    def plane_project(self, pts, plane):
        pts, stride, _ = _cloud(pts); pts = pts.copy()
        plane = np.ascontiguousarray(plane, np.float64)
        _chk(lib().rclc_plane_project(self.h, _p(pts), stride, C.c_int64(len(pts)), _p(plane)))
        return pts

    def plane_fit(self, cfg, pts):
        pts, stride, ioff = _cloud(pts)
        out = np.zeros(4); it = C.c_int32(); nu = C.c_int64()
        rc = lib().rclc_plane_fit(self.h, C.byref(cfg), _p(pts), stride, ioff, C.c_int64(len(pts)), _p(out), C.byref(it), C.byref(nu))
        if rc not in (0, 4):
            _chk(rc)
        return rc, out, it.value, nu.value

    def uniform_sample(self, pts, leaf):
        pts, stride, ioff = _cloud(pts)
        idx = np.zeros(len(pts), np.int32); n = C.c_int64()
        _chk(lib().rclc_uniform_sample(self.h, _p(pts), stride, ioff, C.c_int64(len(pts)), C.c_double(leaf), _p(idx), C.byref(n)))
        return idx[:n.value].copy()

    def euclidean_cluster(self, pts, tol, min_size, max_size):
        pts, stride, ioff = _cloud(pts)
        lab = np.zeros(len(pts), np.int32); sizes = np.zeros(max(len(pts), 1), np.int32); nc = C.c_int32()
        _chk(lib().rclc_euclidean_cluster(self.h, _p(pts), stride, ioff, C.c_int64(len(pts)), C.c_double(tol), C.c_int64(min_size),
                                          C.c_int64(max_size), _p(lab), C.byref(nc), _p(sizes)))
        return lab, sizes[:nc.value].copy()

    def transform_cloud(self, pts, T):
        pts, stride, _ = _cloud(pts); pts = pts.copy()
        T = np.ascontiguousarray(T, np.float32)
        _chk(lib().rclc_transform_cloud(self.h, _p(pts), stride, C.c_int64(len(pts)), _p(T)))
        return pts

    def ransac_plane_score(self, pts, triplets, thr):
        pts, stride, _ = _cloud(pts)
        tr = np.ascontiguousarray(triplets, np.int32); H = tr.shape[0]
        planes = np.zeros((H, 4), np.float32); valid = np.zeros(H, np.int32); counts = np.zeros(H, np.int32)
        _chk(lib().rclc_ransac_plane_score(self.h, _p(pts), stride, C.c_int64(len(pts)), _p(tr), H, C.c_double(thr),
                                           _p(planes), _p(valid), _p(counts)))
        return planes, valid, counts

    def ransac_plane_select(self, pts, plane, thr):
        pts, stride, _ = _cloud(pts)
        plane = np.ascontiguousarray(plane, np.float32)
        mask = np.zeros(len(pts), np.uint8); ref = np.zeros(4, np.float32); cen = np.zeros(3, np.float32); ni = C.c_int64()
        _chk(lib().rclc_ransac_plane_select(self.h, _p(pts), stride, C.c_int64(len(pts)), _p(plane), C.c_double(thr), _p(mask),
                                            _p(ref), _p(cen), C.byref(ni)))
        return mask, ref, cen, ni.value

    def pose_eval(self, pc, img, Tcl):
        pc = np.ascontiguousarray(pc, np.float64); img = np.ascontiguousarray(img, np.float64)
        F, Cn = pc.shape[0], pc.shape[1]
        Tcl = np.ascontiguousarray(Tcl, np.float64); r = np.zeros(F); j = np.zeros((F, 6))
        _chk(lib().rclc_pose_eval(self.h, _p(pc), _p(img), F, Cn, _p(Tcl), _p(r), _p(j)))
        return r, j

    def pose_refine(self, cfg, pc, img, Tcl):
        pc = np.ascontiguousarray(pc, np.float64); img = np.ascontiguousarray(img, np.float64)
        F, Cn = pc.shape[0], pc.shape[1]
        Tcl = np.array(Tcl, np.float64); c0 = C.c_double(); c1 = C.c_double(); it = C.c_int32()
        rc = lib().rclc_pose_refine(self.h, C.byref(cfg), _p(pc), _p(img), F, Cn, _p(Tcl), C.byref(c0), C.byref(c1), C.byref(it))
        if rc not in (0, 4):
            _chk(rc)
        return rc, Tcl, c0.value, c1.value, it.value

    def project_colorize(self, cfg, pts, T_cl, bgr):
        pts, stride, _ = _cloud(pts); n = len(pts)
        T = np.ascontiguousarray(T_cl, np.float32); bgr = np.ascontiguousarray(bgr, np.uint8)
        rgb = np.zeros((n, 3), np.uint8); pix = np.zeros((n, 2), np.int32); infov = np.zeros(n, np.uint8)
        _chk(lib().rclc_project_colorize(self.h, C.byref(cfg), _p(pts), stride, C.c_int64(n), _p(T), _p(bgr), _p(rgb), _p(pix),
                                         _p(infov)))
        return rgb, pix, infov

    # ---- the per-frame chains on resident clouds (csrc/calib.cu) ----
    def segment_board(self, cfg, raw, x_min=3.0, x_max=6.0):
        """apps/BoardSegmentation.cpp:45-165 on one raw scan -> board cloud [m, 4] (white cluster, then black)."""
        raw, stride, ioff = _cloud(raw)
        out = np.zeros((len(raw), 4), np.float32); n = C.c_int64()
        _chk(lib().rclc_segment_board(self.h, C.byref(cfg), _p(raw), stride, ioff, C.c_int64(len(raw)), C.c_double(x_min), C.c_double(x_max),
                                      _p(out), C.byref(n)))
        return out[:n.value].copy()

    def est_board_corner(self, cfg, board):
        """pc_process.h:583-646 on one board-only cloud [n, 4] in the T_base'd laser frame -> (stage_failed, corners, report, cloud as
        the reference leaves it)."""
        pts = np.ascontiguousarray(board, np.float32).copy()
        assert pts.ndim == 2 and pts.shape[1] == 4
        corners = np.zeros((cfg.n_corners(), 3)); rep = GridfitReport(); st = C.c_int32(); T = np.zeros(16)
        _chk(lib().rclc_est_board_corner(self.h, C.byref(cfg), _p(pts), C.c_int64(len(pts)), _p(corners), _p(T), C.byref(rep), C.byref(st)))
        return st.value, corners, rep, pts

    def set_pose_test_hooks(self, shape=0, cluster_debug=0, eval_split=0):
        _chk(lib().rclc_ctx_set_pose_test_hooks(self.h, int(shape), int(cluster_debug), int(eval_split)))

    def set_calib_workers(self, n):
        _chk(lib().rclc_ctx_set_calib_workers(self.h, int(n)))

    def calibrate_frames(self, cfg, scans, img_norm=None, x_min=3.0, x_max=6.0, Tcl0=None):
        """rclc_calibrate_frames: F raw scans (list of [n, 4] float32 arrays; pinned torch-backed arrays are read in place) ->
        dict(corners [F, nc, 3], status [F], n_board [F], reports, Tcl [7], pose_costs, pose_iterations, stage_ms [4])."""
        F = len(scans)
        scans = [np.ascontiguousarray(s, np.float32) for s in scans]
        assert all(s.ndim == 2 and s.shape[1] == 4 for s in scans)
        ptrs = (C.c_void_p * F)(*[s.ctypes.data for s in scans])
        n_pts = np.array([len(s) for s in scans], np.int64)
        nc = cfg.n_corners()
        corners = np.zeros((F, nc, 3)); status = np.zeros(F, np.int32); n_board = np.zeros(F, np.int64)
        reports = (GridfitReport * F)()
        Tcl = np.array([0, 0, 0, 1, 0, 0, 0], np.float64) if Tcl0 is None else np.array(Tcl0, np.float64)
        costs = np.zeros(2); its = C.c_int32(); ms = np.zeros(4)
        img = None if img_norm is None else np.ascontiguousarray(img_norm, np.float64)
        flags = 0 if Tcl0 is None else 1
        _chk(lib().rclc_calibrate_frames(self.h, C.byref(cfg), ptrs, _p(n_pts), F, 16, 12, C.c_double(x_min), C.c_double(x_max), _p(img), flags,
                                         _p(corners), _p(status), _p(n_board), reports, _p(Tcl) if img is not None else None, _p(costs),
                                         C.byref(its), _p(ms)))
        return dict(corners=corners, status=status, n_board=n_board, reports=reports, Tcl=Tcl, pose_costs=costs, pose_iterations=its.value,
                    stage_ms=ms)


class Comm:
    """rclc_comm: NCCL communicator of the library (one per context / GPU)."""

    def __init__(self, ctx: Context, unique_id: bytes, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.h = C.c_void_p()
        assert len(unique_id) == 128
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        _chk(lib().rclc_comm_create(ctx.h, buf, int(rank), int(world), C.byref(self.h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_char * 128)()
        _chk(lib().rclc_comm_get_unique_id(buf))
        return bytes(buf)

    def close(self):
        if self.h:
            lib().rclc_comm_destroy(self.h)
            self.h = C.c_void_p()

    def collectives(self):
        lib().rclc_comm_collectives.restype = C.c_int64
        return lib().rclc_comm_collectives(self.h)

    def pose_refine_sharded(self, cfg, pc_local, img_local, Tcl):
        pc = np.ascontiguousarray(pc_local, np.float64); img = np.ascontiguousarray(img_local, np.float64)
        F = pc.shape[0] if pc.size else 0
        Cn = pc.shape[1] if pc.size else cfg.n_corners()
        Tcl = np.array(Tcl, np.float64); c0 = C.c_double(); c1 = C.c_double(); it = C.c_int32(); ft = C.c_int32()
        rc = lib().rclc_pose_refine_sharded(self.h, C.byref(cfg), _p(pc) if F else None, _p(img) if F else None, F, Cn, _p(Tcl), C.byref(c0),
                                            C.byref(c1), C.byref(it), C.byref(ft))
        if rc not in (0, 4):
            _chk(rc)
        return rc, Tcl, c0.value, c1.value, it.value, ft.value

    def allreduce_acc(self, batch):
        _chk(lib().rclc_batch_allreduce_acc(batch.h, self.h))


class Batch:
    """Frames resident in HBM (rclc_batch_*)."""

    def __init__(self, ctx: Context, cfg: RclcConfig, n_frames: int, max_pts: int):
        self.ctx, self.cfg, self.n_frames = ctx, cfg, n_frames
        self.h = C.c_void_p()
        _chk(lib().rclc_batch_create(ctx.h, C.byref(cfg), n_frames, C.c_int64(max_pts), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().rclc_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, f, pts):
        pts, stride, ioff = _cloud(pts)
        _chk(lib().rclc_batch_upload(self.h, f, _p(pts), stride, ioff, C.c_int64(len(pts))))

    def set_frame_dev(self, f, dev_ptr, n):
        _chk(lib().rclc_batch_set_frame_dev(self.h, f, C.c_void_p(dev_ptr), C.c_int64(n)))

    def total_points(self):
        return lib().rclc_batch_total_points(self.h)

    def bin(self):
        _chk(lib().rclc_batch_bin(self.h))

    def set_exact_sweep(self, on=True):
        """All-fp64 sweep body (the A/B reference of the fast body) for every later sweep / solve of this batch."""
        _chk(lib().rclc_batch_set_exact_sweep(self.h, 1 if on else 0))

    def set_stream_sweep(self, on=True):
        """The streaming kernel (CTA per range of rows, TMA ring) instead of the gather kernel: each other's A/B check."""
        _chk(lib().rclc_batch_set_stream_sweep(self.h, 1 if on else 0))

    def sweep(self, poses):
        poses = np.ascontiguousarray(poses, np.float64)
        assert poses.shape == (self.n_frames, 3)
        _chk(lib().rclc_batch_sweep(self.h, _p(poses)))

    def read_eval(self, want_acc=False):
        nt = self.cfg.n_targets(); F = self.n_frames
        res = np.zeros((F, nt)); jac = np.zeros((F, nt, 3)); acc = np.zeros((F, nt, 16)) if want_acc else None
        _chk(lib().rclc_batch_read_eval(self.h, _p(res), _p(jac), _p(acc)))
        return (res, jac, acc) if want_acc else (res, jac)

    def gridfit_solve(self, T_bl=None):
        F = self.n_frames
        T = None if T_bl is None else np.ascontiguousarray(T_bl, np.float64)
        corners = np.zeros((F, self.cfg.n_corners(), 3)); reps = (GridfitReport * F)()
        _chk(lib().rclc_batch_gridfit_solve(self.h, _p(T), None, _p(corners), reps))
        return corners, list(reps)

    def multistart(self, frame, poses):
        poses = np.ascontiguousarray(poses, np.float64); P = poses.shape[0]; cost = np.zeros(P)
        _chk(lib().rclc_batch_multistart(self.h, frame, _p(poses), P, _p(cost)))
        return cost

    def gmm_fit(self, mu=None, sigma=None):
        F = self.n_frames
        mu = np.tile(np.array(self.cfg.gmm_mu0[:], np.float64), (F, 1)) if mu is None else np.array(mu, np.float64)
        sigma = np.tile(np.array(self.cfg.gmm_sigma0[:], np.float64), (F, 1)) if sigma is None else np.array(sigma, np.float64)
        pri = np.zeros((F, 2))
        _chk(lib().rclc_batch_gmm_fit(self.h, _p(mu), _p(sigma), _p(pri)))
        return mu, sigma, pri
