"""Seeded synthetic inputs for the calibration hot path (numpy only; deterministic across boxes).

The reference ships no data and no scanner model (its simulator is a separate repository,
README.md:66-68), so these generators define the inputs of every BASELINE.json config
(SURVEY.md §8d). Points are float32 rows {x, y, z, intensity}.
"""
import numpy as np

GOLD = 0.6180339887498949
PLASTIC = 0.7548776662466927


def board_extent(W, H, side):
    """Board rectangle in the board frame implied by utils.h:286-316 (targets sit on interior edges)."""
    cx, cy = W // 2, H // 2
    return (-cx * side, (W - cx) * side, -cy * side, (H - cy) * side)


def square_is_black(u, v, W, H, side, black_parity=0):
    x0, _, y0, _ = board_extent(W, H, side)
    ci = np.floor((u - x0) / side).astype(np.int64)
    ri = np.floor((v - y0) / side).astype(np.int64)
    return ((ci + ri) % 2) == black_parity


def reflectance(is_black, rng):
    """black ~ N(15, 4^2), white ~ N(100, 12^2), clipped to [4, 149]."""
    n = is_black.shape[0]
    inten = np.where(is_black, rng.normal(15.0, 4.0, n), rng.normal(100.0, 12.0, n))
    return np.clip(inten, 4.0, 149.0)


def board_frame_cloud(n, seed, W=8, H=6, side=0.1, pert=(0.0, 0.0, 0.0), z_sigma=0.002,
                      pattern="raster", black_parity=0):
    """Board-only cloud already expressed in the (approximate) board frame, i.e. the input of
    opt_grid_fitting (include/opt/opt_grid_fitting.h:25). `pert` = (dx, dy, yaw_deg) is the residual
    misalignment the grid fit has to recover: stored points p satisfy R(yaw)·(p + (dx,dy)) = truth,
    the same convention as BOARD_FITTING_COST::Evaluate (board_fitting_cost.h:132-133).

    pattern "raster": dense Avia/Horizon-like line scan (spatially coherent order);
    pattern "rosette": MID-40-like golden-ratio order (spatially incoherent order).
    """
    rng = np.random.default_rng(seed)
    x0, x1, y0, y1 = board_extent(W, H, side)
    if pattern == "raster":
        lines = max(1, int(round(np.sqrt(n * (y1 - y0) / (x1 - x0)))))
        per = -(-n // lines)
        li = np.arange(n) // per
        ti = np.arange(n) % per
        v = y0 + (li + rng.uniform(0.2, 0.8, n)) * (y1 - y0) / lines
        u = x0 + (ti + rng.uniform(0.0, 1.0, n)) * (x1 - x0) / per
        odd = (li % 2) == 1
        u = np.where(odd, x0 + x1 - u, u)
    elif pattern == "rosette":
        i = np.arange(n, dtype=np.float64)
        u = x0 + np.modf(i * GOLD + rng.uniform())[0] * (x1 - x0)
        v = y0 + np.modf(i * PLASTIC + rng.uniform())[0] * (y1 - y0)
    else:
        raise ValueError(pattern)
    u = np.clip(u, x0 + 1e-9, x1 - 1e-9)
    v = np.clip(v, y0 + 1e-9, y1 - 1e-9)
    inten = reflectance(square_is_black(u, v, W, H, side, black_parity), rng)
    dx, dy, yaw = pert
    th = np.deg2rad(yaw)
    c, s = np.cos(th), np.sin(th)
    # p = R^T truth - t
    px = c * u + s * v - dx
    py = -s * u + c * v - dy
    pz = rng.normal(0.0, z_sigma, n)
    return np.stack([px, py, pz, inten], axis=1).astype(np.float32)


def gridfit_batch(n_frames, n_pts, seed0=1000, W=8, H=6, side=0.1, pattern="raster", max_shift=0.012,
                  max_yaw=1.0):
    """`n_frames` board-frame clouds with per-frame seeded misalignments (BASELINE config 2)."""
    frames, perts = [], []
    for f in range(n_frames):
        rng = np.random.default_rng(seed0 + f)
        pert = (rng.uniform(-max_shift, max_shift), rng.uniform(-max_shift, max_shift),
                rng.uniform(-max_yaw, max_yaw))
        frames.append(board_frame_cloud(n_pts, seed0 + f, W, H, side, pert, pattern=pattern))
        perts.append(pert)
    return frames, perts


def multistart_poses(nx=16, ny=16, nyaw=16, shift=0.03, yaw=4.0):
    """BASELINE config 3 lattice: x,y in [-3 cm, 3 cm] (16x16) x yaw in [-4 deg, 4 deg] (16)."""
    xs = np.linspace(-shift, shift, nx)
    ys = np.linspace(-shift, shift, ny)
    ws = np.linspace(-yaw, yaw, nyaw)
    g = np.stack(np.meshgrid(xs, ys, ws, indexing="ij"), axis=-1).reshape(-1, 3)
    return np.ascontiguousarray(g, dtype=np.float64)


def frustum_cloud(n, seed, zmin=1.0, zmax=100.0, half_fov_x=0.9, half_fov_y=0.6):
    """BASELINE config 5: points uniform in a camera frustum (slightly wider than the image so a
    fraction falls outside the FoV), float32 {x,y,z,pad}."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(zmin, zmax, n)
    x = rng.uniform(-half_fov_x * 1.2, half_fov_x * 1.2, n) * z
    y = rng.uniform(-half_fov_y * 1.2, half_fov_y * 1.2, n) * z
    return np.stack([x, y, z, np.zeros(n)], axis=1).astype(np.float32)


def noise_image(seed, W=1920, H=1080):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)


def _rot_xyz(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def std_corners(W=8, H=6, side=0.1):
    """utils.h:47-66 calc_transfered_standard_corners lattice (board frame, z = 0), row-major over (H-1) x (W-1)."""
    fcx, fcy = float((W - 1) // 2) * side, float((H - 1) // 2) * side
    return np.array([[fcx - c * side, fcy - r * side, 0.0] for r in range(H - 1) for c in range(W - 1)])


def calib_frame(f, n_pts, seed0=5000, W=8, H=6, side=0.1, pattern="rosette", pix_sigma=1e-4, T_cl_gt=None):
    """One frame of BASELINE config 4 (batch calibration), addressed by its GLOBAL index f so that any rank can generate
    exactly its shard: the board-only cloud in its approximate board frame, T_bl (laser -> approximate board frame, the input
    of opt_grid_fitting), the true laser-frame corners, and the normalised image corners seen through the ground-truth
    extrinsic T_cl_gt (camera <- laser) with a little pixel noise."""
    rng = np.random.default_rng(seed0 + f)
    pert = (rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), rng.uniform(-0.8, 0.8))
    cloud = board_frame_cloud(n_pts, seed0 + f, W, H, side, pert, pattern=pattern)
    # approximate board frame b' -> laser: the board 3.5-5.5 m down the lidar x axis, facing it, tilted a little
    R_lb = _rot_xyz(np.deg2rad(rng.uniform(-10, 10)), np.deg2rad(rng.uniform(-20, 20)), np.deg2rad(rng.uniform(-20, 20))) @ \
        np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    T_lb = np.eye(4); T_lb[:3, :3] = R_lb; T_lb[:3, 3] = [rng.uniform(3.5, 5.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3)]
    T_bl = np.linalg.inv(T_lb)
    # true board frame b from b': p_b = Rz(yaw) (p_b' + t)   (board_fitting_cost.h:132-133)
    th = np.deg2rad(pert[2])
    T_b_bp = np.eye(4); T_b_bp[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    T_b_bp[:3, 3] = T_b_bp[:3, :3] @ np.array([pert[0], pert[1], 0.0])
    c_std = np.c_[std_corners(W, H, side), np.ones((W - 1) * (H - 1))]
    corners_l = (T_lb @ np.linalg.inv(T_b_bp) @ c_std.T).T[:, :3]
    if T_cl_gt is None:
        T_cl_gt = gt_extrinsic()
    P = (T_cl_gt[:3, :3] @ corners_l.T).T + T_cl_gt[:3, 3]
    img_norm = P[:, :2] / P[:, 2:3] + rng.normal(0.0, pix_sigma, (len(P), 2))
    return cloud, T_bl, corners_l, img_norm


def gt_extrinsic():
    """Ground-truth camera <- laser transform of the synthetic rig: the camera looks down the lidar's +x axis."""
    T = np.eye(4)
    T[:3, :3] = _rot_xyz(np.deg2rad(1.5), np.deg2rad(-2.0), np.deg2rad(0.8)) @ np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
    T[:3, 3] = [0.06, -0.04, 0.03]
    return T


T_BASE = np.array([[0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 0.0], [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])   # src/global.cpp:116-122


def scan_extrinsic():
    """Ground-truth camera <- (T_base'd) laser transform of the raw-scan rig (the frame est_board_corner's corners live in)."""
    T = np.eye(4)
    T[:3, :3] = _rot_xyz(0.02, -0.03, 0.015)
    T[:3, 3] = [0.07, -0.05, 0.04]
    return T


def raw_scan(f, n_pts=100000, seed0=7000, W=8, H=6, side=0.1, board_share=0.62, pix_sigma=1e-4):
    """One RAW lidar scan of BASELINE configs[0] (the input of apps/BoardSegmentation.cpp), addressed by its global index f:
    a MID-40-like rosette (golden-ratio ray order, density rising towards the optical axis) cast into a scene of a W x H
    chessboard 4-5 m down the lidar's x axis, the ground (z = -1.2 m), a back wall (x = 5.9 m) and scattered clutter; points are
    float32 {x, y, z, intensity} in the LIDAR frame (x forward, z up). `board_share` of the rays are aimed at the board's
    neighbourhood (an integrated non-repetitive scan is densest where the sensor points), the rest sweep the whole 38 deg cone.

    Returns (raw [n, 4], img_norm [(W-1)*(H-1), 2], corners_cam [(W-1)*(H-1), 3]): the normalised image coordinates of the inner
    corners seen through scan_extrinsic() with a little pixel noise, and the true corners in the T_base'd laser frame, both in
    the order est_board_corner returns them for this rig."""
    rng = np.random.default_rng(seed0 + f)
    # the board in the camera-like frame (u right, v down, w forward), top-right square black
    Rb = _rot_xyz(0.12 * (rng.uniform() - 0.5), 0.3 * (rng.uniform() - 0.5), 0.16 * (rng.uniform() - 0.5))
    tb = np.array([0.8 * (rng.uniform() - 0.5), 0.4 * (rng.uniform() - 0.5), 4.0 + 1.0 * rng.uniform()])
    hw, hh = W * side / 2, H * side / 2
    n_board = int(n_pts * board_share)
    i = np.arange(n_pts, dtype=np.float64)
    # ray directions in the camera-like frame, (u/w, v/w)
    g1 = np.modf(i * GOLD + rng.uniform())[0]
    g2 = np.modf(i * PLASTIC + rng.uniform())[0]
    aim = np.arange(n_pts) % 50 < int(round(50 * board_share))            # interleaved, like a scan that keeps revisiting
    # aimed rays: a disc around the board centre's direction, 15 % wider than the board's half diagonal
    rho_b = 1.15 * np.hypot(hw, hh) / tb[2]
    r_a = rho_b * np.sqrt(g1)
    r_w = np.tan(np.deg2rad(19.2)) * g1 ** 0.8                            # whole cone, denser at the centre
    r = np.where(aim, r_a, r_w)
    phi = 2 * np.pi * g2
    du = np.where(aim, tb[0] / tb[2], 0.0) + r * np.cos(phi)
    dv = np.where(aim, tb[1] / tb[2], 0.0) + r * np.sin(phi)
    d = np.stack([du, dv, np.ones(n_pts)], axis=1)
    # board plane: normal = Rb[:, 2], through tb
    nb = Rb[:, 2]
    t_board = (nb @ tb) / (d @ nb)
    hit = d * t_board[:, None]
    ab = (hit - tb) @ Rb[:, :2]
    on_board = (np.abs(ab[:, 0]) <= hw) & (np.abs(ab[:, 1]) <= hh) & (t_board > 0)
    # ground: lidar z = -1.2  <=>  camera-like v = +1.2 ; wall: lidar x = 5.9  <=>  w = 5.9
    t_ground = np.where(d[:, 1] > 1e-6, 1.2 / np.maximum(d[:, 1], 1e-6), np.inf)
    t_wall = np.full(n_pts, 5.9)
    t_bg = np.minimum(t_ground, t_wall)
    is_ground = t_ground < t_wall
    t = np.where(on_board, t_board, t_bg)
    P = d * t[:, None]
    # range noise: 2 mm along the board normal on the board, 1 cm on the background
    P = P + np.where(on_board[:, None], nb[None, :] * rng.normal(0.0, 0.002, n_pts)[:, None],
                     np.where(is_ground[:, None], np.array([0.0, 1.0, 0.0])[None, :], np.array([0.0, 0.0, 1.0])[None, :]) *
                     rng.normal(0.0, 0.01, n_pts)[:, None])
    # reflectance: white 110, black 10 rising to 70 within 12 mm of a square's edge (the beam footprint), +- 2
    fu, fv = (ab[:, 0] + hw) / side, (ab[:, 1] + hh) / side
    ci = np.clip(np.floor(fu), 0, W - 1); ri = np.clip(np.floor(fv), 0, H - 1)
    e = np.minimum(np.minimum(fu - ci, ci + 1 - fu), np.minimum(fv - ri, ri + 1 - fv)) * side
    black = ((ci + ri) % 2) == 1
    I_board = np.maximum(4.0, np.where(black, 10.0 + 60.0 * np.maximum(0.0, 1.0 - e / 0.012), 110.0) + rng.normal(0.0, 2.0, n_pts))
    I = np.where(on_board, I_board, np.where(is_ground, 40.0, 60.0))
    # clutter: 3 % of the background rays return from somewhere in the volume
    cl = (~on_board) & (rng.uniform(size=n_pts) < 0.03)
    P[cl] = np.stack([-3.0 + 6.0 * rng.uniform(size=cl.sum()), 1.2 - 3.0 * rng.uniform(size=cl.sum()), 1.0 + 8.0 * rng.uniform(size=cl.sum())], axis=1)
    I[cl] = 150.0 * rng.uniform(size=cl.sum())
    # camera-like -> lidar: inverse of T_base (x_l = w, y_l = -u, z_l = -v)
    raw = np.stack([P[:, 2], -P[:, 0], -P[:, 1], I], axis=1).astype(np.float32)
    # the inner corners in est_board_corner's order for this rig, in the T_base'd laser frame
    k = np.arange((W - 1) * (H - 1))[::-1]                               # (the board frame of calc_laser_coor_v1 runs the other way)
    cb = np.stack([-(W - 2) * side / 2 + side * (k % (W - 1)), -(H - 2) * side / 2 + side * (k // (W - 1)), np.zeros(len(k))], axis=1)
    corners_cam = cb @ Rb.T + tb
    T = scan_extrinsic()
    Pc = corners_cam @ T[:3, :3].T + T[:3, 3]
    img_norm = Pc[:, :2] / Pc[:, 2:3] + rng.normal(0.0, pix_sigma, (len(k), 2))
    return raw, img_norm, corners_cam
