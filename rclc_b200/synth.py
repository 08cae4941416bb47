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
