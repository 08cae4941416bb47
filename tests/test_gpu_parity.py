"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> librclc_b200.so), against the
CPU oracle on identical seeded inputs. Bars (BASELINE.json north_star): accumulator tables / masks / counts /
pixels bit-exact; grid residuals 1e-5 relative (we get ~1e-13); extrinsic 1e-4 rad / 0.1 mm."""
import copy

import numpy as np
import pytest

from rclc_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from rclc_b200 import capi as m
    return m


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def cfg(capi):
    return capi.config()


POSES = [(0.0, 0.0, 0.0), (0.004, -0.003, 0.7), (-0.02, 0.015, -2.5), (0.03, 0.03, 4.0)]


@pytest.mark.parametrize("pattern", ["raster", "rosette"])
@pytest.mark.parametrize("pose", POSES)
def test_gridfit_eval_bit_exact_accumulators(ctx, cfg, oracle, ocfg, pattern, pose):
    pts = synth.board_frame_cloud(120000, 7, pert=(0.006, 0.004, 0.5), pattern=pattern)
    r0, j0, a0 = oracle.gridfit_eval(ocfg, pts, pose, want_acc=True)
    r1, j1, a1 = ctx.gridfit_eval(cfg, pts, pose, want_acc=True)
    assert np.array_equal(a0, a1), np.abs(a0 - a1).max()          # sums and counts: bit-exact
    np.testing.assert_allclose(r1, r0, rtol=1e-12, atol=1e-15)     # bar: 1e-5 relative
    np.testing.assert_allclose(j1, j0, rtol=1e-10, atol=1e-10)


def test_gridfit_eval_pcl_layout_and_8x5(ctx, cfg, oracle, ocfg):
    """pcl::PointXYZI records (32 B, intensity at byte 16) and the shipped 8x5 board."""
    c5, o5 = copy.copy(cfg), copy.copy(ocfg)
    c5.board_height = 5; o5.board_height = 5
    p4 = synth.board_frame_cloud(60000, 3, W=8, H=5, pert=(0.003, -0.002, 0.2))
    p8 = np.zeros((len(p4), 8), np.float32)
    p8[:, :3] = p4[:, :3]; p8[:, 3] = 1.0; p8[:, 4] = p4[:, 3]
    pose = (0.001, 0.002, 0.1)
    r0, j0, a0 = oracle.gridfit_eval(o5, p4, pose, want_acc=True)
    r1, j1, a1 = ctx.gridfit_eval(c5, p8, pose, want_acc=True)
    assert np.array_equal(a0, a1)
    np.testing.assert_allclose(r1, r0, rtol=1e-12)


def test_gridfit_eval_ragged_and_nan(ctx, cfg, oracle, ocfg):
    """Half a board: some targets see empty half-discs -> NaN residuals exactly where the oracle has them."""
    pts = synth.board_frame_cloud(40000, 5)
    pts = pts[pts[:, 0] < -0.05]
    r0, j0, a0 = oracle.gridfit_eval(ocfg, pts, (0, 0, 0), want_acc=True)
    r1, j1, a1 = ctx.gridfit_eval(cfg, pts, (0, 0, 0), want_acc=True)
    assert np.array_equal(a0, a1)
    assert np.array_equal(np.isnan(r0), np.isnan(r1)) and np.isnan(r1).any()
    np.testing.assert_allclose(r1[~np.isnan(r1)], r0[~np.isnan(r0)], rtol=1e-12)
    # tiny cloud (fewer points than one CTA chunk), far-away points that fall outside the lattice
    tiny = np.concatenate([pts[:37], np.array([[5.0, 5.0, 0.0, 50.0], [-3.0, 0.2, 0.0, 60.0]], np.float32)])
    a0 = oracle.gridfit_eval(ocfg, tiny, (0.001, 0, 0), want_acc=True)[2]
    a1 = ctx.gridfit_eval(cfg, tiny, (0.001, 0, 0), want_acc=True)[2]
    assert np.array_equal(a0, a1)


def _check_solve(corners, rep, oc, orep):
    assert rep.outer_iterations == orep.outer_iterations
    assert rep.pose_evals == orep.pose_evals and rep.lm_iterations == orep.lm_iterations
    assert rep.status == orep.status
    np.testing.assert_allclose(rep.first_initial_cost, orep.first_initial_cost, rtol=1e-11)
    np.testing.assert_allclose(rep.last_final_cost, orep.last_final_cost, rtol=1e-11)
    np.testing.assert_allclose(list(rep.se2_first), list(orep.se2_first), rtol=1e-7, atol=1e-10)
    # bar: corners to 0.1 mm. The corners are the inverse of the composed board frame times the lattice, in double, rounded operation
    # by operation as the oracle rounds them: they differ from the oracle's by at most an ulp (2e-15 m), and not at all whenever the
    # float frame T_base_laser has the oracle's bits - which it has unless the LM iterate behind it (equal to 1e-15: its steps are
    # computed with fused multiply-adds) falls on the other side of a float rounding
    np.testing.assert_allclose(np.array(rep.T_base_laser), np.array(orep.T_base_laser), rtol=0, atol=2e-7)
    np.testing.assert_allclose(corners, oc, rtol=0, atol=2e-15)
    if np.array_equal(np.array(rep.T_base_laser, np.float32), np.array(orep.T_base_laser, np.float32)):
        assert corners.tobytes() == oc.tobytes(), np.abs(corners - oc).max()


@pytest.mark.parametrize("seed,pert,pattern", [(11, (0.01, 0.01, 0.0), "raster"), (12, (0.008, -0.005, 0.6), "rosette"),
                                               (13, (0.0, 0.0, 0.0), "raster"), (14, (-0.004, 0.011, -0.8), "raster")])
def test_gridfit_solve_matches_oracle_trajectory(ctx, cfg, oracle, ocfg, seed, pert, pattern):
    pts = synth.board_frame_cloud(100000, seed, pert=pert, pattern=pattern)
    rc, oc, orep, _ = oracle.gridfit_solve(ocfg, pts)
    assert rc == 0
    corners, rep = ctx.gridfit_solve(cfg, pts)
    _check_solve(corners, rep, oc, orep)


def test_gridfit_solve_with_T_bl(ctx, cfg, oracle, ocfg):
    pts = synth.board_frame_cloud(50000, 21, pert=(0.005, 0.0, 0.0))
    a = 0.3
    T_bl = np.array([[np.cos(a), -np.sin(a), 0, 0.5], [np.sin(a), np.cos(a), 0, -0.2], [0, 0, 1, 4.0], [0, 0, 0, 1]])
    rc, oc, orep, _ = oracle.gridfit_solve(ocfg, pts, T_bl)
    corners, rep = ctx.gridfit_solve(cfg, pts, T_bl)
    _check_solve(corners, rep, oc, orep)


def test_gridfit_solve_failure_status(ctx, cfg, oracle, ocfg):
    pts = synth.board_frame_cloud(20000, 5)
    pts = pts[pts[:, 0] < -0.05]
    rc, oc, orep, _ = oracle.gridfit_solve(ocfg, pts)
    corners, rep = ctx.gridfit_solve(cfg, pts)
    assert rep.status == orep.status == 1
    assert rep.outer_iterations == orep.outer_iterations and rep.pose_evals == orep.pose_evals
    np.testing.assert_allclose(corners, oc, atol=1e-9)


def test_batch_solve_ragged_frames(capi, ctx, cfg, oracle, ocfg):
    sizes = [30000, 80000, 1000, 55555, 0, 64000]
    frames = [synth.board_frame_cloud(max(n, 1), 100 + i, pert=(0.002 * i, -0.003, 0.1 * i))[:n] for i, n in enumerate(sizes)]
    b = capi.Batch(ctx, cfg, len(sizes), max(sizes))
    for f, p in enumerate(frames):
        if len(p):
            b.upload(f, p)
    corners, reps = b.gridfit_solve()
    for f, p in enumerate(frames):
        if len(p) == 0:
            assert reps[f].status == 2
            continue
        rc, oc, orep, _ = oracle.gridfit_solve(ocfg, p)
        _check_solve(corners[f], reps[f], oc, orep)
    # solving again from the resident (pristine) clouds gives identical answers
    corners2, reps2 = b.gridfit_solve()
    assert np.array_equal(corners, corners2)
    mu, sg, pri = b.gmm_fit()
    for f, p in enumerate(frames):
        if len(p) < 10:
            continue
        m0, s0, p0 = oracle.gmm_fit_1d(p[:, 3], [15, 100], [15, 15])
        np.testing.assert_allclose(mu[f], m0, rtol=1e-10); np.testing.assert_allclose(sg[f], s0, rtol=1e-9)
        np.testing.assert_allclose(pri[f], p0, rtol=1e-10)
    b.close()


def test_batch_single_pose_pass(capi, ctx, cfg, oracle, ocfg):
    frames, perts = synth.gridfit_batch(4, 50000, seed0=500)
    b = capi.Batch(ctx, cfg, 4, 50000)
    for f, p in enumerate(frames):
        b.upload(f, p)
    b.bin()
    poses = np.array(perts)
    b.sweep(poses)
    res, jac, acc = b.read_eval(want_acc=True)
    for f in range(4):
        r0, j0, a0 = oracle.gridfit_eval(ocfg, frames[f], poses[f], want_acc=True)
        assert np.array_equal(acc[f], a0)
        np.testing.assert_allclose(res[f], r0, rtol=1e-12)
    # multistart on the resident frame 2
    P = synth.multistart_poses(4, 4, 4)
    cost = b.multistart(2, P)
    for k in range(0, len(P), 5):
        np.testing.assert_allclose(cost[k], oracle.gridfit_cost(ocfg, frames[2], P[k]), rtol=1e-12)
    assert int(np.argmin(cost)) == int(np.argmin([oracle.gridfit_cost(ocfg, frames[2], p) for p in P]))
    b.close()


@pytest.mark.parametrize("n", [30000, 200000, 400000])      # one cluster launch with cached posteriors | without | a launch per pass
def test_gmm_parity_and_labels(ctx, oracle, n):
    pts = synth.board_frame_cloud(n, 21)
    mu, sg, pri = ctx.gmm_fit_1d(pts[:, 3], [15, 100], [15, 15])
    m0, s0, p0 = oracle.gmm_fit_1d(pts[:, 3], [15, 100], [15, 15])
    np.testing.assert_allclose(mu, m0, rtol=1e-11); np.testing.assert_allclose(sg, s0, rtol=1e-10)
    np.testing.assert_allclose(pri, p0, rtol=1e-11)
    # labels (PassThrough [3, thd] / [thd, 150], BoardSegmentation.cpp:140-157): identical except within 1 ulp of thd
    for k in (1.7, 2.0):
        t1, t0 = np.float32(mu[0] + k * sg[0]), np.float32(m0[0] + k * s0[0])
        l1 = (pts[:, 3] >= 3) & (pts[:, 3] <= t1); l0 = (pts[:, 3] >= 3) & (pts[:, 3] <= t0)
        diff = l1 != l0
        assert np.all(np.abs(pts[diff, 3] - t0) <= np.spacing(t0))


def test_plane_project_and_transform_bit_exact(ctx, oracle):
    rng = np.random.default_rng(3)
    n = 100003
    pts = np.zeros((n, 4), np.float32)
    pts[:, 2] = rng.uniform(3, 6, n); pts[:, 0] = rng.uniform(-1, 1, n); pts[:, 1] = rng.uniform(-1, 1, n); pts[:, 3] = rng.uniform(0, 150, n)
    plane = np.array([0.1, -0.2, 1.0, -4.5])
    assert np.array_equal(ctx.plane_project(pts, plane), oracle.plane_project(pts, plane))
    a = 0.01
    T = np.array([[np.cos(a), -np.sin(a), 0, 0.003], [np.sin(a), np.cos(a), 0, -0.002], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    assert np.array_equal(ctx.transform_cloud(pts, T), oracle.transform_cloud(pts, T))
    T3 = np.array([[0, -1, 0, 0.1], [0, 0, -1, 0.2], [1, 0, 0, 0.3], [0, 0, 0, 1]], np.float32)     # T_base (global.cpp:116-122)
    assert np.array_equal(ctx.transform_cloud(pts, T3), oracle.transform_cloud(pts, T3))


def _tilted_board(n, seed, nrm, dist, sigma=0.004, outliers=0):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-0.4, 0.4, n); pts[:, 1] = rng.uniform(-0.3, 0.3, n)
    nrm = np.asarray(nrm, float); nrm = nrm / np.linalg.norm(nrm)
    pts[:, 2] = (dist - nrm[0] * pts[:, 0] - nrm[1] * pts[:, 1]) / nrm[2] + rng.normal(0, sigma, n)
    pts[:, 3] = np.where(rng.uniform(size=n) < 0.5, 15, 110)
    if outliers:
        k = rng.choice(n, outliers, replace=False)
        pts[k, 2] += rng.uniform(0.3, 1.0, outliers)
    return pts, nrm


@pytest.mark.parametrize("case", [(20000, 12, (0.15, -0.1, 1.0), 4.0, 0), (300000, 13, (-0.3, 0.2, 1.0), 5.5, 3000),
                                  (777, 14, (0.0, 0.0, 1.0), 3.0, 20)])
def test_plane_fit_matches_oracle(ctx, cfg, oracle, ocfg, case):
    """pc_plane_fitting (opt_plane_param.hpp:45-88): same Ceres trajectory (iteration count) and the same plane BIT FOR BIT: the sums
    over the points are exact and order-free on both sides and every other expression has one defined association (the chain
    downstream needs the same bits: tests/test_chain_sensitivity.py)."""
    n, seed, nrm, dist, outl = case
    pts, _ = _tilted_board(n, seed, nrm, dist, outliers=outl)
    rc0, p0, it0, used0 = oracle.plane_fit(ocfg, pts)
    rc1, p1, it1, used1 = ctx.plane_fit(cfg, pts)
    assert rc1 == rc0 == 0 and used1 == used0 == int((pts[:, 3] >= 100).sum())
    assert it1 == it0
    assert np.array_equal(p1, p0), (p1 - p0)


def test_plane_fit_pcl_layout_and_degenerate(ctx, cfg, oracle, ocfg):
    pts, _ = _tilted_board(5000, 15, (0.1, 0.1, 1.0), 4.0)
    pcl = np.zeros((len(pts), 8), np.float32); pcl[:, :3] = pts[:, :3]; pcl[:, 4] = pts[:, 3]     # PointXYZI: intensity at +16
    rc0, p0, it0, _ = oracle.plane_fit(ocfg, pts)
    rc1, p1, it1, _ = ctx.plane_fit(cfg, pcl)
    assert rc1 == rc0 == 0 and it1 == it0
    assert np.array_equal(p1, p0)
    # no point passes the reflectance filter: zero residual blocks, the start point comes back in both
    dark = pts.copy(); dark[:, 3] = 10
    rc0, p0, _, u0 = oracle.plane_fit(ocfg, dark)
    rc1, p1, _, u1 = ctx.plane_fit(cfg, dark)
    assert u0 == u1 == 0 and rc0 == rc1
    np.testing.assert_array_equal(p1, p0)


def _cluster_scene(n, seed):
    from tests.test_oracle_restatement import _scene
    return _scene(n, seed)


@pytest.mark.parametrize("n,leaf", [(50000, 0.008), (400000, 0.003), (1000, 0.5)])
def test_uniform_sample_identical_indices(ctx, oracle, n, leaf):
    """pcl::UniformSampling: the same point per voxel, in the same (voxel key) order."""
    pts = _cluster_scene(n, 31)
    pts[5, 1] = np.inf; pts[77, 2] = np.nan
    assert np.array_equal(ctx.uniform_sample(pts, leaf), oracle.uniform_sample(pts, leaf))
    pcl = np.zeros((n, 8), np.float32); pcl[:, :3] = pts[:, :3]; pcl[:, 4] = pts[:, 3]
    assert np.array_equal(ctx.uniform_sample(pcl, leaf), oracle.uniform_sample(pts, leaf))


@pytest.mark.parametrize("n,tol,min_size", [(60000, 0.03, 500), (60000, 0.0085, 3), (3000, 0.05, 1)])
def test_euclidean_cluster_identical_labels(ctx, oracle, n, tol, min_size):
    """pcl::EuclideanClusterExtraction: same partition, same ranking (size descending, smallest member first)."""
    pts = _cluster_scene(n, 32)
    pts = pts[oracle.uniform_sample(pts, 0.004)]                     # the reference always thins before it clusters
    pts[3, 0] = np.nan
    l0, s0 = oracle.euclidean_cluster(pts, tol, min_size, 2500000)
    l1, s1 = ctx.euclidean_cluster(pts, tol, min_size, 2500000)
    assert np.array_equal(s1, s0) and np.array_equal(l1, l0)
    assert len(s0) >= 3 and l0[3] == -1


def test_uniform_and_cluster_degenerate_inputs(ctx, oracle):
    """One point, no finite point, a cluster above max_size, an empty cloud (the C ABI refuses it; the shims never send it)."""
    from rclc_b200.capi import RclcError
    one = np.array([[1, 2, 3, 50]], np.float32)
    assert list(ctx.uniform_sample(one, 0.01)) == list(oracle.uniform_sample(one, 0.01)) == [0]
    l1, s1 = ctx.euclidean_cluster(one, 0.1, 1, 10); l0, s0 = oracle.euclidean_cluster(one, 0.1, 1, 10)
    assert list(l1) == list(l0) == [0] and list(s1) == list(s0) == [1]
    bad = np.full((64, 4), np.nan, np.float32); bad[::2, 0] = np.inf
    assert len(ctx.uniform_sample(bad, 0.01)) == len(oracle.uniform_sample(bad, 0.01)) == 0
    l1, s1 = ctx.euclidean_cluster(bad, 0.1, 1, 10); l0, s0 = oracle.euclidean_cluster(bad, 0.1, 1, 10)
    assert np.array_equal(l1, l0) and (l1 == -1).all() and len(s1) == len(s0) == 0
    # a dense blob of 500 and a far pair: max_size 100 rejects the blob, min_size 3 the pair
    rng = np.random.default_rng(9)
    blob = np.zeros((502, 4), np.float32); blob[:500, :3] = rng.uniform(0, 0.05, (500, 3)); blob[500:, :3] = [[9, 9, 9], [9.01, 9, 9]]
    for mn, mx in ((1, 100), (3, 1000), (1, 1000), (1, 500), (501, 1000)):
        l1, s1 = ctx.euclidean_cluster(blob, 0.06, mn, mx); l0, s0 = oracle.euclidean_cluster(blob, 0.06, mn, mx)
        assert np.array_equal(l1, l0) and np.array_equal(s1, s0), (mn, mx)
    with pytest.raises(RclcError):
        ctx.uniform_sample(np.zeros((0, 4), np.float32), 0.01)
    with pytest.raises(RclcError):
        ctx.euclidean_cluster(np.zeros((0, 4), np.float32), 0.1, 1, 10)


def test_ransac_counts_masks_bit_exact(ctx, oracle):
    rng = np.random.default_rng(5)
    n = 60000
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = 4.0 + rng.normal(0, 0.03, n); pts[:, 1] = rng.uniform(-1, 1, n); pts[:, 2] = rng.uniform(-1, 1, n)
    pts[: n // 3, :3] = rng.uniform(-3, 3, (n // 3, 3)) + [4, 0, 0]
    H = 800
    tr = oracle.ransac_draw_triplets(n, H)
    tr[5] = [7, 7, 9]                                        # degenerate sample -> invalid model
    p0, v0, c0 = oracle.ransac_score(pts, tr, 0.08)
    p1, v1, c1 = ctx.ransac_plane_score(pts, tr, 0.08)
    assert np.array_equal(v0, v1) and v1[5] == 0
    assert np.array_equal(p0.view(np.uint32), p1.view(np.uint32))      # planes bit-exact
    assert np.array_equal(c0, c1)                                      # inlier counts exact
    best, iters = oracle.ransac_replay(c1, v1, n)
    m0 = oracle.plane_select(pts, p1[best], 0.08)
    m1, ref1, cen1, ni = ctx.ransac_plane_select(pts, p1[best], 0.08)
    assert np.array_equal(m0, m1) and ni == int(m0.sum()) == c1[best]
    ref0, cen0 = oracle.plane_refine(pts, m0, p1[best])
    # optimizeModelCoefficients: the inlier moments are exact, order-free sums on both sides => the refined model, its centroid and
    # the mask re-selected with it are the oracle's bit for bit
    assert np.array_equal(ref1.view(np.uint32), ref0.view(np.uint32)) and np.array_equal(cen1.view(np.uint32), cen0.view(np.uint32))
    r0 = oracle.plane_select(pts, ref0, 0.08)
    r1 = ctx.ransac_plane_select(pts, ref1, 0.08)[0]
    assert np.array_equal(r0, r1)


def _pose_problem(oracle, F, seed=9, noise=1e-4, grid=(5, 7)):
    rng = np.random.default_rng(seed)
    Cn = grid[0] * grid[1]
    q = np.array([0.02, -0.015, 0.01, 0.0]); q[3] = np.sqrt(1 - (q[:3] ** 2).sum())
    t = np.array([0.07, -0.08, 0.02])
    R = oracle.quat_to_R([q[3], q[0], q[1], q[2]])
    pc = np.zeros((F, Cn, 3)); img = np.zeros((F, Cn, 2))
    for f in range(F):
        base = np.array([rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), rng.uniform(3.5, 5.5)])
        g = np.array([[0.3 - 0.1 * c, 0.2 - 0.1 * r, 0.0] for r in range(grid[0]) for c in range(grid[1])])
        a = rng.uniform(-0.3, 0.3, 3)
        Rb = oracle.quat_to_R(np.r_[np.sqrt(1 - (a ** 2).sum() / 4), a / 2])
        pc[f] = g @ Rb.T + base
        Pc = pc[f] @ R.T + t
        img[f] = Pc[:, :2] / Pc[:, 2:3] + rng.normal(0, noise, (Cn, 2))
    T0 = np.r_[q, t] + np.array([0.005, -0.004, 0.003, 0, 0.03, -0.02, 0.04]); T0[:4] /= np.linalg.norm(T0[:4])
    return pc, img, T0, np.r_[q, t]


@pytest.mark.parametrize("F,seed", [(1, 2), (2, 3), (5, 9), (20, 9), (33, 5), (64, 3), (256, 9), (257, 6), (300, 4), (700, 5), (1700, 6)])
def test_pose_refine_is_the_oracles_bit_for_bit(ctx, cfg, oracle, ocfg, F, seed):
    """The extrinsic solve is held to the oracle BIT FOR BIT: every frame's residual and local Jacobian, and then the whole
    Levenberg-Marquardt run - iteration count, initial and final cost, the extrinsic's seven doubles - in both launch shapes.
    Nothing weaker is stable: on few-frame problems the solve spends its 1000-iteration cap wandering over a non-smooth floor and
    the capped iterate follows the last bit of every sum (tests/test_pose_sensitivity.py: the oracle itself moves by 0.1 - 6 mm
    under a rounding-level change of its own arithmetic). pose.cu therefore runs the oracle's operations in the oracle's order
    (no fused multiply-add, frame-order sums, correctly rounded pow / sin / cos on both sides). 1700 frames: the frames' rows no
    longer fit shared memory and live in global memory; 1, 2, 33, 257 frames: the edges of the dealing over CTAs, warps and sweeps."""
    pc, img, T0, Tgt = _pose_problem(oracle, F, seed=seed)
    r0, j0 = oracle.pose_eval(pc, img, T0)
    r1, j1 = ctx.pose_eval(pc, img, T0)
    assert r1.tobytes() == r0.tobytes() and j1.tobytes() == j0.tobytes()
    rc0, Ta, c0a, c1a, ita = oracle.pose_refine(ocfg, pc, img, T0)
    assert rc0 == 0
    try:
        for shape in (0, 1, 2):      # the default choice | one CTA | the cluster at any size
            ctx.set_pose_test_hooks(shape=shape)
            rc1, Tb, c0b, c1b, itb = ctx.pose_refine(cfg, pc, img, T0)
            assert rc1 == 0
            assert (itb, c0b, c1b) == (ita, c0a, c1a), (shape, itb, ita, c1b, c1a)
            assert Tb.tobytes() == Ta.tobytes(), (shape, np.abs(Tb - Ta).max())
    finally:
        ctx.set_pose_test_hooks()
    # and the solve does its job: the cost falls by three orders; from 64 frames on the extrinsic is the synthetic truth to 0.2 mm /
    # 1e-4 rad (20 frames: 0.9 mm, 5 frames: 3 cm - under-determined)
    if F >= 5:
        assert c1a < 0.01 * c0a
    if F in (64, 256, 300, 700, 1700):
        assert np.abs(Ta[4:] - Tgt[4:]).max() < 2e-4 and 2 * np.linalg.norm(Ta[:4] * np.sign(Ta[3]) - Tgt[:4] * np.sign(Tgt[3])) < 1e-4


@pytest.mark.parametrize("F,grid", [(12, (3, 4)), (40, (2, 2)), (6, (9, 11)), (3, (1, 1))])
def test_pose_refine_other_board_sizes_bit_for_bit(ctx, cfg, oracle, ocfg, F, grid):
    """Boards with 12, 4, 99 and 1 corner(s) per frame: the per-corner / per-frame split, the eight-lane frame assembly and the strided
    arg-max do not depend on the 35 corners of the 8 x 6 board."""
    pc, img, T0, _ = _pose_problem(oracle, F, seed=17, grid=grid)
    rc0, Ta, c0a, c1a, ita = oracle.pose_refine(ocfg, pc, img, T0)
    try:
        for shape in (1, 2):
            ctx.set_pose_test_hooks(shape=shape)
            rc1, Tb, c0b, c1b, itb = ctx.pose_refine(cfg, pc, img, T0)
            assert (rc1 != 0) == (rc0 != 0)
            assert (itb, c0b, c1b) == (ita, c0a, c1a) or (np.isnan(c0a) and np.isnan(c0b)), (shape, itb, ita)
            assert Tb.tobytes() == Ta.tobytes(), (shape, np.abs(Tb - Ta).max())
    finally:
        ctx.set_pose_test_hooks()


@pytest.mark.parametrize("where", ["start", "later"])
def test_pose_refine_failure_paths_like_the_oracle(ctx, cfg, oracle, ocfg, where):
    """A non-finite residual: at the start Ceres fails the solve and leaves the parameters alone; at a candidate it rejects the step
    and goes on. Both shapes do what the oracle does."""
    pc, img, T0, _ = _pose_problem(oracle, 6, seed=23)
    if where == "start":
        pc[2, 7, :] = np.nan
    else:
        # a corner that sits in the camera plane for the start but not for its neighbours: X / Z blows up only along the way
        R = oracle.quat_to_R([T0[3], T0[0], T0[1], T0[2]])
        p = np.linalg.solve(R, np.array([0.1, 0.1, 1e-9]) - T0[4:])
        pc[2, 7, :] = p
    rc0, Ta, c0a, c1a, ita = oracle.pose_refine(ocfg, pc, img, T0)
    try:
        for shape in (1, 2):
            ctx.set_pose_test_hooks(shape=shape)
            rc1, Tb, c0b, c1b, itb = ctx.pose_refine(cfg, pc, img, T0)
            assert (rc1 != 0) == (rc0 != 0), (rc0, rc1)
            assert Tb.tobytes() == Ta.tobytes() and itb == ita, (shape, itb, ita, Tb, Ta)
            assert (c0b == c0a or (np.isnan(c0a) and np.isnan(c0b))) and (c1b == c1a or (np.isnan(c1a) and np.isnan(c1b)))
    finally:
        ctx.set_pose_test_hooks()


@pytest.mark.parametrize("F,seed", [(5, 9), (20, 9), (64, 3), (256, 9), (300, 4), (700, 5)])
def test_pose_refine_cluster_frame_split_is_exact(ctx, cfg, oracle, F, seed, monkeypatch):
    """The cluster shape deals (frame, corner) pairs over 16 CTAs and rebuilds each frame's sums from per-corner factors. That
    must not change one bit of a frame's residual or Jacobian: the same cluster solve with every frame evaluated by ONE thread
    (the one-CTA kernel's routine) has to return identical bits after up to 1000 iterations."""
    pc, img, T0, _ = _pose_problem(oracle, F, seed=seed)
    out = []
    try:
        for dbg in (0, 1):           # 1: whole-frame evaluation by one thread
            ctx.set_pose_test_hooks(shape=2, cluster_debug=dbg)
            rc, T, c0, c1, it = ctx.pose_refine(cfg, pc, img, T0)
            out.append((rc, T.tobytes(), c0, c1, it))
        assert out[0] == out[1]
        ctx.set_pose_test_hooks()
        r0, j0 = ctx.pose_eval(pc, img, T0)
        ctx.set_pose_test_hooks(eval_split=1)
        r1, j1 = ctx.pose_eval(pc, img, T0)
        assert r0.tobytes() == r1.tobytes() and j0.tobytes() == j1.tobytes()
    finally:
        ctx.set_pose_test_hooks()


@pytest.mark.parametrize("F,seed", [(5, 9), (20, 9), (64, 3), (256, 9), (300, 4), (700, 5)])
def test_pose_refine_cluster_equals_one_cta(ctx, cfg, oracle, ocfg, F, seed, monkeypatch):
    """The two launch shapes of the extrinsic solve return the same bits (each is also held to the oracle's bits in
    test_pose_refine_is_the_oracles_bit_for_bit); this one prints their wall times."""
    import time
    pc, img, T0, _ = _pose_problem(oracle, F, seed=seed)
    out = {}
    try:
        for mode in (1, 2):          # one CTA | the cluster at any size
            ctx.set_pose_test_hooks(shape=mode)
            ctx.pose_refine(cfg, pc, img, T0)
            t0 = time.perf_counter()
            rc, T, c0, c1, it = ctx.pose_refine(cfg, pc, img, T0)
            out[str(mode) + "ms"] = (time.perf_counter() - t0) * 1e3
            out[mode] = (rc, T.tobytes(), c0, c1, it)
    finally:
        ctx.set_pose_test_hooks()
    print(f"pose refine F={F}: one CTA {out['1ms']:.2f} ms, cluster {out['2ms']:.2f} ms, {out[1][4]} iterations")
    assert out[1] == out[2]
    if F in (64, 300, 700):
        rco, To, c0o, c1o, ito = oracle.pose_refine(ocfg, pc, img, T0)
        assert rc == rco == 0
        assert np.abs(T[4:] - To[4:]).max() < 1e-4
        assert 2 * np.linalg.norm(T[:4] * np.sign(T[3]) - To[:4] * np.sign(To[3])) < 1e-4


def test_project_colorize_exact(ctx, cfg, oracle, ocfg):
    pts = synth.frustum_cloud(300000, 2)
    pts[:7, 2] = [0.0, -1.0, 1e-30, 5.0, -5.0, 0.0, 2.0]; pts[5, 0] = 0.0; pts[5, 1] = 0.0     # Z = 0, NaN, behind camera
    img = synth.noise_image(4, cfg.cam_width, cfg.cam_height)
    a = 0.02
    T = np.array([[np.cos(a), 0, np.sin(a), 0.07], [0, 1, 0, -0.08], [-np.sin(a), 0, np.cos(a), 0.02], [0, 0, 0, 1]], np.float32)
    rgb0, pix0, in0 = oracle.project_colorize(ocfg, pts, T, img)
    rgb1, pix1, in1 = ctx.project_colorize(cfg, pts, T, img)
    assert np.array_equal(in0, in1) and np.array_equal(pix0, pix1) and np.array_equal(rgb0, rgb1)
    assert 0.2 < in1.mean() < 0.95


def test_kernel_launch_counter(ctx, cfg):
    before = ctx.kernel_launches()
    ctx.gmm_fit_1d(np.linspace(5, 120, 1000, dtype=np.float32), [15, 100], [15, 15])
    assert ctx.kernel_launches() - before == 1             # the whole EM of a small cloud is one cluster launch
    before = ctx.kernel_launches()
    ctx.gmm_fit_1d(np.linspace(5, 120, 300000, dtype=np.float32), [15, 100], [15, 15])
    assert ctx.kernel_launches() - before == 10            # large clouds: one pass per iteration, ten iterations


def test_fast_sweep_equals_exact_sweep(capi, ctx, cfg, oracle, ocfg, monkeypatch):
    """The production sweep (fp32 predicates + exact re-evaluation inside the rounding band + pose-aware candidate
    culling) must produce the SAME accumulator tables as the all-fp64 sweep and the oracle, for small and large poses."""
    rng = np.random.default_rng(77)
    frames = [synth.board_frame_cloud(150000, 900 + i, pert=(0.004 * i, -0.003 * i, 0.3 * i), pattern="rosette" if i % 2 else "raster")
              for i in range(4)]
    tables = {}
    for mode in ("0", "1"):
        b = capi.Batch(ctx, cfg, 4, 150000)
        b.set_exact_sweep(mode == "1")
        for f, p in enumerate(frames):
            b.upload(f, p)
        b.bin()
        out = []
        prng = np.random.default_rng(5)
        for trial in range(12):
            scale = [1e-4, 2e-3, 0.03][trial % 3]
            poses = np.stack([prng.uniform(-scale, scale, 4), prng.uniform(-scale, scale, 4), prng.uniform(-130 * scale, 130 * scale, 4)], axis=1)
            b.sweep(poses)
            out.append((poses, b.read_eval(want_acc=True)[2].copy()))
        tables[mode] = out
        b.close()
    for (p0, a0), (p1, a1) in zip(tables["0"], tables["1"]):
        assert np.array_equal(p0, p1)
        assert np.array_equal(a0, a1), np.abs(a0 - a1).max()
    # and the oracle on a few of them
    for poses, acc in tables["0"][:6]:
        for f in (0, 3):
            assert np.array_equal(acc[f], oracle.gridfit_eval(ocfg, frames[f], poses[f], want_acc=True)[2])


def test_stream_sweep_equals_gather_sweep(capi, ctx, cfg, oracle, ocfg):
    """Two independent kernels for board_fitting_cost.h:74-262 — k_sweep (a warp per target gathers the rows of its discs) and
    k_sweep_stream (a CTA per range of rows behind a TMA ring, candidates per lattice cell) — must leave the SAME accumulator
    tables, bit for bit: small poses, poses where the kd-tree neighbour test matters (centimetres), a frame whose intensities
    cannot carry the riding count (all-fp64 rows), ragged sizes down to a handful of points; and the oracle on a few of them."""
    rng = np.random.default_rng(78)
    sizes = [150000, 90000, 40, 20000, 3]
    frames = [synth.board_frame_cloud(max(n, 64), 700 + i, pert=(0.004 * i, -0.003 * i, 0.3 * i), pattern="rosette" if i % 2 else "raster")[:n]
              for i, n in enumerate(sizes)]
    frames[3] = frames[3].copy()
    frames[3][::7, 3] = 300.0 + rng.uniform(0, 50, len(frames[3][::7]))          # beyond [2^-3, 256): this frame takes the exact rows
    tables = {}
    for mode in ("gather", "stream"):
        b = capi.Batch(ctx, cfg, len(sizes), max(sizes))
        b.set_stream_sweep(mode == "stream")
        for f, p in enumerate(frames):
            b.upload(f, p)
        b.bin()
        out = []
        prng = np.random.default_rng(6)
        for trial in range(12):
            scale = [1e-4, 2e-3, 0.012, 0.035][trial % 4]
            poses = np.stack([prng.uniform(-scale, scale, len(sizes)), prng.uniform(-scale, scale, len(sizes)),
                              prng.uniform(-100 * scale, 100 * scale, len(sizes))], axis=1)
            b.sweep(poses)
            out.append((poses, b.read_eval(want_acc=True)[2].copy()))
        tables[mode] = out
        b.close()
    for (p0, a0), (p1, a1) in zip(tables["gather"], tables["stream"]):
        assert np.array_equal(p0, p1)
        assert np.array_equal(a0, a1), np.abs(a0 - a1).max()
    for poses, acc in tables["stream"][:4]:
        for f in (0, 3):
            assert np.array_equal(acc[f], oracle.gridfit_eval(ocfg, frames[f], poses[f], want_acc=True)[2])


def test_stream_sweep_solve_equals_gather_solve(capi, ctx, cfg):
    """The full solve (fused LM steps, speculative rounds, in-place moves) through either kernel: identical reports and corners."""
    frames = [synth.board_frame_cloud(60000 + 7000 * i, 300 + i, pert=(0.003 * i, -0.002 * i, 0.25 * i), pattern="rosette" if i % 2 else "raster")
              for i in range(5)]
    res = {}
    for mode in ("gather", "stream"):
        b = capi.Batch(ctx, cfg, len(frames), max(len(p) for p in frames))
        b.set_stream_sweep(mode == "stream")
        for f, p in enumerate(frames):
            b.upload(f, p)
        res[mode] = b.gridfit_solve()
        b.close()
    (c0, r0), (c1, r1) = res["gather"], res["stream"]
    assert np.array_equal(c0, c1)
    for a, bb in zip(r0, r1):
        assert (a.status, a.outer_iterations, a.pose_evals, a.lm_iterations) == (bb.status, bb.outer_iterations, bb.pose_evals, bb.lm_iterations)
        assert a.last_final_cost == bb.last_final_cost


def test_large_board_is_served_by_the_gather_sweep_only(capi, ctx, cfg, oracle, ocfg):
    """A 16 x 12 board (356 targets): the streaming kernel's per-target table no longer fits three CTAs per SM, so the batch
    refuses to switch to it (RCLC_ERR_UNSUPPORTED) and the gather kernel serves the board — bit-exact against the oracle. On a board
    half way (12 x 8, 172 targets) both kernels run and agree."""
    for W, H, stream_ok in ((16, 12, False), (12, 8, True)):
        cw, ow = copy.copy(cfg), copy.copy(ocfg)
        cw.board_width = W; cw.board_height = H; ow.board_width = W; ow.board_height = H
        pts = synth.board_frame_cloud(200000, 9, W=W, H=H, pert=(0.004, -0.003, 0.25))
        pose = np.array([[0.002, -0.001, 0.15]])
        b = capi.Batch(ctx, cw, 1, len(pts))
        b.upload(0, pts); b.bin(); b.sweep(pose)
        acc = b.read_eval(want_acc=True)[2][0].copy()
        assert np.array_equal(acc, oracle.gridfit_eval(ow, pts, pose[0], want_acc=True)[2])
        if stream_ok:
            b.set_stream_sweep(True)
            b.sweep(pose)
            assert np.array_equal(b.read_eval(want_acc=True)[2][0], acc)
        else:
            with pytest.raises(capi.RclcError):
                b.set_stream_sweep(True)
        b.close()


def test_gridfit_points_beyond_the_lattice_box_come_into_reach(ctx, cfg, oracle, ocfg):
    """The reference rebuilds its kd-tree on the FULL cloud in every outer iteration (opt_grid_fitting.h:75-76, 107): points
    that start beyond the lattice box (board holder, surroundings in the board's plane) become neighbours of the outer targets
    once the accumulated motion exceeds the 2 cm between the box and the neighbour radius. They are kept in the border cells
    of the sort, so the solve sees them exactly when the oracle does: same accumulators, same LM path, same corners."""
    rng = np.random.default_rng(41)
    x0, x1, y0, y1 = synth.board_extent(8, 6, 0.1)
    for pert in ((0.034, -0.031, 0.9), (-0.04, 0.035, -1.2)):
        board = synth.board_frame_cloud(120000, 77, pert=pert, pattern="rosette")
        # a frame of surroundings on the plane, out to 30 cm beyond the board, mid reflectance
        m = 60000
        u = rng.uniform(x0 - 0.3, x1 + 0.3, m); v = rng.uniform(y0 - 0.3, y1 + 0.3, m)
        keep = (u < x0) | (u > x1) | (v < y0) | (v > y1)
        th = np.deg2rad(pert[2]); c, s = np.cos(th), np.sin(th)
        ring = np.stack([c * u + s * v - pert[0], -s * u + c * v - pert[1], rng.normal(0, 0.002, m), rng.uniform(30, 90, m)], axis=1)[keep]
        pts = np.concatenate([board, ring.astype(np.float32)])[rng.permutation(len(board) + int(keep.sum()))]
        # accumulators at a pose that pulls outside points into the outer discs
        for pose in ((0.0, 0.0, 0.0), (pert[0], pert[1], pert[2]), (0.05, -0.045, 1.5)):
            _, _, a1 = ctx.gridfit_eval(cfg, pts, pose, want_acc=True)
            _, _, a0 = oracle.gridfit_eval(ocfg, pts, pose, want_acc=True)
            assert np.array_equal(a0, a1), pose
        c1, r1 = ctx.gridfit_solve(cfg, pts)
        rc, c0, r0, _ = oracle.gridfit_solve(ocfg, pts)
        assert rc == 0 and r1.status == 0
        assert r1.pose_evals == r0.pose_evals and r1.outer_iterations == r0.outer_iterations
        assert np.abs(c1 - c0).max() < 1e-6
        assert r1.outer_iterations >= 2                    # the cloud was moved at least once: the motion is what brings them in


def test_comm_world_of_one(capi, ctx, cfg, oracle):
    """The library's NCCL layer (csrc/comm.cu) on a communicator of one rank: the sharded extrinsic solve returns the bits of
    rclc_pose_refine, the accumulator all-reduce leaves a sweep's tables as they are. (Two ranks: scripts/multi_gpu_check.py.)"""
    comm = capi.Comm(ctx, capi.Comm.unique_id(), 0, 1)
    try:
        pc, img, T0, _ = _pose_problem(oracle, 20)
        rc0, Ta, a0, a1, ita = ctx.pose_refine(cfg, pc, img, T0)
        rc1, Tb, b0, b1, itb, ft = comm.pose_refine_sharded(cfg, pc, img, T0)
        assert ft == 20 and (rc0, Ta.tobytes(), a0, a1, ita) == (rc1, Tb.tobytes(), b0, b1, itb)
        pts = synth.board_frame_cloud(60000, 5, pert=(0.004, -0.003, 0.4))
        b = capi.Batch(ctx, cfg, 1, len(pts))
        b.upload(0, pts); b.bin(); b.sweep(np.array([[0.002, 0.001, 0.2]]))
        r0, j0, t0 = b.read_eval(want_acc=True)
        b.sweep(np.array([[0.002, 0.001, 0.2]])); comm.allreduce_acc(b)
        r1, j1, t1 = b.read_eval(want_acc=True)
        assert np.array_equal(t0, t1) and np.array_equal(r0, r1)
        assert comm.collectives() == 3
        b.close()
    finally:
        comm.close()
