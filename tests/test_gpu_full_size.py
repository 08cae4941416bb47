"""Parity at BASELINE.json's FULL sizes, where the oracle is too slow to check everything: size-independent properties of the
domain (additivity over disjoint point sets, invariance under point order, fast == all-fp64 sweep, determinism of the grouped
solve) plus oracle spot checks on a few frames / hypotheses, all through the C ABI on the GPU."""
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


@pytest.fixture(scope="module")
def batch20(capi, ctx, cfg):
    """configs[1]: 20 raster frames x 500 k points, resident and sorted."""
    frames, perts = synth.gridfit_batch(20, 500000, seed0=1000)
    b = capi.Batch(ctx, cfg, 20, 500000)
    for f, p in enumerate(frames):
        b.upload(f, p)
    b.bin()
    yield b, frames, np.array(perts)
    b.close()


def _tables(capi, ctx, cfg, clouds, poses, exact=False):
    b = capi.Batch(ctx, cfg, len(clouds), max(len(c) for c in clouds))
    b.set_exact_sweep(exact)
    for f, p in enumerate(clouds):
        b.upload(f, p)
    b.bin()
    b.sweep(np.asarray(poses, np.float64))
    acc = b.read_eval(want_acc=True)[2].copy()
    b.close()
    return acc


def test_accumulators_are_additive_and_order_free_at_500k(capi, ctx, cfg, batch20):
    """Every accumulator is an exact sum over points: the table of a cloud equals the sum of the tables of any two disjoint
    halves, and does not depend on the order of the points — bit for bit, at 500 k points and the frame's full misalignment."""
    b, frames, perts = batch20
    b.sweep(perts)
    whole = b.read_eval(want_acc=True)[2].copy()
    rng = np.random.default_rng(3)
    for f in (0, 7, 19):
        p = frames[f]
        perm = rng.permutation(len(p))
        half = len(p) // 2 + 12345
        acc = _tables(capi, ctx, cfg, [p[perm], p[perm[:half]], p[perm[half:]]], [perts[f]] * 3)
        assert np.array_equal(acc[0], whole[f])                       # order of the points
        assert np.array_equal(acc[1] + acc[2], whole[f])              # disjoint halves
        assert acc[1][:, 4:8].sum() + acc[2][:, 4:8].sum() == whole[f][:, 4:8].sum() > 0


def test_fast_sweep_equals_all_fp64_sweep_at_full_size(capi, ctx, cfg, batch20, monkeypatch):
    b, frames, perts = batch20
    b.sweep(perts)
    fast = b.read_eval(want_acc=True)[2].copy()
    exact = _tables(capi, ctx, cfg, frames[:6], perts[:6], exact=True)
    assert np.array_equal(exact, fast[:6])


def test_oracle_spot_check_at_full_size(capi, ctx, cfg, oracle, ocfg, batch20):
    b, frames, perts = batch20
    b.sweep(perts)
    res, jac, acc = b.read_eval(want_acc=True)
    for f in (3, 11):
        r0, j0, a0 = oracle.gridfit_eval(ocfg, frames[f], perts[f], want_acc=True)
        assert np.array_equal(acc[f], a0)
        np.testing.assert_allclose(res[f], r0, rtol=1e-12, atol=1e-15)


def test_solve_is_deterministic_and_independent_of_grouping(capi, ctx, cfg, oracle, ocfg, batch20, monkeypatch):
    """The 20-frame solve twice, then one frame on its own: identical corners and trajectories; one frame against the oracle."""
    b, frames, perts = batch20
    c0, r0 = b.gridfit_solve()
    c1, r1 = b.gridfit_solve()
    assert np.array_equal(c0, c1) and [r.pose_evals for r in r0] == [r.pose_evals for r in r1]
    cs, rs = ctx.gridfit_solve(cfg, frames[5])
    assert np.array_equal(cs, c0[5]) and rs.pose_evals == r0[5].pose_evals
    rc_o, co, ro, _ = oracle.gridfit_solve(ocfg, frames[5])
    assert rc_o == 0
    assert ro.pose_evals == r0[5].pose_evals and ro.outer_iterations == r0[5].outer_iterations
    np.testing.assert_allclose(c0[5], co, atol=1e-6)
    assert all(r.status == 0 for r in r0)


def test_all_20_frames_of_configs1_follow_the_oracles_trajectory(capi, ctx, cfg, oracle, ocfg, batch20):
    """configs[1] in full against the oracle: every one of the 20 x 500 k frames — same number of pose evaluations, LM iterations
    and outer iterations, corners within 1 um. (The oracle runs a thread per frame: a few seconds on the box's host cores.)"""
    from concurrent.futures import ThreadPoolExecutor
    b, frames, perts = batch20
    corners, reps = b.gridfit_solve()
    with ThreadPoolExecutor(max_workers=20) as ex:
        ref = list(ex.map(lambda p: oracle.gridfit_solve(ocfg, p), frames))
    for f, (rc_o, co, ro, _) in enumerate(ref):
        assert rc_o == 0 and reps[f].status == 0
        assert (ro.pose_evals, ro.lm_iterations, ro.outer_iterations) == (reps[f].pose_evals, reps[f].lm_iterations, reps[f].outer_iterations), f
        assert np.abs(corners[f] - co).max() <= 2e-15, (f, np.abs(corners[f] - co).max())      # the oracle's to an ulp
    # and the accumulator tables of all 20 frames at their full misalignment, bit for bit
    b.bin()
    b.sweep(perts)
    acc = b.read_eval(want_acc=True)[2]
    with ThreadPoolExecutor(max_workers=20) as ex:
        ref_acc = list(ex.map(lambda fp: oracle.gridfit_eval(ocfg, fp[0], fp[1], want_acc=True)[2], zip(frames, perts)))
    for f in range(20):
        assert np.array_equal(acc[f], ref_acc[f]), f


def test_multistart_4096_hypotheses_200k_points(capi, ctx, cfg, oracle, ocfg):
    """configs[2]: all 4096 costs finite, the oracle on a sample of hypotheses, and the minimum at the lattice point nearest the
    cloud's true misalignment."""
    true = (0.004, -0.003, 0.3)
    pts = synth.board_frame_cloud(200000, 77, pert=true, pattern="rosette")
    P = synth.multistart_poses()
    assert len(P) == 4096
    cost = ctx.gridfit_multistart(cfg, pts, P)
    assert np.isfinite(cost).all()
    rng = np.random.default_rng(0)
    from concurrent.futures import ThreadPoolExecutor
    ks = np.r_[rng.choice(len(P), 64, replace=False), int(np.argmin(cost))]
    with ThreadPoolExecutor(max_workers=16) as ex:
        ref = list(ex.map(lambda k: oracle.gridfit_cost(ocfg, pts, P[k]), ks))
    np.testing.assert_allclose(cost[ks], ref, rtol=1e-12)
    best = P[int(np.argmin(cost))]
    assert abs(best[0] - true[0]) <= 0.004 and abs(best[1] - true[1]) <= 0.004 and abs(best[2] - true[2]) <= 0.54


def test_projection_of_10M_points_is_the_oracle_bit_for_bit(ctx, cfg, oracle, ocfg):
    """configs[4]."""
    pts = synth.frustum_cloud(10_000_000, 5)
    img = synth.noise_image(6, cfg.cam_width, cfg.cam_height)
    T = np.eye(4, dtype=np.float32); T[0, 3] = 0.05; T[1, 3] = -0.02; T[0, 1] = 0.01; T[1, 0] = -0.01
    rgb1, pix1, fov1 = ctx.project_colorize(cfg, pts, T, img)
    rgb0, pix0, fov0 = oracle.project_colorize(ocfg, pts, T, img)
    assert np.array_equal(fov1, fov0) and 0.5 < fov0.mean() < 0.8
    k = fov0.astype(bool)
    assert np.array_equal(pix1[k], pix0[k]) and np.array_equal(rgb1[k], rgb0[k])


def test_gmm_on_20_frames_of_500k(capi, ctx, cfg, oracle, batch20):
    b, frames, perts = batch20
    mu, sg, pri = b.gmm_fit()
    for f in (0, 13):
        m0, s0, p0 = oracle.gmm_fit_1d(frames[f][:, 3], [15, 100], [15, 15])
        np.testing.assert_allclose(mu[f], m0, rtol=1e-10); np.testing.assert_allclose(sg[f], s0, rtol=1e-9); np.testing.assert_allclose(pri[f], p0, rtol=1e-10)
