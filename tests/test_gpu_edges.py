"""Edge cases of the smaller operators against the oracle, exactly: one-point and coarse / fine voxel grids, clusters that the size
limits reject, a plane almost without a normal, triplets with repeated indices and collinear points, points in and behind the
camera plane and NaN points in the projection. (The reference holds no tests; these are the inputs its code paths branch on.)"""
import numpy as np
import pytest

from rclc_b200 import synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


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
def cloud():
    rng = np.random.default_rng(7)
    return np.c_[rng.uniform(-1, 1, (3000, 3)) + [0, 0, 4], rng.uniform(0, 150, 3000)].astype(np.float32)


@pytest.mark.parametrize("leaf", [0.01, 0.5, 100.0])
def test_uniform_sample_leaf_sizes(ctx, oracle, cloud, leaf):
    assert all(_same(x, y) for x, y in zip(oracle.uniform_sample(cloud, leaf), ctx.uniform_sample(cloud, leaf)))
    assert all(_same(x, y) for x, y in zip(oracle.uniform_sample(cloud[:1], leaf), ctx.uniform_sample(cloud[:1], leaf)))


@pytest.mark.parametrize("tol,lo,hi", [(0.03, 1, 100000), (0.5, 5, 100000), (10.0, 1, 100), (0.03, 500, 2500000)])
def test_euclidean_cluster_size_limits(ctx, oracle, cloud, tol, lo, hi):
    assert all(_same(x, y) for x, y in zip(oracle.euclidean_cluster(cloud, tol, lo, hi), ctx.euclidean_cluster(cloud, tol, lo, hi)))
    assert all(_same(x, y) for x, y in zip(oracle.euclidean_cluster(cloud[:1], tol, 1, 10), ctx.euclidean_cluster(cloud[:1], tol, 1, 10)))


@pytest.mark.parametrize("plane", [[0, 0, 1, -4.0], [0.3, -0.2, 0.9, -3.0], [1e-9, 0, 1e-9, 1.0]])
def test_plane_project_planes(ctx, oracle, cloud, plane):
    assert _same(oracle.plane_project(cloud, plane), ctx.plane_project(cloud, plane))


def test_ransac_score_degenerate_triplets(ctx, oracle, cloud):
    tri = np.array([[0, 1, 2], [5, 5, 9], [10, 11, 10], [100, 200, 300]], np.int32)           # repeated indices: no plane / a NaN plane
    for a, b in zip(oracle.ransac_score(cloud, tri, 0.08), ctx.ransac_plane_score(cloud, tri, 0.08)):
        assert np.array_equal(np.asarray(a, np.float64), np.asarray(b, np.float64), equal_nan=True)
    line = np.c_[np.linspace(0, 1, 50), np.zeros(50), np.full(50, 4.0), np.zeros(50)].astype(np.float32)
    tri = np.array([[0, 10, 20], [1, 2, 3]], np.int32)                                           # collinear points
    for a, b in zip(oracle.ransac_score(line, tri, 0.08), ctx.ransac_plane_score(line, tri, 0.08)):
        assert np.array_equal(np.asarray(a, np.float64), np.asarray(b, np.float64), equal_nan=True)


def test_projection_of_points_in_and_behind_the_camera_plane(ctx, cfg, oracle, ocfg):
    rng = np.random.default_rng(9)
    pts = np.c_[rng.uniform(-2, 2, (2000, 2)), rng.uniform(-1, 6, 2000), rng.uniform(0, 150, 2000)].astype(np.float32)
    pts[:5, 2] = 0.0
    pts[5:8, :3] = np.nan
    img = synth.noise_image(3)
    T = np.eye(4, dtype=np.float32)
    assert all(_same(x, y) for x, y in zip(oracle.project_colorize(ocfg, pts, T, img), ctx.project_colorize(cfg, pts, T, img)))


@pytest.mark.parametrize("which", ["one point", "ten points", "300 points", "far from the board"])
def test_grid_fit_on_clouds_that_cannot_be_fitted(ctx, cfg, oracle, ocfg, which):
    """Clouds that leave half-discs empty: NaN residuals, a Ceres FAILURE in every outer iteration (status 1), max_iter_time + 1 outer
    iterations, the corners of the pose the fit started from - the same on both sides, accumulators and NaN pattern included."""
    pts = synth.board_frame_cloud(40000, 3, pert=(0.004, -0.002, 0.2))
    p = {"one point": pts[:1], "ten points": pts[:10], "300 points": pts[::133], "far from the board": pts + np.array([5, 5, 0, 0], np.float32)}[which]
    rc, oc, orep, _ = oracle.gridfit_solve(ocfg, p)
    c, rep = ctx.gridfit_solve(cfg, p)
    assert rc == 0 and (rep.status, rep.outer_iterations, rep.pose_evals) == (orep.status, orep.outer_iterations, orep.pose_evals) and rep.status == 1
    assert np.abs(c - oc).max() <= 2e-15
    r0, j0, a0 = oracle.gridfit_eval(ocfg, p, (0.001, 0.002, 0.1), want_acc=True)
    r1, j1, a1 = ctx.gridfit_eval(cfg, p, (0.001, 0.002, 0.1), want_acc=True)
    assert np.array_equal(a0, a1) and np.array_equal(np.isnan(r0), np.isnan(r1))
    assert np.allclose(r0[~np.isnan(r0)], r1[~np.isnan(r1)], rtol=1e-10, atol=0)
