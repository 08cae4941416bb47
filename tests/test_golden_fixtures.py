"""Frozen vectors under tests/golden/ (written by scripts/make_golden.py from the CPU oracle on seeded synthetic inputs —
the reference itself cannot be built here and holds no vectors for the hot path; see that script's header).
CPU: the oracle still reproduces them (it cannot drift silently). GPU: the CUDA path, through the C ABI, matches them —
bit-exact for accumulators / counts / masks / pixels, within the north-star tolerances for the floating-point outputs."""
import hashlib
import os

import numpy as np
import pytest

from rclc_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def text(a):
    return bytes(a).decode()


def eval_cloud(g):
    pts = synth.board_frame_cloud(int(g["n"]), int(g["seed"]), pert=tuple(g["pert"]), pattern=text(g["pattern"]))
    assert sha(pts) == text(g["input_sha"]), "synthetic generator drifted: regenerate the fixtures deliberately"
    return pts


SOLVE_CASES = [(21, (0.006, -0.004, 0.3), "raster"), (22, (-0.009, 0.007, -0.6), "rosette")]


def ransac_cloud():
    rng = np.random.default_rng(41)
    n = 6000
    yz = rng.uniform(-1, 1, (n, 2))
    x = np.where(rng.random(n) < 0.7, 4.0 + 0.2 * yz[:, 0] - 0.1 * yz[:, 1] + rng.normal(0, 0.01, n), 4.0 + 2 * rng.uniform(-1, 1, n))
    return np.stack([x, yz[:, 0], yz[:, 1], np.full(n, 50.0)], 1).astype(np.float32)


# ------------------------------------------------------------------------------------------------ CPU: oracle vs fixtures
def test_oracle_reproduces_gridfit_eval(oracle, ocfg):
    g = load("gridfit_eval")
    pts = eval_cloud(g)
    for k in range(3):
        r, j, a = oracle.gridfit_eval(ocfg, pts, tuple(g[f"pose{k}"]), want_acc=True)
        assert np.array_equal(a, g[f"acc{k}"])
        np.testing.assert_allclose(r, g[f"res{k}"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(j, g[f"jac{k}"], rtol=1e-12, atol=1e-12)


def test_oracle_reproduces_solve_gmm_ransac_pose_project(oracle, ocfg):
    g = load("gridfit_solve")
    for k, (seed, pert, pat) in enumerate(SOLVE_CASES):
        p = synth.board_frame_cloud(30000, seed, pert=pert, pattern=pat)
        assert sha(p) == text(g[f"input_sha{k}"])
        rc, corners, rep, _ = oracle.gridfit_solve(ocfg, p)
        np.testing.assert_allclose(corners, g[f"corners{k}"], atol=1e-12)
        assert [rep.outer_iterations, rep.pose_evals, rep.lm_iterations, rep.status] == list(g[f"counters{k}"])
    g = load("gmm")
    inten = synth.board_frame_cloud(50000, 31)[:, 3]
    mu, sg, pri = oracle.gmm_fit_1d(inten, [15, 100], [15, 15])
    np.testing.assert_allclose(np.r_[mu, sg, pri], np.r_[g["mu"], g["sigma"], g["priors"]], rtol=1e-13)
    g = load("ransac")
    cloud = ransac_cloud()
    assert sha(cloud) == text(g["input_sha"])
    assert np.array_equal(oracle.ransac_draw_triplets(len(cloud), 64), g["triplets"])
    planes, valid, counts = oracle.ransac_score(cloud, g["triplets"], 0.08)
    assert np.array_equal(counts, g["counts"]) and np.array_equal(valid, g["valid"]) and np.array_equal(planes, g["planes"])
    win, iters = oracle.ransac_replay(counts, valid, len(cloud))
    assert [win, iters] == list(g["winner"])
    g = load("pose")
    r, j = oracle.pose_eval(g["pc"], g["im"], g["T0"])
    np.testing.assert_allclose(r, g["res0"], rtol=1e-13); np.testing.assert_allclose(j, g["jac0"], rtol=1e-12, atol=1e-14)
    g = load("project")
    pts = synth.frustum_cloud(8000, int(g["seed"])); img = synth.noise_image(int(g["img_seed"]))
    rgb, pix, fov = oracle.project_colorize(ocfg, pts, g["T"], img)
    assert np.array_equal(pix, g["pix"]) and np.array_equal(np.packbits(fov.astype(bool)), g["infov_packed"]) and sha(rgb) == text(g["rgb_sha"])


def test_calc_rmse_kat_file_matches_the_test_literals():
    import json
    kat = json.load(open(os.path.join(GOLD, "calc_rmse_kat.json")))
    from tests.test_oracle_golden import CASES
    for name in ("indoor", "outdoor"):
        assert tuple(kat[name]["dt_mm"]) == CASES[name]["dt"] and tuple(kat[name]["T_eul"]) == CASES[name]["T_eul"]


# ------------------------------------------------------------------------------------------------ GPU: CUDA path vs fixtures
@pytest.fixture(scope="module")
def gpu():
    from rclc_b200 import capi
    ctx = capi.Context(0)
    yield capi, ctx, capi.config()
    ctx.close()


@pytest.mark.gpu
def test_cuda_gridfit_eval_matches_fixtures(gpu):
    capi, ctx, cfg = gpu
    g = load("gridfit_eval")
    pts = eval_cloud(g)
    for k in range(3):
        r, j, a = ctx.gridfit_eval(cfg, pts, tuple(g[f"pose{k}"]), want_acc=True)
        assert np.array_equal(a, g[f"acc{k}"])                                    # integer counts and exact sums
        np.testing.assert_allclose(r, g[f"res{k}"], rtol=1e-5)                    # north-star: 1e-5 relative (observed: ~1e-15)
        np.testing.assert_allclose(j, g[f"jac{k}"], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_cuda_solve_gmm_ransac_pose_project_match_fixtures(gpu):
    capi, ctx, cfg = gpu
    g = load("gridfit_solve")
    for k, (seed, pert, pat) in enumerate(SOLVE_CASES):
        p = synth.board_frame_cloud(30000, seed, pert=pert, pattern=pat)
        corners, rep = ctx.gridfit_solve(cfg, p)
        assert np.abs(corners - g[f"corners{k}"]).max() <= 2e-15                  # bar 0.1 mm; they are the oracle's to an ulp
        assert [rep.outer_iterations, rep.pose_evals] == list(g[f"counters{k}"][:2])
    g = load("gmm")
    inten = synth.board_frame_cloud(50000, 31)[:, 3]
    mu, sg, pri = ctx.gmm_fit_1d(inten, [15, 100], [15, 15])
    np.testing.assert_allclose(np.r_[mu, sg, pri], np.r_[g["mu"], g["sigma"], g["priors"]], rtol=1e-9)
    g = load("ransac")
    cloud = ransac_cloud()
    planes, valid, counts = ctx.ransac_plane_score(cloud, g["triplets"], 0.08)
    assert np.array_equal(counts, g["counts"]) and np.array_equal(valid, g["valid"]) and np.array_equal(planes, g["planes"])
    mask = ctx.ransac_plane_select(cloud, g["planes"][int(g["winner"][0])], 0.08)[0]
    assert np.array_equal(np.packbits(mask.astype(bool)), g["mask_packed"])
    g = load("pose")
    r, j = ctx.pose_eval(g["pc"], g["im"], g["T0"])
    np.testing.assert_allclose(r, g["res0"], rtol=1e-10); np.testing.assert_allclose(j, g["jac0"], rtol=1e-8, atol=1e-12)
    rc, Tcl, c0, c1, it = ctx.pose_refine(cfg, g["pc"], g["im"], g["T0"])
    assert Tcl.tobytes() == g["Tcl"].tobytes() and [c0, c1] == list(g["costs"])      # the oracle's solve bit for bit (pose.cu)
    g = load("project")
    pts = synth.frustum_cloud(8000, int(g["seed"])); img = synth.noise_image(int(g["img_seed"]))
    rgb, pix, fov = ctx.project_colorize(cfg, pts, g["T"], img)
    assert np.array_equal(pix, g["pix"]) and np.array_equal(np.packbits(fov.astype(bool)), g["infov_packed"]) and sha(rgb) == text(g["rgb_sha"])
