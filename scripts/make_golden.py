#!/usr/bin/env python
"""Writes the golden fixtures under tests/golden/.

What they are — and are not. The reference (zhijianglu/RCLC) cannot be built or imported in this image (PCL, Ceres,
Eigen and OpenCV are absent) and ships no tests or data for the hot path, so these vectors are NOT
outputs of the reference. They are outputs of the CPU oracle (oracle/, a line-by-line restatement of the reference) on
seeded synthetic inputs, frozen so that (a) the oracle cannot drift silently (tests/test_golden_fixtures.py, CPU) and
(b) the CUDA path is checked against fixed numbers as well as against the live oracle (same file, -m gpu).
The only known answers the reference itself holds (apps/calc_rmse.cpp literals) are pinned in tests/test_oracle_golden.py
and stored here as calc_rmse_kat.json for completeness.

    python scripts/make_golden.py        # regenerates tests/golden/*.npz, *.json
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O          # noqa: E402
from rclc_b200 import synth              # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def cases():
    """name -> dict(inputs regenerated from seeds, compact expected outputs)."""
    cfg = O.config()
    out = {}

    # K4: BOARD_FITTING_COST::Evaluate x 82 at three poses (accumulators are integers / exact sums: bit-exact)
    pts = synth.board_frame_cloud(20000, 11, pert=(0.004, -0.003, 0.25), pattern="rosette")
    ev = {}
    for k, pose in enumerate([(0.0, 0.0, 0.0), (0.002, 0.001, 0.2), (-0.011, 0.008, -0.9)]):
        r, j, a = O.gridfit_eval(cfg, pts, pose, want_acc=True)
        ev[f"pose{k}"] = np.array(pose); ev[f"res{k}"] = r; ev[f"jac{k}"] = j; ev[f"acc{k}"] = a
    out["gridfit_eval"] = dict(seed=11, n=20000, pert=(0.004, -0.003, 0.25), pattern="rosette", input_sha=sha(pts), **ev)

    # opt_grid_fitting: corners + solver trajectory counters
    sol = {}
    for k, (seed, pert, pat) in enumerate([(21, (0.006, -0.004, 0.3), "raster"), (22, (-0.009, 0.007, -0.6), "rosette")]):
        p = synth.board_frame_cloud(30000, seed, pert=pert, pattern=pat)
        rc, corners, rep, Tbl = O.gridfit_solve(cfg, p)
        sol[f"corners{k}"] = corners
        sol[f"counters{k}"] = np.array([rep.outer_iterations, rep.pose_evals, rep.lm_iterations, rep.status], np.int64)
        sol[f"costs{k}"] = np.array([rep.first_initial_cost, rep.last_final_cost])
        sol[f"input_sha{k}"] = np.frombuffer(sha(p).encode(), np.uint8)
    out["gridfit_solve"] = sol

    # GMMFit::fit_1d
    inten = synth.board_frame_cloud(50000, 31)[:, 3]
    mu, sg, pri = O.gmm_fit_1d(inten, [15, 100], [15, 15])
    out["gmm"] = dict(mu=mu, sigma=sg, priors=pri, input_sha=sha(inten))

    # RANSAC scoring / selection on a plane with outliers
    rng = np.random.default_rng(41)
    n = 6000
    yz = rng.uniform(-1, 1, (n, 2))
    x = np.where(rng.random(n) < 0.7, 4.0 + 0.2 * yz[:, 0] - 0.1 * yz[:, 1] + rng.normal(0, 0.01, n), 4.0 + 2 * rng.uniform(-1, 1, n))
    cloud = np.stack([x, yz[:, 0], yz[:, 1], np.full(n, 50.0)], 1).astype(np.float32)
    tri = O.ransac_draw_triplets(n, 64)
    planes, valid, counts = O.ransac_score(cloud, tri, 0.08)
    win, iters = O.ransac_replay(counts, valid, n)
    mask = O.plane_select(cloud, planes[win], 0.08)
    out["ransac"] = dict(seed=41, triplets=tri, planes=planes, valid=valid, counts=counts, winner=np.array([win, iters]),
                         mask_packed=np.packbits(mask.astype(bool)), input_sha=sha(cloud))

    # pose_reprj_opt
    rng = np.random.default_rng(51)
    F, Cn = 8, 35
    q = np.array([0.02, -0.01, 0.015, 0.9996]); q /= np.linalg.norm(q); t = np.array([0.05, -0.03, 0.02])
    x_, y_, z_, w_ = q
    R = np.array([[1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_)],
                  [2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_)],
                  [2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)]])
    pc = np.zeros((F, Cn, 3)); im = np.zeros((F, Cn, 2))
    for f in range(F):
        c0 = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.4, 0.4), rng.uniform(3.5, 5.5)])
        for k in range(Cn):
            p = c0 + np.array([0.1 * (k % 7), 0.1 * (k // 7), 0.02 * (k % 5)])
            P = R @ p + t
            pc[f, k] = p; im[f, k] = P[:2] / P[2] + rng.normal(0, 1e-4, 2)
    T0 = np.array([0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0])
    res, jac = O.pose_eval(pc, im, T0)
    out["pose"] = dict(pc=pc, im=im, T0=T0, res0=res, jac0=jac)
    Tout = O.pose_refine(cfg, pc, im, T0)
    out["pose"]["Tcl"] = np.asarray(Tout[1]); out["pose"]["costs"] = np.array([Tout[2], Tout[3]])

    # projection / colourise
    pts = synth.frustum_cloud(8000, 61)
    img = synth.noise_image(62)
    T = np.eye(4, dtype=np.float32); T[0, 3] = 0.05; T[1, 3] = -0.02; T[0, 1] = 0.01; T[1, 0] = -0.01
    rgb, pix, fov = O.project_colorize(cfg, pts, T, img)
    out["project"] = dict(seed=61, img_seed=62, T=T, pix=pix.astype(np.int32), infov_packed=np.packbits(fov.astype(bool)),
                          rgb_sha=np.frombuffer(sha(rgb).encode(), np.uint8), input_sha=sha(pts))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, d in cases().items():
        arrs = {}
        for k, v in d.items():
            if isinstance(v, str):
                arrs[k] = np.frombuffer(v.encode(), np.uint8)
            else:
                arrs[k] = np.asarray(v)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print(name, {k: tuple(np.shape(v)) for k, v in arrs.items()})
    kat = {
        "source": "apps/calc_rmse.cpp:70-74, 113-121 (inputs); :142-164 (expected, printed with 6 significant digits)",
        "gt_euler_mrad": [5.02947, -1.14864, -0.0513812],
        "indoor": {"T_eul": [4.92869, -1.52763, -0.0462151], "dt_mm": [2.54194, -1.2659, 4.73766], "dr_mrad": [-0.100777, -0.378992, 0.00516618]},
        "outdoor": {"T_eul": [5.28472, -1.09601, -0.19652], "dt_mm": [2.26847, 0.0585949, 5.54309], "dr_mrad": [0.255252, 0.0526271, -0.145139]},
    }
    json.dump(kat, open(os.path.join(OUT, "calc_rmse_kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
