"""How sensitive is the reference's per-frame chain (est_board_corner, include/pc_process.h:583-646) to the last digits of an
upstream stage? Oracle only, CPU.

pc_plane_fitting returns four doubles; pc_plane_prj rounds the projected points to float; everything downstream (voxel winners,
block clustering, the float cloud moves and the accept / reject decisions of the grid fit's LM with its heuristic Jacobian,
include/opt/board_fitting_cost.h:224-260) is discrete. The test perturbs the fitted plane's d by 1e-9 relative — the agreement an
independent implementation of the plane fit reaches when it sums in another order or solves the step by Cholesky instead of QR
(this repository's round-1 kernel did exactly that) — and measures what the ORACLE ITSELF does with it: a few hundred of the
~34 k projected points move by one float ulp, and in about half the cases the grid fit's LM takes another accept / reject path
and the corners come out 0.1 - 0.5 mm elsewhere. One ulp of the double moves no float at all.

So the 0.1 mm bar of the full chain is not attainable by agreeing "to 1e-9" upstream: it takes the SAME BITS upstream. The
product therefore reproduces the oracle's plane bit for bit (exact order-free sums, one defined arithmetic:
tests/test_gpu_parity.py::test_plane_fit_matches_oracle), and the reference's own Ceres / Eigen build would show the same 0.1 - 0.5 mm
scatter against any other build of itself."""
import numpy as np
import pytest

from oracle import pyoracle as O
from rclc_b200 import synth


def _chain_after_plane(cfg, board, plane):
    """est_board_corner from the projection on: pc_plane_prj, GMM, eular_cluster_iter, calc_laser_coor_v1, transform, grid fit."""
    prj = O.plane_project(board, plane)
    mu, sg, _ = O.gmm_fit_1d(prj[:, 3], [15, 100], [15, 15])
    rc, cen = O.eular_cluster_iter(cfg, prj, mu[0] + 2.0 * sg[0])
    assert rc == 0
    rc, Tbl = O.calc_laser_coor(cfg, cen, plane)
    assert rc == 0
    moved = O.transform_cloud(prj, Tbl.astype(np.float32))
    rc, corners, rep, _ = O.gridfit_solve(cfg, moved, Tbl)
    assert rc == 0
    return corners, rep


def test_plane_perturbation_of_1e9_forks_the_grid_fit_one_ulp_does_not():
    cfg = O.config()
    worst, moved_1e9, moved_ulp, forks, cases = 0.0, 0, 0, 0, 0
    for f in range(4):
        raw, _, _ = synth.raw_scan(f)
        rc, board = O.segment_board(cfg, raw)
        assert rc == 0
        board = np.c_[board[:, :3] @ synth.T_BASE[:3, :3].T, board[:, 3]].astype(np.float32)
        rc, plane, _, _ = O.plane_fit(cfg, board)
        assert rc == 0
        # the split chain is the chain: same corners as orc_est_board_corner
        c0, r0 = _chain_after_plane(cfg, board, plane)
        rc, c_ref, _, _ = O.est_board_corner(cfg, board)
        assert rc == 0 and np.array_equal(c0, c_ref)
        base = O.plane_project(board, plane)[:, :3]
        p_ulp = plane.copy()
        p_ulp[3] = np.nextafter(p_ulp[3], np.inf)
        moved_ulp += int((O.plane_project(board, p_ulp)[:, :3] != base).any(axis=1).sum())
        for rel in (1e-9, -1e-9):
            p1 = plane.copy()
            p1[3] *= 1.0 + rel
            moved_1e9 += int((O.plane_project(board, p1)[:, :3] != base).any(axis=1).sum())
            c1, r1 = _chain_after_plane(cfg, board, p1)
            worst = max(worst, float(np.abs(c1 - c0).max()))
            forks += int(r1.pose_evals != r0.pose_evals)
            cases += 1
    print(f"plane d * (1 +- 1e-9): {moved_1e9 / cases:.0f} projected points per frame move by a float ulp, {forks} of {cases} grid fits "
          f"take another LM path, corners move by up to {worst * 1e3:.3f} mm; one ulp of d: {moved_ulp} points move")
    assert moved_ulp <= 2                        # one ulp of the double does not reach the float coordinates
    assert moved_1e9 > 100 * cases               # 1e-9 does: a few hundred points per frame
    assert forks >= 2 and worst > 1e-4           # ... and the oracle's own grid fit answers with > 0.1 mm
