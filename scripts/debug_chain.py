"""Stage-by-stage comparison of the per-frame chain (est_board_corner) between the library's operators and the CPU oracle on the
raw-scan rig: prints the first stage whose output differs in any bit. GPU box only."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import pyoracle as O
from rclc_b200 import capi, synth


def eular_iter_host(cfg, ctx, prj, thd):
    """pc_process.h:104-211 over the library's operators (the product's own shim is C++: include/rclc_ops.hpp)."""
    idx = ctx.uniform_sample(prj, cfg.block_us_radius)
    blocks = prj[idx]
    return blocks


def main():
    cfg, ocfg = capi.config(), O.config()
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    with capi.Context(0) as ctx:
        for f in range(nf):
            raw, _, _ = synth.raw_scan(f)
            rc, board = O.segment_board(ocfg, raw)
            board = np.c_[board[:, :3] @ synth.T_BASE[:3, :3].T, board[:, 3]].astype(np.float32)
            rc0, p0, it0, _ = O.plane_fit(ocfg, board)
            rc1, p1, it1, _ = ctx.plane_fit(cfg, board)
            msg = [f"frame {f}: n {len(board)}"]
            msg.append(f"plane {'==' if np.array_equal(p0, p1) else '!= ' + str(p1 - p0)} its {it0}/{it1}")
            prj0 = O.plane_project(board, p0); prj1 = ctx.plane_project(board, p0)
            msg.append(f"project {'==' if np.array_equal(prj0, prj1) else '!='}")
            m0, s0, _ = O.gmm_fit_1d(prj0[:, 3], [15, 100], [15, 15]); m1, s1, _ = ctx.gmm_fit_1d(prj0[:, 3], [15, 100], [15, 15])
            t0, t1 = m0[0] + 2.0 * s0[0], m1[0] + 2.0 * s1[0]
            msg.append(f"gmm thd rel diff {abs(t1 - t0) / t0:.1e} float-equal {np.float32(t0) == np.float32(t1)}")
            i0 = O.uniform_sample(prj0, ocfg.block_us_radius); i1 = ctx.uniform_sample(prj0, cfg.block_us_radius)
            msg.append(f"us {'==' if np.array_equal(i0, i1) else '!='} ({len(i0)})")
            rc, cen = O.eular_cluster_iter(ocfg, prj0, t0)
            rcT0, T0 = O.calc_laser_coor(ocfg, cen, p0); T1 = capi.calc_laser_coor(cfg, cen, p0)
            T1 = T1[1] if isinstance(T1, tuple) else T1
            msg.append(f"Tbl max diff {np.abs(T0 - T1).max():.1e} float-equal {np.array_equal(T0.astype(np.float32), T1.astype(np.float32))}")
            mv0 = O.transform_cloud(prj0, T0.astype(np.float32)); mv1 = ctx.transform_cloud(prj0, T0.astype(np.float32))
            msg.append(f"move {'==' if np.array_equal(mv0, mv1) else '!='}")
            rcs, c0, r0, _ = O.gridfit_solve(ocfg, mv0, T0); c1, r1 = ctx.gridfit_solve(cfg, mv0, T0)
            msg.append(f"gridfit evals {r0.pose_evals}/{r1.pose_evals} outer {r0.outer_iterations}/{r1.outer_iterations} status {r1.status} corners d {np.abs(c1 - c0).max():.2e}")
            print("; ".join(msg), flush=True)


if __name__ == "__main__":
    main()
