"""Rounds, poses swept and poses consumed per frame of the 20 x 500 k solve; wall time per solve."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from rclc_b200 import capi, synth
F, N = 20, 500000
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000, pattern="raster")
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames): b.upload(f, p)
for _ in range(3): corners, reps = b.gridfit_solve()
t = []
for _ in range(10):
    t0 = time.perf_counter(); corners, reps = b.gridfit_solve(); t.append(time.perf_counter() - t0)
print("solve ms: median %.2f min %.2f" % (1e3 * np.median(t), 1e3 * min(t)))
ev = np.array([r.pose_evals for r in reps]); sw = np.array([r.poses_swept for r in reps]); rd = np.array([r.rounds for r in reps])
print("pose_evals", ev.tolist()); print("poses_swept", sw.tolist()); print("rounds", rd.tolist())
print("sum evals %d, sum swept %d (%.2f x), max rounds %d, mean rounds %.1f" % (ev.sum(), sw.sum(), sw.sum() / ev.sum(), rd.max(), rd.mean()))
for G in (1, 2, 4, 5, 10, 20):
    capi._chk(capi.lib().rclc_ctx_set_solve_groups(ctx.h, G))
    for _ in range(2): b.gridfit_solve()
    t = []
    for _ in range(6):
        t0 = time.perf_counter(); b.gridfit_solve(); t.append(time.perf_counter() - t0)
    print("groups %2d: solve ms median %.2f" % (G, 1e3 * np.median(t)))
