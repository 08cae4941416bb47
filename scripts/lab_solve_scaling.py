"""How the solve's time splits into latency and throughput: the slowest frame of the bench alone, then F copies of it."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from rclc_b200 import capi, synth
N = 500000
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(20, N, seed0=1000, pattern="raster")
slow, fast = frames[2], frames[1]
for F in (1, 2, 5, 10, 20, 40):
    b = capi.Batch(ctx, cfg, F, N)
    for f in range(F): b.upload(f, slow)
    for _ in range(2): corners, reps = b.gridfit_solve()
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); corners, reps = b.gridfit_solve(); t.append(time.perf_counter() - t0)
    ev = sum(r.pose_evals for r in reps); sw = sum(r.poses_swept for r in reps)
    ms = 1e3 * np.median(t)
    print("F=%2d copies of the slowest frame: %.2f ms, %d rounds, %d poses swept, %.1f GB of point-poses -> %.2f TB/s" % (F, ms, reps[0].rounds, sw, sw * N * 16 / 1e9, sw * N * 16 / ms / 1e9))
    b.close()
