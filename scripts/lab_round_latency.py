"""Lab build (-DRCLC_LAB): the slowest frame of the bench alone — per-round latency and the LM step's share of it (stderr lines of
rclc_batch_gridfit_solve), then an ncu-free launch-level view: wall time per solve at 1 frame."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from rclc_b200 import capi, synth
N = 500000
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(20, N, seed0=1000, pattern="raster")
b = capi.Batch(ctx, cfg, 1, N)
b.upload(0, frames[2])
for _ in range(3): corners, reps = b.gridfit_solve()
t0 = time.perf_counter(); corners, reps = b.gridfit_solve(); print("solve ms %.2f rounds %d" % (1e3 * (time.perf_counter() - t0), reps[0].rounds))
