"""Time of the counting sort (rclc_batch_bin: init + count + scatter + row boxes) of 20 x 500 k points on one stream."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from rclc_b200 import capi, synth
F, N = 20, 500000
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000, pattern="raster")
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames): b.upload(f, p)
for _ in range(3): b.bin()
ctx.synchronize()
t = []
for _ in range(10):
    ctx.synchronize(); ctx.timer_start(); b.bin(); t.append(ctx.timer_stop_ms())
print("bin (sort) of 20 x 500k: median %.3f ms  (%s)" % (np.median(t), " ".join("%.3f" % x for x in t)))
