"""Minimal driver for ncu: bin a resident batch once, then run a few single-pose sweeps (k_sweep)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = capi.config()
ctx = capi.Context(0)
F0 = min(F, 20)                      # (larger batches repeat the 20 frames: every copy has its own memory)
frames, perts = synth.gridfit_batch(F0, N, seed0=1000)
frames = [frames[f % F0] for f in range(F)]; perts = [perts[f % F0] for f in range(F)]
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames):
    b.upload(f, p)
b.bin()
poses = np.array(perts)
ms = []
for _ in range(reps):
    ctx.flush_l2()
    b.sweep(poses)
    ctx.synchronize()
    ms.append(ctx.last_sweep_ms())
print("k_sweep us:", " ".join("%.1f" % (m * 1e3) for m in ms), "median %.1f" % (np.median(ms[2:]) * 1e3))
print("ok", b.total_points())
