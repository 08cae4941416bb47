import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from rclc_b200 import capi, synth
F, N = 20, 500000
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000)
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames): b.upload(f, p)
b.bin()
poses = np.array(perts)
for flush in (True, False, True, False):
    ms = []
    for _ in range(8):
        if flush: ctx.flush_l2()
        b.sweep(poses); ctx.synchronize(); ms.append(ctx.last_sweep_ms())
    print("flush" if flush else "no flush", " ".join("%.1f" % (m*1e3) for m in ms))
# fewer frames: L2-resident
for Fs in (1, 2, 5, 10):
    b2 = capi.Batch(ctx, cfg, Fs, N)
    for f in range(Fs): b2.upload(f, frames[f])
    b2.bin()
    for flush in (True, False):
        ms = []
        for _ in range(6):
            if flush: ctx.flush_l2()
            b2.sweep(poses[:Fs]); ctx.synchronize(); ms.append(ctx.last_sweep_ms())
        print(Fs, "frames", "flush" if flush else "no flush", " ".join("%.1f" % (m*1e3) for m in ms))
    b2.close()
