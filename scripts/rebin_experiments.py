import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
cfg = capi.config(); ctx = capi.Context(0)
N = 500000
for F in (1, 2, 20):
    frames, perts = synth.gridfit_batch(F, N, seed0=1000)
    b = capi.Batch(ctx, cfg, F, N)
    for f, p in enumerate(frames): b.upload(f, p)
    for _ in range(3): b.bin()
    ctx.synchronize()
    d = []
    for _ in range(10):
        ctx.timer_start(); b.bin(); d.append(ctx.timer_stop_ms())
    print(f"F={F} bin() (copy + init + count + scatter + rowbox): {np.median(d)*1e3:.1f} us", flush=True)
    b.close()
