"""How much of the 160 MB upload is hidden behind the solve? (A) uploads alone, (B) solve alone, (C) uploads + solve."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from rclc_b200 import capi, synth
F, N = 20, 500000
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000, pattern="raster")
pinned = [torch.from_numpy(p).pin_memory() for p in frames]
frames = [t.numpy() for t in pinned]
b = capi.Batch(ctx, cfg, F, N)
def up():
    for f, p in enumerate(frames): b.upload(f, p)
def med(fn, n=10):
    t = []
    for _ in range(n):
        ctx.synchronize(); t0 = time.perf_counter(); fn(); ctx.synchronize(); t.append(time.perf_counter() - t0)
    return 1e3 * np.median(t)
up(); b.gridfit_solve(); b.gridfit_solve()
print("A uploads alone        %.2f ms" % med(up))
print("B solve alone          %.2f ms" % med(lambda: b.gridfit_solve()))
print("C uploads + solve      %.2f ms" % med(lambda: (up(), b.gridfit_solve())))
order = list(range(F))
def up_rev():
    for f in reversed(order): b.upload(f, frames[f])
print("C' reversed upload order + solve %.2f ms" % med(lambda: (up_rev(), b.gridfit_solve())))
for G in (5, 20):
    capi._chk(capi.lib().rclc_ctx_set_solve_groups(ctx.h, G))
    up(); b.gridfit_solve()
    print("groups %d: B %.2f  C %.2f" % (G, med(lambda: b.gridfit_solve()), med(lambda: (up(), b.gridfit_solve()))))
