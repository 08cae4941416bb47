"""Speculative LM rounds of the batched grid-fit solve: per-frame evaluations consumed / swept, rounds, and the step time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
cfg = capi.config()
ctx = capi.Context(0)
if len(sys.argv) > 3:
    ctx.set_solve_groups(int(sys.argv[3]))
frames, perts = synth.gridfit_batch(F, N)
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames):
    b.upload(f, p)
ctx.synchronize()
for it in range(4):
    ctx.timer_start()
    c, reps = b.gridfit_solve()
    ms = ctx.timer_stop_ms()
    t0 = time.perf_counter(); b.gridfit_solve(); wall = (time.perf_counter() - t0) * 1e3
    print("solve %d: %.2f ms (events), %.2f ms (wall)" % (it, ms, wall))
ev = np.array([r.pose_evals for r in reps]); sw = np.array([r.poses_swept for r in reps]); ro = np.array([r.rounds for r in reps])
print("pose_evals", ev.tolist())
print("poses_swept", sw.tolist())
print("rounds", ro.tolist())
print("outer", [r.outer_iterations for r in reps])
print("spec_mismatch", sum(r.spec_mismatch for r in reps), "status", [r.status for r in reps])
print("totals: evals %d swept %d (x%.2f) max rounds %d (max evals %d)" % (ev.sum(), sw.sum(), sw.sum() / ev.sum(), ro.max(), ev.max()))
