"""Per-task stage clocks of k_sweep (lab build: RCLC_EXTRA_NVCC_FLAGS=-DRCLC_SWEEP_TIMING python rclc_b200/build.py --force)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
flush = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000)
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames): b.upload(f, p)
b.bin()
poses = np.array(perts)
for _ in range(4):
    if flush: ctx.flush_l2()
    b.sweep(poses); ctx.synchronize()
print("F", F, "flush", flush, "sweep us %.1f" % (ctx.last_sweep_ms() * 1e3))
out = np.zeros((8192, 8), np.int64)
capi._chk(capi.lib().rclc_lab_sweep_timing(out.ctypes.data_as(C.c_void_p), 8192))
o = out[out[:, 0] != 0].astype(np.float64)
clk = 1965.0
names = ["control+pose+constants", "cell table", "box tests (+ interleaved streaming)", "last list streamed", "reduce + store"]
print(len(o), "tasks; per-task stage times in us (median / p90 / max)")
for k, nm in enumerate(names):
    d = (o[:, k + 1] - o[:, k]) / clk
    d = d[o[:, k + 1] > 0]
    print("  %-40s %6.2f %6.2f %6.2f" % (nm, np.median(d), np.percentile(d, 90), d.max()))
tot = (o[:, 5] - o[:, 0]) / clk
g0 = o[:, 7].min()
print("  %-40s %6.2f %6.2f %6.2f" % ("whole task", np.median(tot), np.percentile(tot, 90), tot.max()))
print("  task start after the first task's start (globaltimer, us): median %.2f max %.2f" % (np.median(o[:, 7] - g0) / 1e3, (o[:, 7] - g0).max() / 1e3))
