"""Per-task stage clocks of k_sweep (build with RCLC_EXTRA_NVCC_FLAGS=-DRCLC_SWEEP_TIMING python rclc_b200/build.py --force)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
cfg = capi.config(); ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000)
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames): b.upload(f, p)
b.bin()
poses = np.array(perts)
for _ in range(4):
    ctx.flush_l2(); b.sweep(poses)
ctx.synchronize()
print("sweep ms", ctx.last_sweep_ms())
out = np.zeros((32 * 128 * 8, 12), np.int64)
capi.lib().rclc_batch_debug_read_terms.argtypes = [C.c_void_p, C.c_void_p]
rc = capi.lib().rclc_batch_debug_read_terms(b.h, out.ctypes.data_as(C.c_void_p))
assert rc == 0
o = out[out[:, 0] != 0].astype(np.float64)
t0 = o[:, 0].min()
beg, end = (o[:, 0] - t0) / 1e3, (o[:, 1] - t0) / 1e3
clk = 1.965e3   # cycles per us (approx)
def q(x): return "min %.1f p10 %.1f med %.1f p90 %.1f max %.1f sum %.0f" % (x.min(), np.percentile(x, 10), np.median(x), np.percentile(x, 90), x.max(), x.sum())
print("tasks", len(o), "span us", end.max())
busy = o[o[:, 7] > 0]
print("tasks with rows", len(busy))
print("begin us  ", q(beg)); print("end us    ", q(end)); print("dur us    ", q(end - beg))
for name, col in (("setup us", 3), ("stream us", 4), ("fold+atomics us", 5)):
    print("%-16s" % name, q(busy[:, col] / clk))
print("rows of cell    ", q(busy[:, 6])); print("row visits      ", q(busy[:, 7])); print("passes          ", q(busy[:, 8]))
print("cyc per row visit in stream", q(busy[:, 4] / np.maximum(busy[:, 7], 1)))
sm = o[:, 11].astype(int)
per_sm = np.bincount(sm, minlength=148)
print("tasks per SM", q(per_sm.astype(float)))
np.save("gpurun_out/sweep_timeline.npy", out)
sys.stdout.flush(); os._exit(0)
