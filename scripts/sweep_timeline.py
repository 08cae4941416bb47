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
nt = cfg.n_targets()
parts = int(os.environ.get("RCLC_SWEEP_PARTS", "1"))
out = np.zeros((32 * 128 * 8, 12), np.int64)
capi.lib().rclc_batch_debug_read_terms.argtypes = [C.c_void_p, C.c_void_p]
rc = capi.lib().rclc_batch_debug_read_terms(b.h, out.ctypes.data_as(C.c_void_p))
assert rc == 0
out = out[:F * nt * parts]
o = out.astype(np.float64)
nt = nt * parts
t0 = o[:, 0].min()
beg, end = (o[:, 0] - t0) / 1e3, (o[:, 1] - t0) / 1e3
clk = 1.965e3   # cycles per us (approx)
def q(x): return "min %.1f p10 %.1f med %.1f p90 %.1f max %.1f" % (x.min(), np.percentile(x, 10), np.median(x), np.percentile(x, 90), x.max())
print("tasks", len(o), "span us", end.max())
print("begin us  ", q(beg)); print("end us    ", q(end)); print("dur us    ", q(end - beg))
print("dur cyc/us", q(o[:, 2] / clk))
print("list us   ", q(o[:, 3] / clk)); print("stream us ", q(o[:, 4] / clk))
print("prologue (consts, smem init) us", q(o[:, 5] / clk)); print("cell look-up + prefix us     ", q(o[:, 7] / clk))
print("final pump + fold us         ", q(o[:, 8] / clk)); print("epilogue (totals out) us     ", q(o[:, 10] / clk))
print("rows      ", q(o[:, 6]))
print("mode counts", np.bincount(o[:, 9].astype(int)))
print("cyc/row in stream", q(o[:, 4] / np.maximum(o[:, 6], 1)))
print("other us (dur - list - stream)", q((o[:, 2] - o[:, 3] - o[:, 4]) / clk))
sm = o[:, 11].astype(int)
per_sm = np.bincount(sm, minlength=148)
print("tasks per SM", q(per_sm.astype(float)))
# slowest tasks
idx = np.argsort(end)[-8:]
for i in idx: print("late task f=%d t=%d beg %.1f end %.1f rows %d list %.1f stream %.1f sm %d (%d tasks)" % (i // nt, i % nt, beg[i], end[i], o[i, 6], o[i, 3] / clk, o[i, 4] / clk, sm[i], per_sm[sm[i]]))
for n_on in sorted(set(per_sm)): print("SMs with %d tasks: %d, median end %.1f" % (n_on, (per_sm == n_on).sum(), np.median(end[per_sm[sm] == n_on]) if (per_sm[sm] == n_on).any() else 0))
np.save("gpurun_out/sweep_timeline.npy", out)
sys.stdout.flush(); os._exit(0)
