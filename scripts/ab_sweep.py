"""A/B of the two sweep kernels on a resident batch: k_sweep_stream (default) against k_sweep (gather): accumulator tables must
agree bit for bit; prints the CUDA-event time of each (L2 flushed before every launch)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
cfg = capi.config()
ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000)
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames):
    b.upload(f, p)
b.bin()
poses = np.array(perts) * scale
out = {}
for name, stream in (("stream", True), ("gather", False)):
    b.set_stream_sweep(stream)
    ms = []
    for _ in range(reps):
        ctx.flush_l2()
        b.sweep(poses)
        ctx.synchronize()
        ms.append(ctx.last_sweep_ms())
    res, jac, acc = b.read_eval(want_acc=True)
    out[name] = acc
    print("%s us: %s median %.1f" % (name, " ".join("%.1f" % (m * 1e3) for m in ms), np.median(ms[2:]) * 1e3), flush=True)
a, g = out["stream"], out["gather"]
same = np.array_equal(a.view(np.uint64), g.view(np.uint64))
print("acc bit-identical:", same)
if not same:
    bad = np.argwhere(a != g)
    print("mismatches:", len(bad), "of", a.size, "first:", bad[:10].tolist())
    for idx in bad[:10]:
        print(tuple(idx), a[tuple(idx)], g[tuple(idx)])
    sys.exit(1)
print("ok", b.total_points())
