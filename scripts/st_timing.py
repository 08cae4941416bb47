"""Per-warp stage cycles of k_sweep_stream (lab build: RCLC_EXTRA_NVCC_FLAGS=-DRCLC_ST_TIMING python rclc_b200/build.py --force)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cfg = capi.config()
ctx = capi.Context(0)
frames, perts = synth.gridfit_batch(F, N, seed0=1000)
b = capi.Batch(ctx, cfg, F, N)
for f, p in enumerate(frames):
    b.upload(f, p)
b.bin()
b.set_stream_sweep(True)
poses = np.array(perts) * scale
for _ in range(4):
    ctx.flush_l2()
    b.sweep(poses)
    ctx.synchronize()
    print("k_sweep_stream us: %.1f" % (ctx.last_sweep_ms() * 1e3))
nw = 4096
out = np.zeros((nw, 8), np.int64)
capi._chk(capi.lib().rclc_lab_st_timing(out.ctypes.data_as(C.c_void_p), nw))
out = out[out[:, 5] > 0]
names = ["prologue", "tile waits", "next_segment", "flush", "segments(rows+waits+flush)", "total", "n segments", "n tiles"]
print(len(out), "warps")
for k, nm in enumerate(names):
    col = out[:, k]
    print("%-28s median %9.0f  mean %9.0f  max %9.0f" % (nm, np.median(col), col.mean(), col.max()))
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "st_timing.npy"), out)
