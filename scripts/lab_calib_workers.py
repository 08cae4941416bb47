"""Wall time of rclc_calibrate_frames (20 scans x 100 k) against the number of host workers."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from rclc_b200 import capi, synth
cfg = capi.config(); ctx = capi.Context(0)
scans, imgs = zip(*[synth.raw_scan(f)[:2] for f in range(20)])
host = [torch.from_numpy(s).pin_memory().numpy() for s in scans]
imgs = np.stack(imgs)
print("host cores", os.cpu_count())
for w in (0, 8, 16, 20, 24, 32, 40):
    capi._chk(capi.lib().rclc_ctx_set_calib_workers(ctx.h, w))
    for _ in range(3): out = ctx.calibrate_frames(cfg, host, imgs)
    t, st = [], []
    for _ in range(8):
        t0 = time.perf_counter(); out = ctx.calibrate_frames(cfg, host, imgs); t.append(time.perf_counter() - t0); st.append(out["stage_ms"].copy())
    s = np.median(np.stack(st), axis=0)
    print("workers %2d: total %.2f ms (chains %.2f, grid fit %.2f, extrinsic %.2f)" % (w, 1e3 * np.median(t), s[0], s[1], s[2]))
