#!/bin/bash
# lab build of the library with per-stage timers of the calibration chains, then the timing script at a few worker counts
set -e
RCLC_EXTRA_NVCC_FLAGS=-DRCLC_LAB python rclc_b200/build.py --force > /dev/null
python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from rclc_b200 import capi, synth
cfg = capi.config()
with capi.Context(0) as ctx:
    scans, imgs = zip(*[synth.raw_scan(f)[:2] for f in range(20)])
    host = [torch.from_numpy(s).pin_memory().numpy() for s in scans]
    keep = [torch.from_numpy(s).pin_memory() for s in scans]; host = [k.numpy() for k in keep]
    for W in (1, 4, 16):
        ctx.set_calib_workers(W)
        for _ in range(4):
            out = ctx.calibrate_frames(cfg, host, np.stack(imgs))
PY
python rclc_b200/build.py --force > /dev/null
