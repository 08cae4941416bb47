"""One whole calibration (20 raw scans: stage 1 + stage 2 + extrinsic) for a launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/profile_calibration.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from rclc_b200 import capi, synth

cfg = capi.config()
with capi.Context(0) as ctx:
    scans, imgs = zip(*[synth.raw_scan(f)[:2] for f in range(20)])
    ctx.set_calib_workers(1)                       # one frame after the other: the list reads in chain order
    out = ctx.calibrate_frames(cfg, list(scans), np.stack(imgs))
    print(out["status"].tolist(), out["pose_iterations"], ctx.kernel_launches())
