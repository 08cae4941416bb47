"""One pose refine of 256 frames x 35 corners (cluster shape, then one CTA) - the target of the ncu capture in profiles/."""
import os, sys, time
sys.path.insert(0, '.')
from rclc_b200 import capi
from oracle import pyoracle as oracle            # problem generator only (quat_to_R)
from tests.test_gpu_parity import _pose_problem
ctx = capi.Context(0); cfg = capi.config()
pc, img, T0, _ = _pose_problem(oracle, 256, seed=9)
for mode in ("2", "0"):
    ctx.set_pose_test_hooks(shape=2 if mode == "2" else 1)
    t0 = time.perf_counter()
    rc, T, c0, c1, it = ctx.pose_refine(cfg, pc, img, T0)
    print("cluster" if mode == "2" else "one CTA", rc, it, f"{(time.perf_counter() - t0) * 1e3:.2f} ms")
