"""Per-iteration cycle breakdown of the pose refine (build with RCLC_EXTRA_NVCC_FLAGS=-DRCLC_POSE_TIMING)."""
import os, sys
sys.path.insert(0, '.')
from rclc_b200 import capi
from oracle import pyoracle as oracle
from tests.test_gpu_parity import _pose_problem
ctx = capi.Context(0); cfg = capi.config()
for F, seed in ((20, 9), (64, 3), (256, 9)):
    pc, img, T0, _ = _pose_problem(oracle, F, seed=seed)
    for cl in ("0", "2"):
        ctx.set_pose_test_hooks(shape=2 if cl == "2" else 1)
        ctx.pose_refine(cfg, pc, img, T0)
