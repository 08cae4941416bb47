"""One extrinsic solve (20 frames x 35 corners, cluster shape) for `ncu -k regex:k_pose_refine -c 1`."""
import sys
sys.path.insert(0, '.')
from rclc_b200 import capi
from oracle import pyoracle as O
from tests.test_gpu_parity import _pose_problem
ctx = capi.Context(0); cfg = capi.config()
pc, img, T0, _ = _pose_problem(O, 20, seed=9)
rc, T, c0, c1, it = ctx.pose_refine(cfg, pc, img, T0)
print(rc, it, c0, c1)
