"""The extrinsic solve against the oracle, bit for bit, with the wall time of both launch shapes (lab script; the test of record is
tests/test_gpu_parity.py::test_pose_refine_is_the_oracles_bit_for_bit)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from rclc_b200 import capi
from oracle import pyoracle as O
from tests.test_gpu_parity import _pose_problem
ctx = capi.Context(0); cfg = capi.config(); ocfg = O.config()
g = np.load('tests/golden/pose.npz')
probs = [('golden', g['pc'], g['im'], g['T0'])]
for F, seed in ((5, 9), (20, 9), (64, 3), (256, 9), (300, 4), (700, 5), (1700, 6)):
    pc, img, T0, _ = _pose_problem(O, F, seed=seed); probs.append((f'F={F}', pc, img, T0))
for name, pc, im, T0 in probs:
    r0, j0 = O.pose_eval(pc, im, T0); r1, j1 = ctx.pose_eval(pc, im, T0)
    a = O.pose_refine(ocfg, pc, im, T0)
    line = f"{name}: eval bits equal {r0.tobytes() == r1.tobytes() and j0.tobytes() == j1.tobytes()}; oracle {a[4]} its cost {a[3]:.17g};"
    for shape in (1, 2):
        ctx.set_pose_test_hooks(shape=shape)
        ctx.pose_refine(cfg, pc, im, T0)
        t0 = time.perf_counter(); b = ctx.pose_refine(cfg, pc, im, T0); ms = (time.perf_counter() - t0) * 1e3
        same = b[1].tobytes() == a[1].tobytes() and b[2] == a[2] and b[3] == a[3] and b[4] == a[4]
        line += f" shape {shape}: {b[4]} its, {ms:.2f} ms, identical to oracle {same}, dt {np.abs(b[1][4:]-a[1][4:]).max():.2e};"
    ctx.set_pose_test_hooks()
    print(line, flush=True)
