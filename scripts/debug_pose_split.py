import os, sys, numpy as np
sys.path.insert(0, '.')
from rclc_b200 import capi
from oracle import pyoracle as oracle
from tests.test_gpu_parity import _pose_problem
ctx = capi.Context(0); cfg = capi.config()
for F, seed in ((5, 9), (20, 9), (256, 9), (300, 4)):
    pc, img, T0, _ = _pose_problem(oracle, F, seed=seed)
    out = {}
    for name, cl, dbg in (("one", "0", "0"), ("cluster", "2", "0"), ("cluster+block_dev", "2", "1")):
        ctx.set_pose_test_hooks(shape={'0': 1, '2': 2}[cl], cluster_debug=int(dbg))
        rc, T, c0, c1, it = ctx.pose_refine(cfg, pc, img, T0)
        out[name] = (rc, T.tobytes(), c0, c1, it)
        print(F, name, rc, it, repr(c0), repr(c1), T[4:])
    print('   cluster==one', out['cluster'] == out['one'], ' cluster+block_dev==one', out['cluster+block_dev'] == out['one'])
