import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
from oracle import pyoracle as O
cfg, ocfg = capi.config(), O.config()
ctx = capi.Context(0)
frames = [synth.board_frame_cloud(150000, 900 + i, pert=(0.004 * i, -0.003 * i, 0.3 * i), pattern="rosette" if i % 2 else "raster") for i in range(4)]
tables = {}
for mode in ("0", "1"):
    b = capi.Batch(ctx, cfg, 4, 150000)
    b.set_exact_sweep(mode == "1")
    for f, p in enumerate(frames): b.upload(f, p)
    b.bin()
    prng = np.random.default_rng(5); out = []
    for trial in range(12):
        scale = [1e-4, 2e-3, 0.03][trial % 3]
        poses = np.stack([prng.uniform(-scale, scale, 4), prng.uniform(-scale, scale, 4), prng.uniform(-130 * scale, 130 * scale, 4)], axis=1)
        b.sweep(poses); out.append((poses, b.read_eval(want_acc=True)[2].copy()))
    tables[mode] = out; b.close()
xy, isrow = O.generate_targets(ocfg)
for trial, ((p0, a0), (p1, a1)) in enumerate(zip(tables["0"], tables["1"])):
    d = np.argwhere(a0 != a1)
    if len(d) == 0: continue
    print("trial", trial, "n diffs", len(d))
    for f in sorted(set(d[:, 0])):
        ao = O.gridfit_eval(ocfg, frames[f], p0[f], want_acc=True)[2]
        print("  frame", f, "pose", p0[f], "fast==oracle", np.array_equal(a0[f], ao), "exact==oracle", np.array_equal(a1[f], ao))
        for (ff, t, j) in d[d[:, 0] == f][:8]:
            print("    target", t, xy[t], "acc", j, "fast", a0[ff, t, j], "exact", a1[ff, t, j], "oracle", ao[t, j])
