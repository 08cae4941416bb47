"""Timing experiments for k_sweep: batch size (L2-resident or not), flush or not, pose magnitude."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
cfg = capi.config()
ctx = capi.Context(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
for F in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["5","20","40"])]:
    frames, perts = synth.gridfit_batch(F, N, seed0=1000)
    b = capi.Batch(ctx, cfg, F, N)
    for f, p in enumerate(frames):
        b.upload(f, p)
    b.bin()
    for name, poses in (("perts", np.array(perts)), ("zero", np.zeros((F, 3))), ("small", np.full((F, 3), 1e-3))):
        for flush in (True, False):
            for _ in range(3):
                b.sweep(poses)
            ctx.synchronize()
            d = []; dk = []
            for _ in range(15):
                if flush:
                    ctx.flush_l2()
                ctx.synchronize()
                ctx.timer_start()
                b.sweep(poses)
                d.append(ctx.timer_stop_ms())
                dk.append(ctx.last_sweep_ms())
            ms = float(np.median(d)); mk = float(np.median(dk))
            print(f"F={F:3d} N={N} pose={name:6s} flush={int(flush)}  call {ms*1e3:7.1f} us  kernel {mk*1e3:7.1f} us  {F*N*16/mk/1e6:8.1f} GB/s  {1e3*mk/F:6.2f} us/frame", flush=True)
    b.close()
