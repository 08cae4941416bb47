"""Wall time of the single-frame chains and of rclc_calibrate_frames at several worker counts (GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from rclc_b200 import capi, synth

cfg = capi.config()
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
with capi.Context(0) as ctx:
    raw, img, _ = synth.raw_scan(0)
    for _ in range(3):
        board = ctx.segment_board(cfg, raw)
    bc = np.c_[board[:, :3] @ synth.T_BASE[:3, :3].T, board[:, 3]].astype(np.float32)
    for _ in range(3):
        ctx.est_board_corner(cfg, bc)
    if mode == "ncu":
        ctx.segment_board(cfg, raw); ctx.est_board_corner(cfg, bc)
        sys.exit(0)
    t = []
    for _ in range(10):
        l0 = ctx.kernel_launches(); t0 = time.perf_counter(); ctx.segment_board(cfg, raw); t.append(time.perf_counter() - t0); l1 = ctx.kernel_launches()
    print(f"segment_board: {1e3 * np.median(t):.2f} ms, {l1 - l0} launches")
    t = []
    for _ in range(10):
        l0 = ctx.kernel_launches(); t0 = time.perf_counter(); ctx.est_board_corner(cfg, bc); t.append(time.perf_counter() - t0); l1 = ctx.kernel_launches()
    print(f"est_board_corner: {1e3 * np.median(t):.2f} ms, {l1 - l0} launches")
    scans, imgs = zip(*[synth.raw_scan(f)[:2] for f in range(20)])
    import torch
    pinned = [torch.from_numpy(s).pin_memory() for s in scans]
    host = [p.numpy() for p in pinned]
    for W in (1, 2, 4, 8, 16, 20):
        ctx.set_calib_workers(W)
        for _ in range(2):
            ctx.calibrate_frames(cfg, host, np.stack(imgs))
        ms = []
        for _ in range(5):
            out = ctx.calibrate_frames(cfg, host, np.stack(imgs)); ms.append(out["stage_ms"].copy())
        print(f"workers {W}: chains / grid fit / extrinsic / total = {np.round(np.median(np.stack(ms), axis=0), 2)} ms")
