#!/usr/bin/env python
"""BASELINE config 4: batch calibration of N synthetic frames sharded over the GPUs of one box (one process per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        scripts/run_config4.py --frames 256 --pts 100000

Every rank grid-fits its own frames, the corners are all-gathered once (NCCL), every rank refines the same extrinsic.
Rank 0 prints one JSON line: error against the ground-truth extrinsic, device time (max over ranks), and whether all ranks
ended with the bit-identical result."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rot_to_quat_xyzw(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w); y = (R[0, 2] - R[2, 0]) / (4 * w); z = (R[1, 0] - R[0, 1]) / (4 * w)
    return np.array([x, y, z, w])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--pts", type=int, default=100000)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rclc_b200 import capi, multi, synth
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = capi.config(); ctx = capi.Context(local)
    T_gt = synth.gt_extrinsic()
    mine = multi.shard(a.frames, rank, world)
    scene = [synth.calib_frame(f, a.pts, T_cl_gt=T_gt) for f in mine]
    frames = [s[0] for s in scene]; T_bl = np.stack([s[1] for s in scene]) if scene else np.zeros((0, 4, 4))
    img = np.stack([s[3] for s in scene]) if scene else np.zeros((0, cfg.n_corners(), 2))
    # initial guess: ground truth perturbed by ~1 degree / 2 cm (the reference starts from solvePnPRansac, Calibrate.cpp:43-51)
    dR = synth._rot_xyz(0.012, -0.015, 0.01)
    q0 = rot_to_quat_xyzw(dR @ T_gt[:3, :3]); Tcl0 = np.r_[q0, T_gt[:3, 3] + [0.02, -0.015, 0.01]]
    ms_runs = []
    for _ in range(2):                   # the first call pays the allocations (device, pinned) and the CUDA-graph capture
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.timer_start()
        Tcl, corners_all, reports, (c0, c1, it) = multi.calibrate_sharded(ctx, cfg, frames, img, a.frames, rank, world, Tcl0,
                                                                          device=torch.device("cuda", local), T_bl_local=T_bl)
        ms_runs.append(ctx.timer_stop_ms())
    t = torch.tensor(ms_runs, dtype=torch.float64, device="cuda")
    same = torch.tensor(Tcl, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lo, hi = same.clone(), same.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool(torch.equal(lo, hi))
    else:
        identical = True
    if rank == 0:
        q_gt = rot_to_quat_xyzw(T_gt[:3, :3])
        q = Tcl[:4] / np.linalg.norm(Tcl[:4])
        ang = 2 * np.arcsin(min(1.0, np.linalg.norm(q * np.sign(q[3]) - q_gt * np.sign(q_gt[3])) / 2))
        corner_err = float(np.abs(corners_all[mine] - np.stack([s[2] for s in scene])).max()) if scene else None
        print(json.dumps({"config": "configs[3]: batch calibration, frames sharded round-robin, one all-gather of corners",
                          "frames": a.frames, "pts_per_frame": a.pts, "n_gpus": world, "ms_first_call": float(t[0].item()), "ms": float(t[1].item()),
                          "rot_err_rad_vs_gt": float(ang), "trans_err_m_vs_gt": float(np.abs(Tcl[4:] - T_gt[:3, 3]).max()),
                          "max_corner_err_m_rank0": corner_err, "pose_cost": [c0, c1], "pose_iterations": it,
                          "all_ranks_identical": identical}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
