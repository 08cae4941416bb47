"""Multi-GPU host logic of the calibration path: one process per GPU, frames sharded by rank, ONE exchange.

Every stage up to the 35 board corners of a frame is independent per frame (the reference already runs frames in an OpenMP
`parallel for`, apps/Calibrate.cpp:90-92), so frame f goes to rank f mod world and nothing is exchanged until the final
extrinsic solve (pose_reprj_opt, include/opt/opt_reprj_calib.h:21-66), which needs every frame's corners. Of the two
exchanges §8(e) weighs — an all-reduce of the 6x6 normal equations per LM iteration, or one all-gather of the corners
(35 x 3 doubles per frame, 215 KB at 256 frames) followed by a replicated solve — this module does the second: one
collective instead of ~1000, and every rank ends with the bit-identical extrinsic because the solve is deterministic.
`torch.distributed` is plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard(n_frames, rank, world):
    """Global frame indices owned by `rank` (round-robin: f mod world == rank)."""
    return list(range(rank, n_frames, world))


def allgather_frames(local, n_frames, rank, world, device=None):
    """All-gather per-frame arrays. `local` has shape (len(shard(...)), ...) in shard order; returns (n_frames, ...) in
    global frame order on every rank. Ranks may own different numbers of frames (ragged): shards are padded to the longest."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, np.float64)
    tail = local.shape[1:]
    if world == 1:
        assert local.shape[0] == n_frames
        return local.copy()
    longest = -(-n_frames // world)
    buf = torch.zeros((longest,) + tail, dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(local).to(buf.device)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    out = np.zeros((n_frames,) + tail)
    for r in range(world):
        idx = shard(n_frames, r, world)
        if idx:
            out[idx] = gathered[r][:len(idx)].cpu().numpy()
    return out


def calibrate_sharded(ctx, cfg, frames_local, img_norm_local, n_frames, rank, world, Tcl0, device=None, T_bl_local=None):
    """Stage 2 of the reference on a shard (GMM EM + grid fit of the local frames on this rank's GPU), the one exchange, and
    the replicated extrinsic refinement. Returns (Tcl, corners_all, reports_local, (initial_cost, final_cost, iterations))."""
    from . import capi
    F = len(frames_local)
    if F:
        batch = capi.Batch(ctx, cfg, F, max(len(p) for p in frames_local))
        for f, p in enumerate(frames_local):
            batch.upload(f, p)
        batch.gmm_fit()
        corners_local, reports = batch.gridfit_solve(T_bl_local)
        batch.close()
    else:
        corners_local, reports = np.zeros((0, cfg.n_corners(), 3)), []
    corners_all = allgather_frames(corners_local, n_frames, rank, world, device)
    img_all = allgather_frames(np.asarray(img_norm_local, np.float64).reshape(F, cfg.n_corners(), 2), n_frames, rank, world, device)
    rc, Tcl, c0, c1, it = ctx.pose_refine(cfg, corners_all, img_all, Tcl0)
    return Tcl, corners_all, reports, (c0, c1, it)
