"""Two-or-more-GPU check of the library's NCCL layer (csrc/comm.cu). Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py

torch.distributed only carries the NCCL unique id to the other ranks (and compares results); every exchange below is the library's.
  (1) rclc_pose_refine_sharded with ragged shards == rclc_pose_refine on the concatenated frames, same bits on every rank
  (2) rclc_batch_allreduce_acc: one frame's points split across the ranks == the unsplit sweep, bit for bit
"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth          # noqa: E402
from oracle import pyoracle as O           # noqa: E402  (problem generator only)
from tests.test_gpu_parity import _pose_problem   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ids = [capi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    cfg = capi.config()
    ctx = capi.Context(local)
    comm = capi.Comm(ctx, ids[0], rank, world)
    ok = True
    # (1) ragged shards: rank r gets frames [cut[r], cut[r+1])
    F = 41
    pc, img, T0, _ = _pose_problem(O, F, seed=3)
    cut = [0] + sorted(np.random.default_rng(1).choice(np.arange(1, F), world - 1, replace=False).tolist()) + [F]
    lo, hi = cut[rank], cut[rank + 1]
    rc, T, c0, c1, it, ft = comm.pose_refine_sharded(cfg, pc[lo:hi], img[lo:hi], T0)
    rc1, T1, d0, d1, it1 = ctx.pose_refine(cfg, pc, img, T0)
    same = ft == F and (rc, T.tobytes(), c0, c1, it) == (rc1, T1.tobytes(), d0, d1, it1)
    got = [None] * world
    dist.all_gather_object(got, T.tobytes())
    same = same and all(g == got[0] for g in got)
    print(f"[rank {rank}] sharded extrinsic solve, frames {lo}..{hi} of {F}: {'same bits as the one-GPU solve on every rank' if same else 'MISMATCH'} ({it} iterations)", flush=True)
    ok = ok and same
    # (2) one frame split by points
    pts = synth.board_frame_cloud(400000, 11, pert=(0.006, -0.004, 0.5), pattern="rosette")
    pose = np.array([[0.003, -0.002, 0.3]])
    share = pts[rank::world]
    b = capi.Batch(ctx, cfg, 1, len(pts))
    b.upload(0, share); b.bin(); b.sweep(pose); comm.allreduce_acc(b)
    r, j, a = b.read_eval(want_acc=True)
    b.upload(0, pts); b.bin(); b.sweep(pose)
    r0, j0, a0 = b.read_eval(want_acc=True)
    same = np.array_equal(a, a0) and np.array_equal(r, r0) and np.array_equal(j, j0)
    print(f"[rank {rank}] point-split sweep ({len(share)} of {len(pts)} points here) + all-reduce of the accumulator tables: "
          f"{'== unsplit, bit for bit' if same else 'MISMATCH'}", flush=True)
    ok = ok and same
    print(f"[rank {rank}] NCCL collectives issued by the library: {comm.collectives()}", flush=True)
    b.close(); comm.close(); ctx.close()
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
