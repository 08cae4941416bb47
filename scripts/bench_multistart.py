"""BASELINE configs[2]: multi-start grid-pose search, 4096 hypotheses x 200 k board points of one frame (points resident and
sorted; the timed call uploads the poses, sweeps all of them and reads the 4096 costs back). One JSON line."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = capi.config(); ctx = capi.Context(0)
pts = synth.board_frame_cloud(N, 77, pert=(0.004, -0.003, 0.3), pattern="rosette")
poses = synth.multistart_poses()              # 16 x 16 x 16 lattice: +-3 cm, +-4 deg
b = capi.Batch(ctx, cfg, 1, N)
b.upload(0, pts); b.bin()
cost = b.multistart(0, poses); ctx.synchronize()
t = []
for _ in range(reps):
    l0 = ctx.kernel_launches()
    t0 = time.perf_counter(); cost = b.multistart(0, poses); ctx.synchronize(); t.append(time.perf_counter() - t0)
    launches = ctx.kernel_launches() - l0
ms = float(np.median(t)) * 1e3
k = int(np.argmin(cost))
print(json.dumps({"config": "configs[2]: multi-start grid-pose search", "hypotheses": len(poses), "points": N, "ms_per_search": ms,
                  "Mpts_poses_per_s": len(poses) * N / ms / 1e3, "kernel_launches": int(launches), "best_pose": [float(x) for x in poses[k]],
                  "best_cost": float(cost[k]), "true_misalignment": [0.004, -0.003, 0.3]}))
