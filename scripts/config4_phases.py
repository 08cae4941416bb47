"""Wall-clock phases of BASELINE configs[3] on one GPU (256 frames x 100 k points): create / upload / GMM / solve / close / pose refine."""
import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from rclc_b200 import capi, synth
cfg = capi.config(); ctx = capi.Context(0)
T_gt = synth.gt_extrinsic()
scene = [synth.calib_frame(f, 100000, T_cl_gt=T_gt) for f in range(256)]
frames = [s[0] for s in scene]; T_bl = np.stack([s[1] for s in scene]); img = np.stack([s[3] for s in scene])
def rot_to_quat_xyzw(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
Tcl0 = np.r_[rot_to_quat_xyzw(synth._rot_xyz(0.012, -0.015, 0.01) @ T_gt[:3, :3]), T_gt[:3, 3] + [0.02, -0.015, 0.01]]
for rep in range(5):
    t = [time.perf_counter()]
    batch = capi.Batch(ctx, cfg, 256, 100000); ctx.synchronize(); t.append(time.perf_counter())
    for f, p in enumerate(frames): batch.upload(f, p)
    ctx.synchronize(); t.append(time.perf_counter())
    batch.gmm_fit(); t.append(time.perf_counter())
    corners, reports = batch.gridfit_solve(T_bl); t.append(time.perf_counter())
    batch.close(); t.append(time.perf_counter())
    rc, Tcl, c0, c1, it = ctx.pose_refine(cfg, corners, img, Tcl0); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("create %.1f upload %.1f gmm %.1f solve %.1f close %.1f pose %.1f  total %.1f ms" % (*d, d.sum()), flush=True)
