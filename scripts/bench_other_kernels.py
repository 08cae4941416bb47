"""Kernel-level timings of the smaller kernels of the path: RANSAC scoring (K1), GMM EM (K3), pose refine (K6). Run once plain,
then under `ncu --metrics gpu__time_duration.sum` for the per-launch times (profiles/r01_other_kernels.md)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
cfg = capi.config(); ctx = capi.Context(0)
out = {}
# K1: 801 hypotheses (800 iterations + 1, apps/BoardSegmentation.cpp:71) x 300 k points
rng = np.random.default_rng(1)
n = 300000
pts = np.zeros((n, 4), np.float32); pts[:, 0] = rng.uniform(3, 6, n); pts[:, 1] = rng.uniform(-2, 2, n); pts[:, 2] = rng.uniform(-1, 1.5, n)
pts[: n // 3, 0] = 4.5 + 0.1 * pts[: n // 3, 1] + rng.normal(0, 0.01, n // 3)
tri = rng.integers(0, n, (801, 3)).astype(np.int32)
ctx.ransac_plane_score(pts, tri, 0.08); ctx.synchronize()
t0 = time.perf_counter(); planes, valid, counts = ctx.ransac_plane_score(pts, tri, 0.08); ctx.synchronize()
out["ransac_score_ms_host_buffers"] = (time.perf_counter() - t0) * 1e3
out["ransac_point_hypotheses"] = n * 801
# K3: GMM EM on 20 resident frames x 500 k
frames, _ = synth.gridfit_batch(20, 500000, seed0=1000)
b = capi.Batch(ctx, cfg, 20, 500000)
for f, p in enumerate(frames): b.upload(f, p)
b.gmm_fit(); ctx.synchronize()
t0 = time.perf_counter(); b.gmm_fit(); ctx.synchronize()
out["gmm_20x500k_ms"] = (time.perf_counter() - t0) * 1e3
b.close()
# K6: pose refine, 256 frames x 35 corners
T = synth.gt_extrinsic()
pcs, ims = [], []
for f in range(256):
    _, _, corners_l, img_norm = synth.calib_frame(f, 200, T_cl_gt=T)
    pcs.append(corners_l); ims.append(img_norm)
q = np.array([0.0, 0.0, 0.0, 1.0]); 
def rot_to_quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
Tcl0 = np.r_[rot_to_quat(synth._rot_xyz(0.01, -0.012, 0.008) @ T[:3, :3]), T[:3, 3] + [0.02, -0.01, 0.01]]
ctx.pose_refine(cfg, np.stack(pcs), np.stack(ims), Tcl0); ctx.synchronize()
t0 = time.perf_counter(); res = ctx.pose_refine(cfg, np.stack(pcs), np.stack(ims), Tcl0); ctx.synchronize()
out["pose_refine_256x35_ms"] = (time.perf_counter() - t0) * 1e3
out["pose_refine_iterations"] = int(res[-1]) if np.isscalar(res[-1]) or isinstance(res[-1], (int, np.integer)) else None
print(json.dumps(out))
