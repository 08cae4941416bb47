"""Small end-to-end pass over every device entry point (a quick "does everything still run" check on a GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
cfg = capi.config(); ctx = capi.Context(0)
pts = synth.board_frame_cloud(30000, 3, pert=(0.004, -0.003, 0.3))
print("eval", ctx.gridfit_eval(cfg, pts, (0.001, 0.002, 0.1))[0][:2])
print("solve", ctx.gridfit_solve(cfg, pts)[1].pose_evals)
print("multistart", ctx.gridfit_multistart(cfg, pts, synth.multistart_poses(3, 3, 3)).min())
print("gmm", ctx.gmm_fit_1d(pts[:, 3], [15, 100], [15, 15])[0])
pl = pts.copy(); pl[:, 2] += 4.0
print("plane_fit", ctx.plane_fit(cfg, pl)[1])
idx = ctx.uniform_sample(pl, 0.008); print("uniform", len(idx))
lab, sizes = ctx.euclidean_cluster(pl[idx], 0.03, 10, 100000); print("cluster", sizes[:3])
print("project", ctx.plane_project(pl, np.array([0.0, 0.0, 1.0, -4.0]))[:1])
b = capi.Batch(ctx, cfg, 4, 30000)
frames, perts = synth.gridfit_batch(4, 30000, seed0=50)
for f, p in enumerate(frames): b.upload(f, p)
c, r = b.gridfit_solve(); print("batch solve", [x.pose_evals for x in r])
b.gmm_fit(); b.close()
img = synth.noise_image(1, cfg.cam_width, cfg.cam_height)
fr = synth.frustum_cloud(50000, 2)
print("colorize", ctx.project_colorize(cfg, fr, np.eye(4, dtype=np.float32), img)[2].sum())
ctx.close(); print("done")
