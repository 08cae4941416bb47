"""BASELINE configs[4]: reprojection of 10 M points into a 1920 x 1080 pinhole image (Pt2img + is_in_FoV + colour look-up),
through the C ABI with host buffers. One JSON line; run under `ncu --metrics gpu__time_duration.sum` for the kernel's own time."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = capi.config(); ctx = capi.Context(0)
pts = synth.frustum_cloud(N, 5); img = synth.noise_image(6, cfg.cam_width, cfg.cam_height)
T = np.eye(4, dtype=np.float32); T[0, 3] = 0.05; T[1, 3] = -0.02
rgb, pix, fov = ctx.project_colorize(cfg, pts, T, img); ctx.synchronize()
t = []
for _ in range(reps):
    t0 = time.perf_counter(); rgb, pix, fov = ctx.project_colorize(cfg, pts, T, img); ctx.synchronize(); t.append(time.perf_counter() - t0)
ms = float(np.median(t)) * 1e3
print(json.dumps({"config": "configs[4]: reprojection + colourise", "points": N, "in_fov": int(fov.sum()), "ms_host_buffers": ms,
                  "Mpts_per_s_host_buffers": N / ms / 1e3, "bytes_in": N * 16 + img.nbytes, "bytes_out": N * 12}))
