"""Timing of the SURVEY.md §8(f) "next" rows (plane fit, uniform sampling, Euclidean clustering) through the C ABI with host
buffers (upload included), next to the CPU oracle on the same inputs. One JSON line per op. Run on a GPU box."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rclc_b200 import capi, synth
from oracle import pyoracle as orc

cfg = capi.config(); ocfg = orc.config(); ctx = capi.Context(0)


def timed(fn, reps):
    fn(); ctx.synchronize()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ctx.synchronize(); t.append(time.perf_counter() - t0)
    return float(np.median(t)) * 1e3


def scene(n, seed):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, 4), np.float32)
    m = int(n * 0.8)
    pts[:m, 0] = rng.uniform(-0.4, 0.4, m); pts[:m, 1] = rng.uniform(-0.3, 0.3, m); pts[:m, 2] = 4.0 + 0.1 * pts[:m, 0] + rng.normal(0, 0.003, m)
    pts[m:, :3] = rng.uniform(-2, 2, (n - m, 3)) + [0, 0, 7]
    pts[:, 3] = np.where(rng.uniform(size=n) < 0.5, 15, 110)
    return pts


out = []
# ---- pc_plane_fitting: 500 k points, half of them above the reflectance threshold ----
pts = scene(500000, 1)[:400000]
l0 = ctx.kernel_launches()
rc, plane, its, used = ctx.plane_fit(cfg, pts)
launches = ctx.kernel_launches() - l0
ms = timed(lambda: ctx.plane_fit(cfg, pts), 5)
t0 = time.perf_counter(); orc.plane_fit(ocfg, pts); cpu_ms = (time.perf_counter() - t0) * 1e3
out.append({"op": "rclc_plane_fit", "n": len(pts), "used": int(used), "ceres_iterations": int(its), "kernel_launches": int(launches),
            "ms_per_fit_host_buffers": ms, "cpu_oracle_ms": cpu_ms, "speedup": cpu_ms / ms})
# ---- UniformSampling: 2 M points, 8 mm leaf ----
pts = scene(2000000, 2)
ms = timed(lambda: ctx.uniform_sample(pts, 0.008), 5)
kept = len(ctx.uniform_sample(pts, 0.008))
t0 = time.perf_counter(); orc.uniform_sample(pts, 0.008); cpu_ms = (time.perf_counter() - t0) * 1e3
out.append({"op": "rclc_uniform_sample", "n": len(pts), "leaf": 0.008, "kept": kept, "ms_host_buffers": ms, "Mpts_per_s": len(pts) / ms / 1e3,
            "cpu_oracle_ms": cpu_ms, "speedup": cpu_ms / ms})
# ---- EuclideanClusterExtraction: the thinned cloud, 3 cm tolerance (EulerCluster) and 8.5 mm (eular_cluster_iter) ----
thin = pts[ctx.uniform_sample(pts, 0.004)]
for tol, mn in ((0.03, 500), (0.0085, 3)):
    ms = timed(lambda: ctx.euclidean_cluster(thin, tol, mn, 2500000), 5)
    lab, sizes = ctx.euclidean_cluster(thin, tol, mn, 2500000)
    t0 = time.perf_counter(); orc.euclidean_cluster(thin, tol, mn, 2500000); cpu_ms = (time.perf_counter() - t0) * 1e3
    out.append({"op": "rclc_euclidean_cluster", "n": len(thin), "tol": tol, "min_size": mn, "clusters": len(sizes), "largest": int(sizes[0]) if len(sizes) else 0,
                "ms_host_buffers": ms, "cpu_oracle_ms": cpu_ms, "speedup": cpu_ms / ms})
for o in out:
    print(json.dumps(o))
