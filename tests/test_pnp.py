"""PnP initialisation (rclc_pnp_ransac, host code) against OpenCV's solvePnPRansac — the call the reference makes at
apps/Calibrate.cpp:41-50 — on the same synthetic correspondences. cv2 is the reference's own dependency and plays the oracle
here; the random sequences differ, so parity is on the pose: both must recover the ground truth to the same accuracy."""
import numpy as np
import pytest

from rclc_b200 import capi, synth

cv2 = pytest.importorskip("cv2")
F_PIX = 1000.0
K = np.array([[F_PIX, 0, 960.0], [0, F_PIX, 540.0], [0, 0, 1]])


def _corr(frames, noise, n_out, seed):
    rng = np.random.default_rng(seed)
    T = synth.gt_extrinsic()
    obj = np.concatenate([synth.calib_frame(f, 500)[2] for f in frames])
    P = (T[:3, :3] @ obj.T).T + T[:3, 3]
    uv = P[:, :2] / P[:, 2:3] + rng.normal(0, noise, (len(P), 2))
    if n_out:
        k = rng.choice(len(uv), n_out, replace=False)
        uv[k] += rng.uniform(-0.2, 0.2, (n_out, 2))
    return obj, uv, T


def _err(R, t, T):
    ang = np.degrees(np.arccos(np.clip((np.trace(R @ T[:3, :3].T) - 1) / 2, -1, 1)))
    return ang, float(np.linalg.norm(np.ravel(t) - T[:3, 3]))


def _cv(obj, uv):
    px = uv * F_PIX + [960.0, 540.0]
    ok, rv, tv, inl = cv2.solvePnPRansac(obj, px, K, None, flags=cv2.SOLVEPNP_EPNP)      # what SOLVEPNP_UPNP runs since OpenCV 3
    assert ok
    return cv2.Rodrigues(rv)[0], tv.ravel(), len(inl)


def test_exact_correspondences_give_the_exact_pose():
    obj, uv, T = _corr(range(4), 0.0, 0, 1)
    R, t, mask = capi.pnp_ransac(obj, uv, 8 / F_PIX)
    ang, dt = _err(R, t, T)
    assert mask.all() and ang < 1e-6 and dt < 1e-7
    assert abs(np.linalg.det(R) - 1) < 1e-12 and np.allclose(R @ R.T, np.eye(3), atol=1e-12)


@pytest.mark.parametrize("frames,n_out,seed", [(range(5), 20, 2), (range(12), 60, 3), (range(2), 0, 4)])
def test_matches_opencv_accuracy_with_outliers(frames, n_out, seed):
    obj, uv, T = _corr(frames, 2e-4, n_out, seed)
    R, t, mask = capi.pnp_ransac(obj, uv, 8 / F_PIX)
    Rc, tc, n_inl = _cv(obj, uv)
    (a0, d0), (a1, d1) = _err(Rc, tc, T), _err(R, t, T)
    assert a1 <= 1.5 * a0 + 0.02 and d1 <= 1.5 * d0 + 2e-3, ((a1, d1), (a0, d0))
    assert abs(int(mask.sum()) - n_inl) <= 3
    assert mask.sum() >= len(uv) - n_out - 2


def test_single_board_is_planar_and_still_solved():
    """One frame = 35 coplanar corners: the homography branch. OpenCV's EPnP handles the same case."""
    obj, uv, T = _corr([3], 1e-4, 0, 5)
    R, t, mask = capi.pnp_ransac(obj, uv, 8 / F_PIX)
    Rc, tc, _ = _cv(obj, uv)
    (a0, d0), (a1, d1) = _err(Rc, tc, T), _err(R, t, T)
    assert a1 <= 1.5 * a0 + 0.1 and d1 <= 1.5 * d0 + 0.01, ((a1, d1), (a0, d0))
    assert mask.all()


def test_too_few_points_are_refused():
    """Bad arguments are an error, not a guess."""
    with pytest.raises(Exception):
        capi.pnp_ransac(np.zeros((3, 3)), np.zeros((3, 2)), 0.01)
