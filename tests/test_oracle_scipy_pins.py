"""Independent pins of the oracle's two smooth-enough solves: the reference holds no vectors on the hot path,
so the restated Ceres LM (oracle/ceres_lm.hpp) is checked at its CONVERGED COST against scipy.optimize.least_squares — another
minimiser, another implementation — on the same residuals:

  * pc_plane_fitting (include/opt/opt_plane_param.hpp:24-88): point-to-plane distances under HuberLoss(plane_huber_a);
  * pose_reprj_opt (include/opt/opt_reprj_calib.h:21-66): one POSE_REPRJ_COST residual per frame over the quaternion's tangent space.

What is compared is the value of the objective at the two solutions (and the solutions themselves to the accuracy the objective
pins them), not iterates: both minimisers stop on their own tolerances."""
import numpy as np
import pytest

scipy_opt = pytest.importorskip("scipy.optimize")


def _huber_cost(r, a):
    s = r * r
    rho = np.where(s <= a * a, s, 2 * a * np.sqrt(s) - a * a)          # ceres::HuberLoss
    return 0.5 * rho.sum()


def test_plane_fit_reaches_scipys_minimum(oracle, ocfg):
    rng = np.random.default_rng(21)
    n = 6000
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-0.4, 0.4, n); pts[:, 1] = rng.uniform(-0.3, 0.3, n)
    nrm = np.array([0.2, -0.12, 1.0]); nrm /= np.linalg.norm(nrm)
    pts[:, 2] = (4.0 - nrm[0] * pts[:, 0] - nrm[1] * pts[:, 1]) / nrm[2] + rng.normal(0, 0.004, n)
    out = rng.choice(n, n // 15, replace=False)
    pts[out, 2] += rng.uniform(0.2, 0.6, len(out))                      # outliers far beyond the Huber knee
    pts[:, 3] = 120.0                                                    # every point is used (plane_reflactive_thd)
    rc, plane, its, used = oracle.plane_fit(ocfg, pts)
    assert rc == 0 and used == n
    X = pts[:, :3].astype(np.float64)
    a = ocfg.plane_huber_a

    def resid(p):
        return np.abs(X @ p[:3] + p[3]) / np.linalg.norm(p[:3])

    sol = scipy_opt.least_squares(resid, np.array([0.1, 0.1, 1.0, 1.0]), loss="huber", f_scale=a, xtol=1e-14, ftol=1e-14, gtol=1e-14)
    c_or, c_sp = _huber_cost(resid(plane), a), _huber_cost(resid(sol.x), a)
    assert abs(c_or - c_sp) <= 2e-5 * c_sp                               # Ceres stops on function_tolerance 1e-6
    po, ps = plane / np.linalg.norm(plane[:3]), sol.x / np.linalg.norm(sol.x[:3])
    if po @ ps < 0:
        ps = -ps
    np.testing.assert_allclose(po, ps, atol=6e-4)            # (a relative cost change of 1e-6 stops Ceres: the parameters are pinned to ~1e-3 of their scale)


def _plus(T, d):
    """ceres::EigenQuaternionParameterization::Plus on (qx, qy, qz, qw | t)."""
    dq = d[:3]; nd = np.linalg.norm(dq)
    if nd > 0:
        s = np.sin(nd) / nd; dw, dv = np.cos(nd), s * dq
        xw, xv = T[3], T[:3]
        qn = np.r_[dw * xv + xw * dv + np.cross(dv, xv), dw * xw - dv @ xv]
    else:
        qn = T[:4]
    return np.r_[qn, T[4:] + d[3:]]


@pytest.mark.parametrize("F,seed", [(20, 9), (64, 3)])
def test_pose_refine_reaches_scipys_minimum(oracle, ocfg, F, seed):
    rng = np.random.default_rng(seed)
    Cn = 35
    q = np.array([0.02, -0.015, 0.01, 0.0]); q[3] = np.sqrt(1 - (q[:3] ** 2).sum())
    t = np.array([0.07, -0.08, 0.02])
    R = oracle.quat_to_R([q[3], q[0], q[1], q[2]])
    pc = np.zeros((F, Cn, 3)); img = np.zeros((F, Cn, 2))
    for f in range(F):
        base = np.array([rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), rng.uniform(3.5, 5.5)])
        g = np.array([[0.3 - 0.1 * c, 0.2 - 0.1 * r, 0.0] for r in range(5) for c in range(7)])
        a = rng.uniform(-0.3, 0.3, 3)
        Rb = oracle.quat_to_R(np.r_[np.sqrt(1 - (a ** 2).sum() / 4), a / 2])
        pc[f] = g @ Rb.T + base
        Pc = pc[f] @ R.T + t
        img[f] = Pc[:, :2] / Pc[:, 2:3] + rng.normal(0, 1e-4, (Cn, 2))
    T0 = _plus(np.r_[q, t], np.array([0.01, -0.008, 0.006, 0.03, -0.02, 0.04]))
    rc, T1, c0, c1, its = oracle.pose_refine(ocfg, pc, img, T0)
    assert rc == 0 and c1 < c0

    def resid(d):
        return oracle.pose_eval(pc, img, _plus(T0, d))[0]

    # the residual of a frame is a sum of norms plus a maximum: not smooth at the minimum, so scipy gets several starts around the
    # oracle's answer as well as the common start, and the best of them stands for "the minimum"
    best = np.inf
    starts = [np.zeros(6)]
    r = np.random.default_rng(1)
    for _ in range(3):
        starts.append(r.normal(0, 1e-3, 6))
    for d0 in starts:
        sol = scipy_opt.least_squares(resid, d0, method="trf", x_scale=np.r_[np.full(3, 1e-2), np.full(3, 1e-1)], xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
        best = min(best, 0.5 * float(sol.fun @ sol.fun))
    c_or = 0.5 * float(np.sum(oracle.pose_eval(pc, img, T1)[0] ** 2))
    assert abs(c_or - c1) <= 1e-12 * c1                                  # the cost the solve reports is the cost at what it returns
    assert c_or <= best * (1 + 2e-3)                                     # as low as another minimiser gets, to the flatness of the valley
    assert best <= c_or * (1 + 2e-3)
