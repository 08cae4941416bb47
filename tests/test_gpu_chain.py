"""The per-frame chains on resident clouds (csrc/calib.cu) against the CPU oracle's restatement of the same chains
(oracle/rclc_pipeline.cpp) on raw scans of the synthetic MID-40 rig: BASELINE configs[0].

  rclc_segment_board      apps/BoardSegmentation.cpp:45-165      -> the oracle's points, bit for bit, in the oracle's order
  rclc_est_board_corner   include/pc_process.h:583-646           -> the oracle's LM trajectory and corners (1 um)
  rclc_calibrate_frames   both stages for F scans + calib_opt    -> per-frame results equal to the single-frame calls;
                                                                    extrinsic within 1e-4 rad / 0.1 mm of the oracle's
"""
import numpy as np
import pytest

from rclc_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from rclc_b200 import capi
    with capi.Context(0) as c:
        yield c


@pytest.fixture(scope="module")
def cfg():
    from rclc_b200 import capi
    return capi.config()


@pytest.fixture(scope="module")
def oracle():
    from oracle import pyoracle
    return pyoracle


def _to_cam(board):
    return np.c_[board[:, :3] @ synth.T_BASE[:3, :3].T, board[:, 3]].astype(np.float32)


@pytest.mark.parametrize("f", [0, 1, 5])
def test_segment_board_bit_exact(ctx, cfg, oracle, f):
    raw, _, _ = synth.raw_scan(f)
    rc, want = oracle.segment_board(oracle.config(), raw)
    assert rc == 0 and len(want) > 30000
    for _ in range(4):                                      # (repeated: the chain is a sequence of racy-by-nature parallel primitives)
        got = ctx.segment_board(cfg, raw)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_segment_board_no_board_and_strided_input(ctx, cfg, oracle):
    raw, _, _ = synth.raw_scan(2)
    # PointXYZI layout (32-byte records, intensity at +16)
    pcl = np.zeros((len(raw), 8), np.float32); pcl[:, :3] = raw[:, :3]; pcl[:, 4] = raw[:, 3]
    rc, want = oracle.segment_board(oracle.config(), raw)
    got = ctx.segment_board(cfg, pcl)
    assert rc == 0 and np.array_equal(got, want)
    # nothing inside the x range: no plane, empty result (the oracle returns 1)
    far = raw.copy(); far[:, 0] += 10.0
    rc, want = oracle.segment_board(oracle.config(), far)
    assert rc == 1 and len(ctx.segment_board(cfg, far)) == 0
    # a scene without the board: planes are found, none faces the lidar within reach or the candidate has no cluster
    rng = np.random.default_rng(3)
    n = 60000
    ground = np.c_[rng.uniform(3, 6, n), rng.uniform(-2, 2, n), -1.2 + rng.normal(0, 0.01, n), np.full(n, 40.0)].astype(np.float32)
    rc, want = oracle.segment_board(oracle.config(), ground)
    got = ctx.segment_board(cfg, ground)
    assert (rc == 0) == (len(got) > 0) and (rc != 0 or np.array_equal(got, want))


@pytest.mark.parametrize("f", [0, 3, 5])
def test_est_board_corner_matches_oracle(ctx, cfg, oracle, f):
    raw, _, truth = synth.raw_scan(f)
    rc, board = oracle.segment_board(oracle.config(), raw)
    board = _to_cam(board)
    rc0, c0, r0, cloud0 = oracle.est_board_corner(oracle.config(), board)
    st, c1, r1, cloud1 = ctx.est_board_corner(cfg, board)
    assert rc0 == 0 and st == 0
    assert np.array_equal(cloud1.view(np.uint32), cloud0.view(np.uint32))          # projected + moved cloud: the same floats
    assert r1.pose_evals == r0.pose_evals and r1.outer_iterations == r0.outer_iterations
    assert np.abs(c1 - c0).max() <= 2e-15                                           # the corners: the oracle's to an ulp
    if np.array_equal(np.array(r1.T_base_laser, np.float32), np.array(r0.T_base_laser, np.float32)):
        assert c1.tobytes() == c0.tobytes()                                         # (bit for bit when the float frame agrees)
    assert np.linalg.norm(c1 - truth, axis=1).max() < 6e-3                          # and they are the board's corners


def test_est_board_corner_on_degenerate_boards_like_the_oracle(ctx, cfg, oracle):
    """Thinned-out boards stop at the block centroids on both sides; a board without its black squares runs into NaN residuals - a
    Ceres FAILURE, which the reference never looks at (opt_grid_fitting.h:96-131): no failed stage, status 1 in the report, and the
    corners of the pose the fit started from, the oracle's to an ulp."""
    raw, _, _ = synth.raw_scan(0)
    rc, board = oracle.segment_board(oracle.config(), raw)
    board = _to_cam(board)
    for thin in (board[::50], board[:100]):
        rc0 = oracle.est_board_corner(oracle.config(), thin)[0]
        st = ctx.est_board_corner(cfg, thin)[0]
        assert rc0 == st == 2
    white = board[board[:, 3] > 60]
    rc0, c0, r0, cloud0 = oracle.est_board_corner(oracle.config(), white)
    st, c1, r1, cloud1 = ctx.est_board_corner(cfg, white)
    assert rc0 == 0 and st == 0 and r0.status == r1.status == 1
    assert np.array_equal(cloud1.view(np.uint32), cloud0.view(np.uint32))
    assert np.abs(c1 - c0).max() <= 2e-15


def test_calibrate_frames_with_scans_that_hold_no_board(ctx, cfg, oracle):
    """Two of six scans yield no board (nothing in the region of interest; a scan too thin for a plane): status 10 for them, exactly
    where the oracle's BoardSegmentation gives up, and the extrinsic over the four that remain - the oracle's solve over its own
    corners of those four, from the same start - within the bar."""
    F = 6
    ocfg = oracle.config()
    scans, imgs, _ = zip(*[synth.raw_scan(f) for f in range(F)])
    scans = list(scans)
    scans[2] = scans[2] + np.array([100, 0, 0, 0], np.float32)
    scans[4] = scans[4][::40].copy()
    out = ctx.calibrate_frames(cfg, scans, np.stack(imgs))
    oc, keep = [], []
    for f in range(F):
        rc, ob = oracle.segment_board(ocfg, scans[f])
        assert (rc != 0) == (out["status"][f] == 10), f
        if rc != 0:
            continue
        rc, c0, _, _ = oracle.est_board_corner(ocfg, _to_cam(ob))
        assert rc == 0 and out["status"][f] == 0 and out["n_board"][f] == len(ob)
        assert np.abs(out["corners"][f] - c0).max() <= 2e-15
        oc.append(c0); keep.append(f)
    assert keep == [0, 1, 3, 5]
    C0, I0 = np.stack(oc), np.stack([imgs[f] for f in keep])
    rc, T0, _, _, it0 = oracle.pose_refine(ocfg, C0, I0, out_start(ctx, cfg, C0, I0))
    assert rc == 0 and np.abs(out["Tcl"] - T0).max() < 1e-4


def test_calibrate_frames_equals_single_frame_calls_and_oracle_extrinsic(ctx, cfg, oracle):
    F = 6
    ocfg = oracle.config()
    scans, imgs, truth = zip(*[synth.raw_scan(f) for f in range(F)])
    out = ctx.calibrate_frames(cfg, list(scans), np.stack(imgs))
    assert (out["status"] == 0).all()
    oc = []
    for f in range(F):
        board = ctx.segment_board(cfg, scans[f])
        assert out["n_board"][f] == len(board)
        st, c1, r1, _ = ctx.est_board_corner(cfg, _to_cam(board))
        assert st == 0 and np.array_equal(out["corners"][f], c1)                   # batched == single frame, bit for bit
        assert out["reports"][f].pose_evals == r1.pose_evals
        rc, ob = oracle.segment_board(ocfg, scans[f])
        rc, c0, _, _ = oracle.est_board_corner(ocfg, _to_cam(ob))
        assert rc == 0 and np.abs(c1 - c0).max() <= 2e-15                           # the oracle's corners to an ulp
        oc.append(c0)
    # the extrinsic: the same PnP start, pose_reprj_opt on the oracle's corners by the oracle
    rc, T0, _, _, _ = oracle.pose_refine(ocfg, np.stack(oc), np.stack(imgs), out_start(ctx, cfg, np.stack(oc), np.stack(imgs)))
    T1 = out["Tcl"]
    q0, q1 = T0[:4] / np.linalg.norm(T0[:4]), T1[:4] / np.linalg.norm(T1[:4])
    if q0 @ q1 < 0:
        q1 = -q1
    assert 2 * np.abs(q0 - q1).max() < 1e-4 and np.abs(T0[4:] - T1[4:]).max() < 1e-4          # the bar
    # the solve itself is the oracle's bit for bit (tests/test_gpu_parity.py): fed the ORACLE's corners and the same start it
    # returns the oracle's extrinsic exactly; what separates T1 from T0 is the ulp in the corners of a few frames
    rc2, T2, _, _, _ = ctx.pose_refine(cfg, np.stack(oc), np.stack(imgs), out_start(ctx, cfg, np.stack(oc), np.stack(imgs)))
    assert rc2 == 0 and T2.tobytes() == T0.tobytes()
    # and it is the rig's extrinsic
    Tgt = synth.scan_extrinsic()
    R1 = oracle.quat_to_R([q1[3], q1[0], q1[1], q1[2]])
    ang = np.arccos(np.clip((np.trace(R1 @ Tgt[:3, :3].T) - 1) / 2, -1, 1))
    assert ang < 3e-3 and np.abs(T1[4:] - Tgt[:3, 3]).max() < 0.02


def test_calibration_of_configs0_in_full_against_the_oracle(ctx, cfg, oracle):
    """BASELINE configs[0] at its full size — 20 raw MID-40 scans x 100 k points through both stages and the extrinsic solve —
    against the oracle's chains on every frame (a thread per frame on the host): the same board points, corners within 1 um,
    the extrinsic within the north star's 1e-4 rad / 0.1 mm."""
    from concurrent.futures import ThreadPoolExecutor
    F = 20
    ocfg = oracle.config()
    scans, imgs, _ = zip(*[synth.raw_scan(f) for f in range(F)])
    out = ctx.calibrate_frames(cfg, list(scans), np.stack(imgs))
    assert (out["status"] == 0).all()

    def chain(raw):
        rc, board = oracle.segment_board(ocfg, raw)
        assert rc == 0
        rc, c0, r0, _ = oracle.est_board_corner(ocfg, _to_cam(board))
        assert rc == 0
        return len(board), c0, r0

    with ThreadPoolExecutor(max_workers=F) as ex:
        ref = list(ex.map(chain, scans))
    for f, (nb, c0, r0) in enumerate(ref):
        assert out["n_board"][f] == nb
        assert out["reports"][f].pose_evals == r0.pose_evals and out["reports"][f].outer_iterations == r0.outer_iterations
        assert np.abs(out["corners"][f] - c0).max() <= 2e-15, (f, np.abs(out["corners"][f] - c0).max())
    print("frames whose corners are the oracle's bit for bit:", sum(out["corners"][f].tobytes() == ref[f][1].tobytes() for f in range(len(ref))), "of", len(ref))
    oc = np.stack([r[1] for r in ref])
    rc, T0, _, c1o, _ = oracle.pose_refine(ocfg, oc, np.stack(imgs), out_start(ctx, cfg, oc, np.stack(imgs)))
    T1 = out["Tcl"]
    q0, q1 = T0[:4] / np.linalg.norm(T0[:4]), T1[:4] / np.linalg.norm(T1[:4])
    if q0 @ q1 < 0:
        q1 = -q1
    assert 2 * np.abs(q0 - q1).max() < 1e-4 and np.abs(T0[4:] - T1[4:]).max() < 1e-4          # the bar
    # the solve itself is the oracle's bit for bit (tests/test_gpu_parity.py): fed the ORACLE's corners and the same start it
    # returns the oracle's extrinsic exactly; what separates T1 from T0 is the ulp in the corners of a few frames
    rc2, T2, _, _, _ = ctx.pose_refine(cfg, oc, np.stack(imgs), out_start(ctx, cfg, oc, np.stack(imgs)))
    assert rc2 == 0 and T2.tobytes() == T0.tobytes()


def out_start(ctx, cfg, corners, imgs):
    """The PnP start rclc_calibrate_frames uses (apps/Calibrate.cpp:41-62), rebuilt through the same entry points."""
    from rclc_b200 import capi
    import ctypes as C
    R, t, _ = capi.pnp_ransac(corners.reshape(-1, 3), imgs.reshape(-1, 2), 8.0 / (0.5 * (cfg.fx + cfg.fy)))[:3]
    q = np.zeros(4)
    capi.lib().rclc_quat_from_R(np.ascontiguousarray(R, np.float64).ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
    return np.r_[q, t]
