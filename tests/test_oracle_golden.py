"""Pin the oracle's pose utilities against the ONLY known answers the reference holds:
the literals and expected-output comments of apps/calc_rmse.cpp (inputs :70-74, :113-121;
expected :142-164). They pin quaternion->R (Eigen toRotationMatrix), rotationMatrixToEulerAngles
(include/utils.h:422-442 == apps/calc_rmse.cpp:47-66) and rigid composition/inversion — adjacent to
the hot path. Nothing in the reference pins the hot path itself."""
import numpy as np

T_LR_GT = np.array([-0.119994768246006, -0.000174184734047787, 0.000301631548705875])
R_LR_GT = np.array([[0.999999338992956, 4.56035660014599e-05, -0.00114888379145621],
                    [-5.13812077544178e-05, 0.999987351218247, -0.00502938997149094],
                    [0.00114863990135836, 0.00502944567806551, 0.999986692562730]])

CASES = {
    # apps/calc_rmse.cpp:113-116 (indoor), :118-121 (outdoor); q = (w, x, y, z)
    "indoor": dict(qR=(0.999776, -0.0177691, -0.00314187, 0.0110607), tR=(-0.0426148, -0.087164, -0.0198392),
                   qL=(0.999731, -0.0202243, -0.00235128, 0.0111032), tL=(0.0748039, -0.0858426, -0.02457),
                   # expected :146-153
                   R_eul=(-35.6079, -5.88929, 22.2303), L_eul=(-40.5014, -4.2522, 22.2976),
                   T_eul=(4.92869, -1.52763, -0.0462151), dt=(2.54194, -1.2659, 4.73766),
                   dr=(-0.100777, -0.378992, 0.00516618)),
    "outdoor": dict(qR=(0.999795, -0.0168687, -0.000874041, 0.0111297), tR=(-0.0496541, -0.0806369, 0.0058639),
                    qL=(0.999747, -0.0195042, -0.00029811, 0.011238), tL=(0.068088, -0.0805071, 0.000370015),
                    # expected :157-164
                    R_eul=(-33.7564, -1.37224, 22.2862), L_eul=(-39.0151, -0.157693, 22.4838),
                    T_eul=(5.28472, -1.09601, -0.19652), dt=(2.26847, 0.0585949, 5.54309),
                    dr=(0.255252, 0.0526271, -0.145139)),
}


def _close6(got, exp):
    """The reference prints with cout's default 6 significant digits (inputs are 6-digit literals too)."""
    got, exp = np.asarray(got, float), np.asarray(exp, float)
    assert np.all(np.abs(got - exp) <= 1.0e-5 * np.abs(exp) + 1e-9), (got, exp)


def test_gt_euler(oracle):
    # apps/calc_rmse.cpp:142-143 "gt eular 5.02947 -1.14864 -0.0513812"
    e = oracle.R_to_euler(R_LR_GT) * 1000.0
    _close6(e, (5.02947, -1.14864, -0.0513812))


def test_calc_rmse_known_answers(oracle):
    eul_gt = oracle.R_to_euler(R_LR_GT)
    for name, c in CASES.items():
        RR = oracle.quat_to_R(c["qR"])
        RL = oracle.quat_to_R(c["qL"])
        _close6(oracle.R_to_euler(RR) * 1000.0, c["R_eul"])
        _close6(oracle.R_to_euler(RL) * 1000.0, c["L_eul"])
        TR = np.eye(4); TR[:3, :3] = RR; TR[:3, 3] = c["tR"]
        TL = np.eye(4); TL[:3, :3] = RL; TL[:3, 3] = c["tL"]
        T = TR @ oracle.inverse4(TL)                       # calc_rmse.cpp:136
        _close6(oracle.R_to_euler(T[:3, :3]) * 1000.0, c["T_eul"])
        _close6((T[:3, 3] - T_LR_GT) * 1000.0, c["dt"])
        _close6((oracle.R_to_euler(T[:3, :3]) - eul_gt) * 1000.0, c["dr"])


def test_median_error_literal(oracle):
    # apps/calc_rmse.cpp:87-100: sorted err list, element size/2 -> printed "0.00599" region
    err = [0.00681016, 0.00754979, 0.00543653, 0.00620938, 0.00664616, 0.00542783, 0.00776292, 0.00686705,
           0.00473081, 0.00515263, 0.00697178, 0.00625069, 0.0064709, 0.00647485, 0.00598959, 0.00676013,
           0.00605955, 0.00527193, 0.0058098, 0.00494896, 0.00551903, 0.00576172, 0.0066745, 0.00528792,
           0.00622681, 0.00480505, 0.00581097, 0.0056832, 0.00565476, 0.0056538]
    s = sorted((e, i) for i, e in enumerate(err))
    assert abs(s[len(err) // 2][0] - 0.00598959) < 1e-12
