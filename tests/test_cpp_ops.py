"""The reference-shaped C++ operator surface (include/rclc_ops.hpp): compiles with g++ against the stand-in PCL/Eigen
types, links against librclc_b200.so, and (on a GPU box) runs every operator end to end (tests/cpp/test_ops.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_ops")


def _build():
    from rclc_b200 import build
    lib = build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_ops.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(lib),
                                                              os.path.getmtime(os.path.join(ROOT, "include", "rclc_ops.hpp"))):
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", os.path.dirname(lib), "-lrclc_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def test_ops_header_compiles_links_and_refuses_to_compute_without_a_device():
    import torch
    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("GPU box: covered by test_ops_run_on_gpu")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)          # loud, no CPU fallback
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_ops_run_on_gpu():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-3000:], r.stderr[-2000:])
    assert "all operator checks passed" in r.stdout


def test_pcd_io_round_trip_and_formats(tmp_path):
    """include/rclc_pcd.hpp (loadPCDFile / savePCDFileBinary / T_base): host code, runs anywhere."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_pcd")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_pcd.cpp")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "livox_ascii.pcd"), str(tmp_path / "pcd")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "all PCD checks passed" in r.stdout


def _build_pipeline():
    from rclc_b200 import build
    lib = build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_pipeline")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_pipeline.cpp")
    from oracle import pyoracle
    pyoracle.build()                               # the checker: every stage of the pipeline is compared with the CPU oracle
    odir = os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
                           src, "-o", exe, "-L", os.path.dirname(lib), "-lrclc_b200", "-Wl,-rpath," + os.path.dirname(lib),
                           "-L", odir, "-lrclc_oracle", "-Wl,-rpath," + odir])
    return exe


def test_pipeline_compiles_and_needs_a_device():
    import torch
    exe = _build_pipeline()
    if torch.cuda.is_available():
        pytest.skip("GPU box: covered by test_pipeline_scene_to_extrinsic")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_pipeline_scene_to_extrinsic():
    """BASELINE configs[0] in miniature: raw scans -> segment_board -> PCD hand-off -> est_board_corner -> pnp_init + pose_reprj_opt."""
    exe = _build_pipeline()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.returncode, r.stdout[-3000:], r.stderr[-2000:])
    assert "pipeline checks passed" in r.stdout
