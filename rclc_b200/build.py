"""Builds librclc_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librclc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function,-fopenmp", "-ccbin", "/usr/bin/g++"] + \
    os.environ.get("RCLC_EXTRA_NVCC_FLAGS", "").split()          # experiments only (e.g. -DRCLC_SWEEP_TIMING)
# pose.cu spells out the oracle's arithmetic operation by operation: no multiply-add is fused that the source does not fuse
FILE_FLAGS = {"pose.cu": ["-fmad=false"]}


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs = []
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(HERE, "_obj", os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(
                os.path.getmtime(d) for d in [s] + glob.glob(os.path.join(CSRC, "*.cuh")) +
                glob.glob(os.path.join(HERE, "..", "include", "*.h"))):
            continue
        cmd = [NVCC] + FLAGS + FILE_FLAGS.get(os.path.basename(s), []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + s)
        if verbose or "warning" in out:
            sys.stderr.write(out)
    subprocess.check_call([NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart_static", "-ldl", "-lrt", "-lgomp",
                                                                 "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
