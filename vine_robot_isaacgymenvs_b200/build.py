"""In-tree build of ``csrc/libvine_b200.so`` for sm_100a (B200) with plain nvcc.

The .so stays in the tree (git-ignored) so it travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_DIR, "csrc")
SOURCES = ["vine_b200.cu", "vine_mlp.cu", "vine_ppo.cu", "vine_rollout.cu", "vine_lstm.cu", "vine_lstm_net.cu"]
HEADERS = ["vine_device.cuh", "vine_params.h", "vine_umma.cuh", "vine_mlp_common.cuh", "vine_p2p.cuh", "vine_launch.cuh", os.path.join("..", "..", "include", "vine_b200.h")]
OUT = os.path.join(CSRC, "libvine_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


PEAK_SRC = "bench_peak.cu"
PEAK_OUT = os.path.join(CSRC, "libvine_benchpeak.so")   # measurement helper for bench.py only


def _nvcc(out, sources, verbose):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)


def build(force=False, verbose=False):
    if force or needs_build():
        _nvcc(OUT, SOURCES, verbose)
    src = os.path.join(CSRC, PEAK_SRC)
    if force or not os.path.exists(PEAK_OUT) or os.path.getmtime(src) > os.path.getmtime(PEAK_OUT):
        _nvcc(PEAK_OUT, [PEAK_SRC], False)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
