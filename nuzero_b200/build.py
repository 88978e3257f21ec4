"""Builds libnz_engine.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m nuzero_b200.build            # compile if sources are newer than the library
    python -m nuzero_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnz_engine.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "game_ttt.cuh", "game_scs.cuh", "mcts.cuh", "hexgemm.cuh", os.path.join("..", "..", "include", "nz_engine.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (e.g. defines=["NZ_TTT_TILE=16"]) next to the default."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out or LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libnz_engine.so")
    if out is None:
        with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:
            f.write(res.stdout)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
