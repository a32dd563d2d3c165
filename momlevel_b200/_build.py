"""Build libmomlevel_b200.so in-tree with nvcc for sm_100a (no JIT cache, no pip install).

    python -m momlevel_b200._build [--force] [--verbose]

The built library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""

import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmomlevel_b200.so"
SOURCES = ["ml_api.cu", "ml_tma.cu", "ml_tma3.cu", "ml_stream.cu", "ml_hostpath.cu", "ml_strat.cu", "ml_pack.cpp"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "--threads",
    "0",
    "-shared",
    "-Xcompiler",
    "-fPIC,-pthread",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not pathlib.Path(exe).exists():
        raise RuntimeError("nvcc not found; libmomlevel_b200.so cannot be built")
    return exe


def needs_build():
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*.c*")) + [PKG.parent / "include" / "momlevel_b200.h"])
    return newest > LIB.stat().st_mtime


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library. Returns the library path."""
    if not force and not needs_build():
        return LIB
    import os
    import shlex

    # extra flags for experiment builds, e.g. MOMLEVEL_B200_NVCC_FLAGS="-DML_TMA_FENCED_RELEASE" (csrc/ml_tma_dev.cuh)
    extra = shlex.split(os.environ.get("MOMLEVEL_B200_NVCC_FLAGS", ""))
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else [])
    cmd += ["-o", str(LIB)] + [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libmomlevel_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
