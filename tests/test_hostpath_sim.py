"""The host entry points (csrc/ml_hostpath.cu + csrc/ml_pack.cpp) on a box without a GPU.

tests/sim/hostpath_sim.cpp compiles the real host-side source against a simulated CUDA runtime
(tests/sim/cuda_runtime.h: streams that run behind the host on their own threads, events, pinned / pageable
bookkeeping) with toy kernels in place of the device ones, and runs every packing mode x thread count x window
width x pinned / pageable source x slow / fast copies under ThreadSanitizer.  Each run must reproduce, bit for bit,
what the toy kernels give on the caller's whole arrays.  Removing any one of the event waits that guard the staging
buffers, the ring slots or the device windows makes it fail (checked by hand when it was written).
"""

import pathlib
import shutil
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "sim" / "hostpath_sim.cpp"


@pytest.mark.timeout(900)
def test_host_path_under_a_simulated_runtime(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = tmp_path / "hostpath_sim"
    base = [gxx, "-std=c++17", "-O1", "-g", "-pthread", "-I", str(ROOT / "tests" / "sim"), "-x", "c++", str(SRC), "-o", str(exe)]
    res = subprocess.run(base[:5] + ["-fsanitize=thread"] + base[5:], capture_output=True, text=True, cwd=ROOT)
    sanitized = res.returncode == 0
    if not sanitized:  # no libtsan here: the results are still checked
        res = subprocess.run(base, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=800)
    out = run.stdout + run.stderr
    assert "ThreadSanitizer" not in out, out[-4000:]
    assert run.returncode == 0 and ", 0 mismatches" in out, out[-4000:]


def test_packing_tuner_unit(tmp_path):
    """csrc/ml_hosttune.h on its own: trial order, the margin plain copies enjoy, the settled-on-none shortcut and its
    retry point, a rank's share of the cores (tests/sim/tuner_unit.cpp)."""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = tmp_path / "tuner_unit"
    src = ROOT / "tests" / "sim" / "tuner_unit.cpp"
    res = subprocess.run([gxx, "-std=c++17", "-DML_HOSTPATH_TEST_HOOKS", "-I", str(ROOT / "tests" / "sim"), "-x", "c++", str(src),
                          "-o", str(exe)], capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0 and "0 failed" in run.stdout, run.stdout[-2000:]
