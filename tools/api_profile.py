"""Where the wall time of the public call goes: cProfile of momlevel_b200.steric(dset) on device-resident OM4p25 x 12."""
import cProfile
import pathlib
import pstats
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import momlevel_b200 as ml  # noqa: E402
from momlevel_b200 import core, synth  # noqa: E402

nt, nz, ny, nx = synth.CONFIGS["om4p25"]
dev = torch.device("cuda", 0)
grid = synth.make_grid(nz, ny, nx, seed=123, device=dev)
T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
dset = synth.dataset_from_fields(grid, T, S, V)
for _ in range(3):
    ml.steric(dset)
torch.cuda.synchronize()
n = 30
t0 = time.perf_counter()
for _ in range(n):
    ml.steric(dset)
torch.cuda.synchronize()
print("wall ms per call", (time.perf_counter() - t0) / n * 1e3)
pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
t0 = time.perf_counter()
for _ in range(n):
    core.steric_local_selfref(T, S, V, grid["z_i"], grid["deptho"], pres, want_rho_ref=False)
torch.cuda.synchronize()
print("core call ms", (time.perf_counter() - t0) / n * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    ml.steric(dset)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
