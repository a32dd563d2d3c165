"""Config-4 shaped window (OM4p125, 12 steps): global / local kernel time, optionally with ML_TMA_COLSPLIT set."""
import json, os, pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from momlevel_b200 import core, synth
nt, nz, ny, nx = 12, 75, 2240, 2880
grid = synth.make_grid(nz, ny, nx, seed=11, device="cuda")
T, S, V = synth.make_fields(grid, nt, seed=55, dtype=torch.float32)
pres = (grid["z_l"] * 1e4 + 101325.0).contiguous()
z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
pts = nt * nz * ny * nx
def timed(fn, n=4):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
res = {"split": os.environ.get("ML_TMA_COLSPLIT", "1")}
res["global"] = round(pts / timed(lambda: core.steric_global(T, S, V, pres)) / 1e6, 1)
res["selfref"] = round(pts / timed(lambda: core.steric_local_selfref(T, S, V, z_i, depth, pres)) / 1e6, 1)
print(json.dumps(res), flush=True)
