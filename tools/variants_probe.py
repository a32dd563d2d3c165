import pathlib, sys, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from momlevel_b200 import core, synth
nt, nz, ny, nx = 12, 75, 1080, 1440
grid = synth.make_grid(nz, ny, nx, seed=123, device="cuda")
T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
pres = (grid["z_l"] * 1e4 + 101325.0).contiguous()
z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
pts = nt * nz * ny * nx
def timed(fn, n=4):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); del r
        best = min(best, a.elapsed_time(b))
    return best
ms = timed(lambda: core.steric_local_variants(T, S, V, z_i, depth, pres))
print(json.dumps({"tag": sys.argv[1] if len(sys.argv) > 1 else "", "variants_ms": round(ms, 3), "gpts": round(pts / ms / 1e6, 1)}), flush=True)
