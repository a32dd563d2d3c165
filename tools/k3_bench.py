"""Quick A/B harness for the fused kernels on the OM4p25 workload (kernel time only, CUDA events)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from momlevel_b200 import core, synth

nt, nz, ny, nx = 12, 75, 1080, 1440
grid = synth.make_grid(nz, ny, nx, seed=123, device="cuda")
T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
pres = grid["z_l"] * 1e4 + 101325.0
z_i, depth = grid["z_i"], grid["deptho"]
pts = nt * nz * ny * nx

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

eta, rho_ref, sums = core.steric_local_selfref(T, S, V, z_i, depth, pres)
res = {"tag": sys.argv[1] if len(sys.argv) > 1 else ""}
res["selfref"] = pts / timed(lambda: core.steric_local_selfref(T, S, V, z_i, depth, pres)) / 1e6
res["local"] = pts / timed(lambda: core.steric_local(T, S, rho_ref, V, z_i, depth, pres)) / 1e6
res["global"] = pts / timed(lambda: core.steric_global(T, S, V, pres)) / 1e6
if "--all" in sys.argv:
    res["thermo"] = pts / timed(lambda: core.steric_local(T, S[0], rho_ref, V, z_i, depth, pres, s_bcast=True)) / 1e6
    res["halo"] = pts / timed(lambda: core.steric_local(T[0], S, rho_ref, V, z_i, depth, pres, t_bcast=True)) / 1e6
    res["linear"] = pts / timed(lambda: core.steric_local(T, S, rho_ref, V, z_i, depth, pres, eos="linear")) / 1e6
print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
