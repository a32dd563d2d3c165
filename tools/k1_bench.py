"""Timing of the elementwise / auxiliary kernels on OM4p25-sized fields (kernel time, CUDA events)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from momlevel_b200 import core, synth

nt, nz, ny, nx = 6, 75, 1080, 1440
grid = synth.make_grid(nz, ny, nx, seed=123, device="cuda")
T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
pres = grid["z_l"] * 1e4 + 101325.0
pts = nt * nz * ny * nx
N = nz * ny * nx

def timed(fn, n=3):
    r = fn(); torch.cuda.synchronize(); del r
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); del r
        best = min(best, a.elapsed_time(b))
    return best

res = {}
ms = timed(lambda: core.eos_eval("wright", "density", T, S, pres, z_axis=1)); res["eos_wright_density"] = (pts / ms / 1e6, pts * 16 / ms / 1e6)
ms = timed(lambda: core.eos_eval("linear", "density", T, S, None)); res["eos_linear_density"] = (pts / ms / 1e6, pts * 16 / ms / 1e6)
ms = timed(lambda: core.eos_eval("wright", "alpha", T, S, pres, z_axis=1)); res["eos_wright_alpha"] = (pts / ms / 1e6, pts * 16 / ms / 1e6)
ms = timed(lambda: core.flament_spice(T, S)); res["spice"] = (pts / ms / 1e6, pts * 16 / ms / 1e6)
ms = timed(lambda: core.reference_state(T[0], S[0], V, pres)); res["reference_state"] = (N / ms / 1e6, N * 20 / ms / 1e6)
ms = timed(lambda: core.calc_dz(grid["z_i"], grid["deptho"])); res["calc_dz"] = (N / ms / 1e6, N * 8 / ms / 1e6)
rho_ref, _ = core.reference_state(T[0], S[0], V, pres)
ms = timed(lambda: core.steric_local(T, S, rho_ref, V, grid["z_i"], grid["deptho"], pres, want_delta_rho=True)); res["local_with_delta_rho(direct)"] = (pts / ms / 1e6, pts * 16 / ms / 1e6)
ms = timed(lambda: core.delta_rho(T, S, rho_ref, V, pres)); res["delta_rho_entry"] = (pts / ms / 1e6, pts * 16.2 / ms / 1e6)
core.force_direct(True)
ms = timed(lambda: core.steric_local(T, S, rho_ref, V, grid["z_i"], grid["deptho"], pres)); res["local_direct_wright"] = (pts / ms / 1e6, pts * 8.2 / ms / 1e6)
ms = timed(lambda: core.steric_local(T, S, rho_ref, V, grid["z_i"], grid["deptho"], pres, eos="linear")); res["local_direct_linear"] = (pts / ms / 1e6, pts * 8.2 / ms / 1e6)
ms = timed(lambda: core.steric_global(T, S, V, pres)); res["global_direct_wright"] = (pts / ms / 1e6, pts * 8.1 / ms / 1e6)
T64, S64 = T[:2].double(), S[:2].double()
p2 = 2 * N
ms = timed(lambda: core.steric_local(T64, S64, rho_ref, V, grid["z_i"], grid["deptho"], pres)); res["local_direct_fp64_inputs"] = (p2 / ms / 1e6, p2 * 16.5 / ms / 1e6)
print(json.dumps({k: "%.0f Gpts/s %.0f GB/s" % v for k, v in res.items()}, indent=1))
