import pathlib, sys, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from momlevel_b200 import core, synth
nt, nz, ny, nx = 6, 75, 1080, 1440
grid = synth.make_grid(nz, ny, nx, seed=123, device="cuda")
T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
pres = (grid["z_l"] * 1e4 + 101325.0).contiguous()
pts = nt * nz * ny * nx
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); del r
        best = min(best, a.elapsed_time(b))
    return best
res = {"tag": sys.argv[1] if len(sys.argv) > 1 else ""}
res["spice"] = round(pts / timed(lambda: core.flament_spice(T, S)) / 1e6, 1)
res["eos_density"] = round(pts / timed(lambda: core.eos_eval("Wright", "density", T, S, pres, z_axis=1)) / 1e6, 1)
res["calc_n2"] = round(pts / timed(lambda: core.calc_n2(T, S, grid["z_l"])) / 1e6, 1)
rho_ref, _ = core.reference_state(T[0], S[0], V, pres)
res["delta_rho"] = round(pts / timed(lambda: core.delta_rho(T, S, rho_ref, V, pres)) / 1e6, 1)
res["reference_state"] = round(nz * ny * nx / timed(lambda: core.reference_state(T[0], S[0], V, pres)) / 1e6, 1)
print(json.dumps(res), flush=True)
