"""Elementwise / sweep kernels on OM4p25 x 6 steps: G points/s and share of the HBM copy bandwidth.

    python tools/spice_probe.py [tag]

One JSON line for the default kernels (ring-staged streaming kernels of csrc/ml_stream.cu where the fields
suit them) and one with ``ml_set_force_direct(1)`` -- the plain-load kernels -- for the A/B.
"""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from momlevel_b200 import core, synth  # noqa: E402

PEAK = 6547.5
try:
    PEAK = float(json.loads((pathlib.Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except (OSError, ValueError, KeyError):
    pass
nt, nz, ny, nx = 6, 75, 1080, 1440
grid = synth.make_grid(nz, ny, nx, seed=123, device="cuda")
T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
pres = (grid["z_l"] * 1e4 + 101325.0).contiguous()
pts = nt * nz * ny * nx


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        del r
        best = min(best, a.elapsed_time(b))
    return best


def entry(points, bytes_per_point, ms):
    return {"gpts": round(points / ms / 1e6, 1), "ms": round(ms, 4),
            "hbm_frac": round(points * bytes_per_point / (ms * 1e-3) / 1e9 / PEAK, 3)}


rho_ref, _ = core.reference_state(T[0], S[0], V, pres)
for mode, name in ((0, "default"), (1, "plain-load kernels (ml_set_force_direct(1))")):
    core.force_direct(mode)
    res = {"tag": sys.argv[1] if len(sys.argv) > 1 else "", "kernels": name, "peak_gbs": PEAK}
    res["spice"] = entry(pts, 16, timed(lambda: core.flament_spice(T, S)))
    res["eos_density"] = entry(pts, 16, timed(lambda: core.eos_eval("Wright", "density", T, S, pres, z_axis=1)))
    res["linear_density"] = entry(pts, 16, timed(lambda: core.eos_eval("linear", "density", T, S, pres, z_axis=1)))
    res["calc_n2"] = entry(pts, 16, timed(lambda: core.calc_n2(T, S, grid["z_l"])))
    res["delta_rho"] = entry(pts, 16, timed(lambda: core.delta_rho(T, S, rho_ref, V, pres)))
    res["reference_state"] = entry(nz * ny * nx, 20, timed(lambda: core.reference_state(T[0], S[0], V, pres)))
    print(json.dumps(res), flush=True)
core.force_direct(0)
