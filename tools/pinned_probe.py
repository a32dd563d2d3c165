"""Single-height launches on OM4p25 x 12 by CUDA events (best of 7): steric, thermosteric, halosteric, global, and the
pinned variants of the global series.  MOMLEVEL_B200_LIB selects an experiment build.  One JSON line."""
import json
import os
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from momlevel_b200 import core, synth  # noqa: E402


def best_ms(fn, n=7):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    nt, nz, ny, nx = synth.CONFIGS["om4p25"]
    dev = torch.device("cuda", 0)
    grid = synth.make_grid(nz, ny, nx, seed=123, device=dev)
    T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
    pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
    z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
    T0, S0 = T[0].contiguous(), S[0].contiguous()
    eta, rho, _ = core.steric_local_selfref(T, S, V, z_i, depth, pres, want_rho_ref=True)
    out = {"lib": os.environ.get("MOMLEVEL_B200_LIB", "default")}
    buf = torch.empty_like(eta)
    legs = {
        "selfref": lambda: core.steric_local_selfref(T, S, V, z_i, depth, pres, want_rho_ref=False),
        "local": lambda: core.steric_local(T, S, rho, V, z_i, depth, pres, want_delta_rho=False, eta_out=buf),
        "thermo": lambda: core.steric_local(T, S0, rho, V, z_i, depth, pres, want_delta_rho=False, s_bcast=True, eta_out=buf),
        "halo": lambda: core.steric_local(T0, S, rho, V, z_i, depth, pres, want_delta_rho=False, t_bcast=True, eta_out=buf),
        "global": lambda: core.steric_global(T, S, V, pres),
        "global_thermo": lambda: core.steric_global(T, S0, V, pres, s_bcast=True),
        "global_halo": lambda: core.steric_global(T0, S, V, pres, t_bcast=True),
    }
    for k, fn in legs.items():
        out[k] = round(best_ms(fn), 4)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
