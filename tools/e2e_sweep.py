"""A/B of the host path on an OM4p25 year: how level rows cross PCIe (as they are / packed / balanced), how many
host threads pack, how many steps a window holds.  One JSON line per setting.

    python tools/e2e_sweep.py [reps] [ring]
"""

import json
import pathlib
import sys
import time

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

from momlevel_b200 import core, synth  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    quick = len(sys.argv) > 2 and sys.argv[2] == "ring"  # only the staging-ring A/B
    nt, nz, ny, nx = synth.CONFIGS["om4p25"]
    dev = torch.device("cuda", 0)
    grid = synth.make_grid(nz, ny, nx, seed=123, device=dev)
    T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
    pres = (grid["z_l"] * 1.0e4 + 101325.0).cpu().numpy()
    z_i, depth = grid["z_i"].cpu().numpy(), grid["deptho"].cpu().numpy()
    Th, Sh, Vh = (torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x) for x in (T, S, V))
    eta_h = torch.empty((nt, ny, nx), dtype=torch.float64, pin_memory=True)
    del T, S, V
    torch.cuda.synchronize()
    points = nt * nz * ny * nx
    first = None
    settings = [(0, 0, 1), (1, 0, 1), (2, 0, 1), (1, 4, 1), (1, 6, 1), (1, 8, 1), (1, 10, 1), (1, 12, 1), (2, 8, 1),
                (2, 12, 1), (1, 0, 2), (1, 0, 3), (1, 0, 4), (1, 0, 6), (1, 0, 12), (1, 8, 3), (1, 0, 1)]
    if quick:
        settings = [(1, 8, 1), (3, 8, 1), (4, 8, 1), (4, 0, 1), (4, 12, 1), (4, 6, 1), (1, 0, 1), (4, 15, 1), (2, 12, 1)]
    for mode, threads, spw in settings:
        core.host_packing(mode, threads)
        run = lambda: core.steric_local_host(Th, Sh, Vh, z_i, depth, pres, steps_per_window=spw, eta_out=eta_h)  # noqa: E731
        run()
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        nbytes, frac = core.host_last_transfer()
        if first is None:
            first = eta_h.clone()
        same = bool(torch.equal(first.view(torch.int64), eta_h.view(torch.int64)))
        print(json.dumps({"mode": mode, "threads": threads or "default", "steps_per_window": spw, "ms": round(best * 1e3, 2),
                          "gpts": round(points / best / 1e9, 3), "h2d_gb": round(nbytes / 1e9, 3),
                          "rows_packed": round(frac, 3), "pcie_gbs": round(nbytes / best / 1e9, 1),
                          "host_ms": {k: round(v, 1) for k, v in core.host_last_timings().items()},
                          "bit_identical": same}), flush=True)
    core.host_packing(1, 0)


if __name__ == "__main__":
    main()
