"""BASELINE config 4 at full length through the PUBLIC call, on one GPU, with bounded host memory.

    python tools/config4_public.py [--steps 365] [--block 1] [--check 3]

``momlevel_b200.steric(dset, domain="global")`` on a Dataset whose ``thetao`` / ``so`` exist only as blocks along
time -- what ``xr.open_mfdataset(..., chunks={"time": 1})`` gives (example.ipynb cell 4).  The 365-day OM4p125 series
is 1.41 TB; here every block is produced on demand (the synthetic generator stands in for the file reader: a block is
generated on the device and handed over as a plain numpy array, i.e. pageable host memory), streamed through
``core.HostStream`` and dropped.  Reports the wall time, the peak resident set size of the process (it has to stay
near two blocks plus the library's pinned staging, not near 1.41 TB) and compares a few steps of the mass series with
the device-resident call bit for bit.
"""

import argparse
import json
import pathlib
import resource
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import momlevel_b200 as ml  # noqa: E402
from momlevel_b200 import core, synth  # noqa: E402
from momlevel_b200.labeled import ChunkedArray, DataArray, Dataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=365)
    ap.add_argument("--block", type=int, default=1)
    ap.add_argument("--check", type=int, default=3, help="blocks whose masses are compared with the resident call")
    ap.add_argument("--grid", default="om4p125")
    args = ap.parse_args()
    _, nz, ny, nx = synth.CONFIGS[args.grid]
    nt = args.steps
    dev = torch.device("cuda", 0)
    grid = synth.make_grid(nz, ny, nx, seed=11, device=dev)
    chunks = tuple([args.block] * (nt // args.block) + ([nt % args.block] if nt % args.block else []))
    produced = {"thetao": 0, "so": 0}
    cache = {}

    def source(which):
        def blocks():
            t = 0
            for n in chunks:
                key = (t, n)
                if key not in cache:  # T and S of a block come out of one generator call
                    T, S, _ = synth.make_fields(grid, n, seed=55, dtype=torch.float32, t_first=t)
                    cache.clear()
                    cache[key] = {"thetao": T.cpu().numpy(), "so": S.cpu().numpy()}
                    del T, S
                produced[which] += 1
                yield cache[key].pop(which)
                t += n

        return blocks

    _, _, V = synth.make_fields(grid, 1, seed=55, dtype=torch.float32, t_first=0)
    dims = ("time", "z_l", "yh", "xh")
    ds = Dataset()
    ds["time"] = DataArray(np.arange(nt, dtype=np.float64), ("time",))
    for k in ("z_l", "z_i"):
        ds[k] = DataArray(grid[k].cpu().numpy(), (k,))
    ds["yh"] = DataArray(np.arange(ny, dtype=np.float64), ("yh",))
    ds["xh"] = DataArray(np.arange(nx, dtype=np.float64), ("xh",))
    shape = (nt, nz, ny, nx)
    ds["thetao"] = DataArray(ChunkedArray(shape, np.float32, chunks, source("thetao")), dims)
    ds["so"] = DataArray(ChunkedArray(shape, np.float32, chunks, source("so")), dims)
    Vh = V.cpu().numpy()
    ds["volcello"] = DataArray(ChunkedArray(shape, np.float32, chunks, lambda: iter([Vh[None]] * len(chunks))), dims)
    ds["areacello"] = DataArray(grid["areacello"].cpu().numpy(), ("yh", "xh"))
    ds["deptho"] = DataArray(grid["deptho"].cpu().numpy(), ("yh", "xh"))
    del V
    torch.cuda.empty_cache()

    rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    t0 = time.perf_counter()
    result, reference = ml.steric(ds, domain="global")
    wall = time.perf_counter() - t0
    rss1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    eta = result["steric"].values
    h2d, packed = core.host_last_transfer()

    # a few blocks again, resident on the device: the masses must agree bit for bit
    volo, rhoga = float(reference["volo"]), float(reference["rhoga"])
    pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
    worst = 0.0
    checked = []
    Vd = torch.from_numpy(Vh).to(dev)
    for i in np.linspace(0, len(chunks) - 1, max(1, args.check)).astype(int).tolist():
        t = int(sum(chunks[:i]))
        T, S, _ = synth.make_fields(grid, chunks[i], seed=55, dtype=torch.float32, t_first=t)
        m = core.steric_global(T, S, Vd, pres).cpu().numpy()
        want = (volo / float(np.nansum(ds["areacello"].values))) * np.log(rhoga / (m / volo))
        worst = max(worst, float(np.max(np.abs(want - eta[t: t + chunks[i]]))))
        checked.append(t)
        del T, S
    step_bytes = 2 * nz * ny * nx * 4
    print(json.dumps({
        "workload": f"{args.grid} {nx}x{ny}x{nz}, {nt} steps in blocks of {args.block}, steric(dset, domain='global') on a "
                    "Dataset of chunked (dask-like) fields in pageable host memory",
        "wall_s": wall, "points": nt * nz * ny * nx, "gpts": nt * nz * ny * nx / wall / 1e9,
        "field_bytes_total": nt * step_bytes, "block_bytes": args.block * step_bytes,
        "h2d_bytes": h2d, "level_rows_sent_packed": packed,
        "peak_rss_bytes_before": rss0 * 1024, "peak_rss_bytes_after": rss1 * 1024,
        "rss_growth_in_blocks": (rss1 - rss0) * 1024 / (args.block * step_bytes),
        "blocks_produced": produced, "eta_first_m": float(eta[0]), "eta_last_m": float(eta[-1]),
        "steps_checked_against_resident_call": checked, "max_abs_diff_vs_resident_m": worst,
    }), flush=True)


if __name__ == "__main__":
    main()
