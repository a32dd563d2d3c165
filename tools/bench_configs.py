"""Throughput of the remaining BASELINE.json configs (3, 4, 5); bench.py covers config 2.

    python tools/bench_configs.py [3] [4] [5]          (1 GPU)
    torchrun --nproc-per-node N tools/bench_configs.py 3 4     (members / time steps sharded over N ranks)

Prints one JSON line per config (rank 0).  Kernel time is measured with CUDA events around the
library calls only; generating the synthetic fields is not timed.  Datasets that do not fit in
HBM (config 4: 3.9 GB per step x 365) are streamed through a window that is regenerated in place.
"""

import json
import os
import pathlib
import sys

import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from momlevel_b200 import core, synth  # noqa: E402
from momlevel_b200 import distributed as mld  # noqa: E402

PEAK = 6547.5
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except (OSError, ValueError, KeyError):
    pass


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn):
    a, b = ev(), ev()
    a.record()
    out = fn()
    b.record()
    return out, (a, b)


def max_over_ranks(ms, dev, world):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def config3(rank, world, dev):
    """SPEAR 1 deg, 30 members x 120 months, local steric; members sharded over ranks."""
    nt, nz, ny, nx = synth.CONFIGS["spear1deg"]
    members = mld.assign_members(30, world, rank)
    if world == 1:
        members = members[:4]  # what rank 0 of an 8-GPU run owns
    grid = synth.make_grid(nz, ny, nx, seed=7, device=dev)
    pres = grid["z_l"] * 1.0e4 + 101325.0
    z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
    # all members of the rank are resident (4 x 8.3 GB) -- generating them is not timed
    fields = [synth.make_fields(grid, nt, seed=1000 + m, dtype=torch.float32) for m in members]
    n_streams = int(os.environ.get("ML_MEMBER_STREAMS", "4"))
    mld.steric_local_members(fields, z_i, depth, pres, n_streams=n_streams)  # warm-up
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(3):
        a, b = ev(), ev()
        a.record()
        res = mld.steric_local_members(fields, z_i, depth, pres, n_streams=n_streams)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
        del res
    del fields
    ms = max_over_ranks(best, dev, world)
    n_members = 30 if world > 1 else len(members)
    pts = n_members * nt * nz * ny * nx
    # algorithmic bytes: T,S once, volcello(t=0) once, rho_ref written once, deptho + eta (the later chunks' reads of
    # rho_ref / volcello are served by L2 because the chunks of a tile run together)
    per_member_bytes = nt * nz * ny * nx * 8 + nz * ny * nx * (4 + 8) + ny * nx * 8 * (nt + 1)
    return {"config": 3, "workload": f"SPEAR 1deg {nx}x{ny}x{nz}, {n_members} members x {nt} months, local steric, Wright",
            "n_gpus": world, "members_per_rank_max": len(members), "member_streams": n_streams, "kernel_ms_max_rank": ms,
            "value": pts / (ms * 1e-3), "unit": "grid-points/s",
            "hbm_frac": per_member_bytes * len(members) / (ms * 1e-3) / 1e9 / PEAK}


def config4(rank, world, dev, window=12):
    """OM4p125 daily x 365, global steric series, time-sharded; each rank streams its block."""
    nt, nz, ny, nx = synth.CONFIGS["om4p125"]
    lo, hi = mld.shard_range(nt, world, rank)
    if world == 1:
        lo, hi = 0, 46  # what rank 0 of an 8-GPU run owns
    grid = synth.make_grid(nz, ny, nx, seed=11, device=dev)
    pres = grid["z_l"] * 1.0e4 + 101325.0
    # every rank regenerates step 0 itself (no broadcast of the 5.8 GB reference state)
    T0, S0, V = synth.make_fields(grid, 1, seed=55, dtype=torch.float32, t_first=0)
    rho_ref, sums = core.reference_state(T0[0], S0[0], V, pres)
    volo, masso_ref = (float(x) for x in sums.cpu())
    del T0, S0, rho_ref
    torch.cuda.empty_cache()
    pairs, parts = [], []
    first = True
    for t in range(lo, hi, window):
        n = min(window, hi - t)
        T, S, _ = synth.make_fields(grid, n, seed=55, dtype=torch.float32, t_first=t)
        if first:
            core.steric_global(T, S, V, pres)  # warm-up
            torch.cuda.synchronize()
            first = False
        m, p = timed(lambda: core.steric_global(T, S, V, pres))
        pairs.append(p)
        parts.append(m)
        torch.cuda.synchronize()
        del T, S
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs), dev, world)
    masso_local = torch.cat(parts)
    g0, g1 = ev(), ev()
    g0.record()
    masso = mld.gather_series(masso_local, nt) if world > 1 else masso_local
    g1.record()
    torch.cuda.synchronize()
    eta, href = mld.global_sea_level(masso.cpu().numpy(), volo, masso_ref / volo, float(torch.nansum(grid["areacello"])))
    n_steps = nt if world > 1 else hi - lo
    pts = n_steps * nz * ny * nx
    steps_rank = hi - lo
    bytes_rank = steps_rank * nz * ny * nx * 8 + (steps_rank + window - 1) // window * nz * ny * nx * 4
    return {"config": 4, "workload": f"OM4p125 {nx}x{ny}x{nz}, {n_steps} daily steps, global steric series, Wright, "
                                     f"streamed in {window}-step windows",
            "n_gpus": world, "steps_per_rank_max": steps_rank, "kernel_ms_max_rank": ms,
            "gather_ms": g0.elapsed_time(g1), "value": pts / (ms * 1e-3), "unit": "grid-points/s",
            "hbm_frac": bytes_rank / (ms * 1e-3) / 1e9 / PEAK, "eta_first": float(eta[0]), "eta_last": float(eta[-1]),
            "reference_height_m": float(href)}


def config5(rank, world, dev):
    """OM4p25 x 12, linear EOS local steric + Flament spiciness over the 4-D fields."""
    nt, nz, ny, nx = synth.CONFIGS["om4p25"]
    grid = synth.make_grid(nz, ny, nx, seed=123, device=dev)
    pres = grid["z_l"] * 1.0e4 + 101325.0
    T, S, V = synth.make_fields(grid, nt, seed=123 + rank, dtype=torch.float32)
    pts = nt * nz * ny * nx
    out = {"config": 5, "workload": f"OM4p25 {nx}x{ny}x{nz}, {nt} steps, linear EOS local steric + Flament spiciness",
           "n_gpus": world, "unit": "grid-points/s"}
    core.steric_local_selfref(T, S, V, grid["z_i"], grid["deptho"], pres, eos="linear")
    torch.cuda.synchronize()
    _, (a, b) = timed(lambda: core.steric_local_selfref(T, S, V, grid["z_i"], grid["deptho"], pres, eos="linear"))
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(b), dev, world)
    N = nz * ny * nx
    out["steric_linear_ms"] = ms
    out["steric_linear_value"] = world * pts / (ms * 1e-3)
    out["steric_linear_hbm_frac"] = (nt * N * 8 + N * 12 + ny * nx * 8 * (nt + 1)) / (ms * 1e-3) / 1e9 / PEAK
    # spice: fp64 output as large as both inputs together -> 6 steps at a time keeps HBM use bounded
    half = nt // 2
    res = core.flament_spice(T[:half], S[:half])
    torch.cuda.synchronize()
    del res
    tot = 0.0
    for h in range(2):
        res, (a, b) = timed(lambda: core.flament_spice(T[h * half:(h + 1) * half], S[h * half:(h + 1) * half]))
        torch.cuda.synchronize()
        del res  # the 5.6 GB result goes back to the caching allocator before the next half
        tot += a.elapsed_time(b)
    ms = max_over_ranks(tot, dev, world)
    out["spice_ms"] = ms
    out["spice_value"] = world * pts / (ms * 1e-3)
    out["spice_hbm_frac"] = pts * 16 / (ms * 1e-3) / 1e9 / PEAK
    return out


def main():
    which = [int(a) for a in sys.argv[1:]] or [3, 4, 5]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for c in which:
        res = {3: config3, 4: config4, 5: config5}[c](rank, world, dev)
        torch.cuda.empty_cache()
        if rank == 0:
            print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
