"""BASELINE.json configs 3, 4 and 5 (bench.py's headline is config 2), each with an in-run oracle check.

    python tools/bench_configs.py [3] [4] [5]                          (1 GPU)
    torchrun --nproc-per-node N tools/bench_configs.py 3 4 5           (N ranks)

``bench.py`` imports this module and puts what ``run_configs`` returns under ``extras.configs`` of its
JSON line, so the driver's 1/2/4/8-GPU runs time the north-star splits themselves:

* config 3 -- 30 SPEAR members x 120 months, local steric: the (member, 12-step block) list is cut into
  contiguous shares (``distributed.assign_member_blocks``), STRONG scaling, no collective;
* config 4 -- 365 daily OM4p125 steps, global steric series: the time axis is cut into contiguous blocks,
  each rank streams its block through 12-step windows that are regenerated in place (1.41 TB does not fit),
  and the ONE collective of the design -- ``gather_series`` on a warmed NCCL communicator -- sits inside the
  event pair of the last window.  STRONG scaling;
* config 5 -- OM4p25 x 12, linear EOS local steric + Flament spiciness, one batch per rank (weak, like config 2).

Kernel time is measured with CUDA events around the library calls only (generating synthetic fields is not
timed); ``ms`` is the max over ranks of each rank's summed event time.  Parity (rank 0): the oracle on
columns sampled from every part of the grid (first / last / edge tiles and a stride in between) for the
local configs, one whole step of the mass series for the global one.
"""

import json
import os
import pathlib
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from momlevel_b200 import core, synth  # noqa: E402
from momlevel_b200 import distributed as mld  # noqa: E402

UNIT = "grid-points/s"
TILE = 256  # columns per CTA of the TMA family (csrc/ml_tma.cu)


def hbm_peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, ValueError, KeyError):
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def ev():
    return torch.cuda.Event(enable_timing=True)


def max_over_ranks(x, dev, world):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def roofline(alg_bytes, ms):
    peak, src = hbm_peak()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "algorithmic_bytes": int(alg_bytes), "peak_source": src}


def n1_record(key):
    """The committed one-GPU record of the same leg (profiles/r02_bench_1gpu.json), for ``efficiency_vs_n1``."""
    try:
        rec = json.loads((ROOT / "profiles" / "r02_bench_1gpu.json").read_text())
        return float(rec["extras"]["configs"][key]["value"])
    except (OSError, ValueError, KeyError, TypeError):
        return None


def with_efficiency(out, key, world):
    base = n1_record(key)
    if base:
        out["efficiency_vs_n1"] = out["value"] / (world * base)  # value is the whole job's rate either way
        out["n1_value_used"] = base
        out["n1_source"] = "profiles/r02_bench_1gpu.json (committed one-GPU run of this leg)"
    return out


def sample_columns(ncol, n_stride=72):
    """Columns from every part of the grid: both ends of the first, second and last 256-column tile and a
    stride across everything in between -- not one slab from the middle."""
    tiles = (ncol + TILE - 1) // TILE
    picks = {0, 1, TILE - 1, TILE, TILE + 1, 2 * TILE - 1, (tiles - 1) * TILE - 1, (tiles - 1) * TILE,
             (tiles - 1) * TILE + 1, ncol - 2, ncol - 1, (tiles // 2) * TILE - 1, (tiles // 2) * TILE}
    picks |= {int(x) for x in np.linspace(0, ncol - 1, n_stride)}
    return np.array(sorted(p for p in picks if 0 <= p < ncol), dtype=np.int64)


def oracle_local_parity(T, S, V, ref, grid, eta, eos="Wright", variant="steric"):
    """max |eta - oracle| over sampled columns.  ``ref``: None (reference = step 0 of T, S) or ``(T0, S0)``."""
    from oracle import steric as osteric

    ncol = eta[0].numel()
    cols = sample_columns(ncol)
    idx = torch.as_tensor(cols, device=T.device)
    f64 = lambda x: x.cpu().numpy().astype(np.float64)  # noqa: E731 -- the parity definition: fp64 upcast
    Tc = f64(T.flatten(2)[:, :, idx])[:, :, None, :]
    Sc = f64(S.flatten(2)[:, :, idx])[:, :, None, :]
    Vc = f64(V.flatten(1)[:, idx])[None, :, None, :]
    depth = f64(grid["deptho"].flatten()[idx])[None, :]
    area = f64(grid["areacello"].flatten()[idx])[None, :]
    z_l, z_i = f64(grid["z_l"]), f64(grid["z_i"])
    if ref is None:
        oref = osteric.reference_state(Tc, Sc, Vc, area, z_l, eos=eos)
    else:
        T0 = f64(ref[0].flatten(1)[:, idx])[None, :, None, :]
        S0 = f64(ref[1].flatten(1)[:, idx])[None, :, None, :]
        oref = osteric.reference_state(T0, S0, Vc, area, z_l, eos=eos)
    want, _ = osteric.steric_local(Tc, Sc, z_l, z_i, depth, oref, eos=eos, variant=variant)
    got = f64(eta.flatten(1)[:, idx])[:, None, :]
    same_nan = bool(np.array_equal(np.isnan(got), np.isnan(want)))
    m = ~np.isnan(want)
    err = float(np.max(np.abs(got[m] - want[m]))) if m.any() else 0.0
    return {"max_abs_err_m": err, "nan_pattern_equal": same_nan, "columns": int(cols.size),
            "wet_columns": int(m[0].sum()), "steps": int(T.shape[0]),
            "sample": "both ends of the first / second / middle / last 256-column tile + 72 strided columns"}


# ------------------------------------------------------------------------------ config 3


def config3(rank, world, dev, reps=2, parity=True):
    """SPEAR 1 deg, 30 members x 120 months, local steric, Wright; member x time blocks sharded over ranks."""
    nt, nz, ny, nx = synth.CONFIGS["spear1deg"]
    n_members = 30
    N = nz * ny * nx
    pieces = mld.assign_member_blocks(n_members, nt, world, rank)
    grid = synth.make_grid(nz, ny, nx, seed=7, device=dev)
    pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
    z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
    n_streams = int(os.environ.get("ML_MEMBER_STREAMS", "4"))
    # resident batches of <= 36 GB of T and S (one rank of an 8-GPU run holds its whole share at once; one GPU
    # alone takes the 249 GB ensemble in eight batches)
    batches, cur, cur_bytes = [], [], 0
    for pc in pieces:
        b = (pc[2] - pc[1]) * N * 8
        if cur and cur_bytes + b > 36e9:
            batches.append(cur)
            cur, cur_bytes = [], 0
        cur.append(pc)
        cur_bytes += b
    if cur:
        batches.append(cur)

    totals = [0.0] * reps
    alg_bytes = 0
    checks = []
    warmed = False
    for bi, batch in enumerate(batches):
        fields = []
        for (m, t0, t1) in batch:
            T, S, V = synth.make_fields(grid, t1 - t0, seed=1000 + m, dtype=torch.float32, t_first=t0)
            ref = None
            if t0 > 0:  # the member's reference state lives in its step 0: one more step to load
                T0, S0, _ = synth.make_fields(grid, 1, seed=1000 + m, dtype=torch.float32, t_first=0)
                ref = (T0[0].contiguous(), S0[0].contiguous())
            fields.append((T, S, V, ref))
            steps = t1 - t0
            # T, S once; volcello(t=0) once; rho_ref written once (the later chunks of a tile re-read both from L2:
            # they are launched next to each other); deptho in, eta out; a block that does not start at step 0
            # also reads the step-0 slabs
            alg_bytes += steps * N * 8 + N * (4 + 8) + ny * nx * 8 * (steps + 1) + (N * 8 if ref is not None else 0)
        # the result tensors exist before the clock starts (a cudaMalloc inside the timed region is the allocator's
        # time, not the path's)
        res = [core.selfref_outputs(T, S) for T, S, _, _ in fields]
        if not warmed:
            mld.steric_local_pieces(fields, z_i, depth, pres, n_streams=n_streams, outs=res)
            warmed = True
        torch.cuda.synchronize()
        for r in range(reps):
            a, b = ev(), ev()
            a.record()
            mld.steric_local_pieces(fields, z_i, depth, pres, n_streams=n_streams, outs=res)
            b.record()
            torch.cuda.synchronize()
            totals[r] += a.elapsed_time(b)
        if parity and rank == 0:  # rank 0's first piece (starts at a step 0) and its last one (usually does not)
            todo = ([0] if bi == 0 else []) + ([len(batch) - 1] if bi == len(batches) - 1 else [])
            for k in sorted(set(todo)):
                T, S, V, ref = fields[k]
                chk = oracle_local_parity(T, S, V, ref, grid, res[k][0])
                chk["piece"] = {"member": batch[k][0], "steps": [batch[k][1], batch[k][2]]}
                checks.append(chk)
        del fields, res
    ms = max_over_ranks(sum(totals) / reps, dev, world)
    own_ms = sum(totals) / reps
    pts = n_members * nt * N
    blocks = sum(-(-(p[2] - p[1]) // 12) for p in pieces)
    out = {"workload": f"SPEAR-ocean 1deg {nx}x{ny}x{nz}, {n_members} members x {nt} months, local steric, Wright "
                       f"(BASELINE configs[2])",
           "points": pts, "ms": ms, "value": pts / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "strong",
           "sharding": "(member, 12-step block) list cut into contiguous shares, no collective; a block that does not "
                       "start at its member's step 0 loads that step and evaluates the reference density first",
           "blocks_on_rank0": blocks, "pieces_on_rank0": len(pieces), "batches_on_rank0": len(batches),
           "member_streams": n_streams, "timed_passes": [t for t in totals],
           "roofline": roofline(alg_bytes, own_ms), "parity": checks}
    out["roofline"]["note"] = "rank 0's algorithmic bytes over rank 0's own event time"
    return with_efficiency(out, "3", world)


# ------------------------------------------------------------------------------ config 4


def _oracle_mass_of_step(T_h, S_h, V_h, pres_h, workers):
    """sum(rho * volcello) of ONE step by the numpy oracle (derived.py:435-438), y-slabs on a thread pool."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import eos as oeos

    nz, ny, nx = T_h.shape
    rows = 32

    def slab(y0):
        ys = slice(y0, min(ny, y0 + rows))
        rho = oeos.density("Wright", T_h[:, ys].astype(np.float64), S_h[:, ys].astype(np.float64),
                           pres_h[:, None, None])
        return float(np.nansum(rho * V_h[:, ys].astype(np.float64)))

    with ThreadPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(slab, range(0, ny, rows)))
    return float(np.sum(np.array(parts, dtype=np.float64)))


def config4(rank, world, dev, window=12, parity=True, nt_override=None):
    """OM4p125 daily x 365, global steric series, time-sharded; windows regenerated in place; gather timed."""
    nt, nz, ny, nx = synth.CONFIGS["om4p125"]
    if nt_override:
        nt = int(nt_override)
    N = nz * ny * nx
    # rank 0 owns step 0 and sums the reference volume on top of its steps: it takes a short block when there is one
    lo, hi = mld.shard_range(nt, world, rank, light_first=True)
    grid = synth.make_grid(nz, ny, nx, seed=11, device=dev)
    pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
    area_sum = float(torch.nansum(grid["areacello"]))
    starts = list(range(lo, hi, window))
    total_ms, gather_ms = 0.0, 0.0
    parts, ref_sums = [], None
    check = None
    eta = href = None
    alg_bytes = 0
    # warm the communicator (the first NCCL call sets up channels and costs milliseconds) with the same
    # message shape as the timed gather
    if world > 1:
        dummy = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
        mld.gather_series(dummy, nt, extra=torch.zeros(2, dtype=torch.float64, device=dev), light_first=True)
        mld.gather_series(dummy, nt, extra=torch.zeros(2, dtype=torch.float64, device=dev), light_first=True)
    for wi, t in enumerate(starts):
        n = min(window, hi - t)
        T, S, V = synth.make_fields(grid, n, seed=55, dtype=torch.float32, t_first=t)
        if wi == 0:  # warm-up, untimed
            core.steric_global(T, S, V, pres)
            if t == 0:
                core.weighted_nansum(V)
        last = wi == len(starts) - 1
        if last:
            barrier(world)  # the gather must not be charged for ranks that are still generating their window
        torch.cuda.synchronize()
        a, b, g = ev(), ev(), ev()
        a.record()
        parts.append(core.steric_global(T, S, V, pres))
        alg_bytes += n * N * 8 + N * 4
        if t == 0:
            # The rank that owns step 0 owns the scalars of the reference state (reference.py:74-80): masso_ref IS
            # the mass of step 0, which the series holds already (calc_masso(rho(t=0), volcello), the same sum), and
            # volo is one skipna sum over volcello (derived.py:787-789) -- no pass over T and S for the reference
            # density, which the global branch never reads (steric.py:134-142).
            ref_sums = torch.cat([core.weighted_nansum(V), parts[0][:1]])
            alg_bytes += N * 4
        if last:
            g.record()
            mine = ref_sums if ref_sums is not None else torch.zeros(2, dtype=torch.float64, device=dev)
            series, extras = mld.gather_series(torch.cat(parts), nt, extra=mine, light_first=True)  # the one collective of the design
        b.record()
        torch.cuda.synchronize()
        total_ms += a.elapsed_time(b)
        if last:
            gather_ms = g.elapsed_time(b)
            # the read-back of 367 doubles and the ln formula (steric.py:136-142): host arithmetic, behind the clock
            eta, href = mld.finish_global_series(series, extras, area_sum)
        if parity and rank == 0 and wi == 0 and n > 1:
            k = 1  # a perturbed step (step 0 is the unperturbed mean state)
            t0 = time.perf_counter()
            workers = min(32, len(os.sched_getaffinity(0)))
            want = _oracle_mass_of_step(T[k].cpu().numpy(), S[k].cpu().numpy(), V.cpu().numpy(), pres.cpu().numpy(),
                                        workers)
            got = float(parts[0][k])
            check = {"step": lo + k, "masso_kernel_kg": got, "masso_oracle_kg": want,
                     "masso_rel_err": abs(got - want) / abs(want), "oracle_seconds": time.perf_counter() - t0,
                     "oracle_threads": workers,
                     "sample": "one whole step: every column of the 2880x2240x75 grid (the global sum has no columns "
                               "to sample)"}
        del T, S
    ms = max_over_ranks(total_ms, dev, world)
    gms = max_over_ranks(gather_ms, dev, world)
    pts = nt * N
    if check is not None and href is not None:
        # what the mass error means in metres: eta = href * ln(rhoga_ref * volo / M)  =>  d(eta) = href * dM / M
        check["eta_abs_err_m"] = float(href) * check["masso_rel_err"]
    out = {"workload": f"OM4p125 {nx}x{ny}x{nz}, {nt} daily steps, global steric series, Wright (BASELINE configs[3])",
           "points": pts, "ms": ms, "value": pts / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "strong",
           "sharding": f"time axis cut into contiguous blocks ({hi - lo} steps on rank 0), streamed through {window}-step "
                       "windows regenerated in place; reference state on the rank that owns step 0, its two scalars ride "
                       "in the all-gather of the mass series",
           "collective": "one all_gather_into_tensor of (steps per rank + 2) doubles, inside the event pair of the last "
                         "window; the read-back of the 367 doubles and the ln formula follow it on the host"
                         if world > 1 else "none (one rank)",
           "gather_ms": gms, "steps_on_rank0": hi - lo, "windows_on_rank0": len(starts),
           "roofline": roofline(alg_bytes, total_ms),
           "eta_first_m": float(eta[0]), "eta_last_m": float(eta[-1]), "reference_height_m": float(href),
           "parity": check}
    out["roofline"]["note"] = "rank 0's algorithmic bytes over rank 0's own event time, gather included"
    return with_efficiency(out, "4", world)


# ------------------------------------------------------------------------------ config 5


def config5(rank, world, dev, parity=True, reps=3):
    """OM4p25 x 12, linear EOS local steric + Flament spiciness over the 4-D fields; one batch per rank."""
    from oracle import spice as ospice

    nt, nz, ny, nx = synth.CONFIGS["om4p25"]
    N = nz * ny * nx
    pts = nt * N
    grid = synth.make_grid(nz, ny, nx, seed=123, device=dev)
    pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
    z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
    T, S, V = synth.make_fields(grid, nt, seed=123 + rank, dtype=torch.float32)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return out, a.elapsed_time(b) / reps

    (eta, _, _), ms_lin = timed(lambda: core.steric_local_selfref(T, S, V, z_i, depth, pres, eos="linear",
                                                                   want_rho_ref=False))
    lin_bytes = nt * N * 8 + N * 4 + ny * nx * 8 * (nt + 1)
    chk_lin = oracle_local_parity(T, S, V, None, grid, eta, eos="linear") if parity and rank == 0 else None
    del eta
    spice = torch.empty(T.shape, dtype=torch.float64, device=dev)  # 11.2 GB: allocated outside the timed region
    _, ms_sp = timed(lambda: core.flament_spice(T, S, out=spice))
    chk_sp = None
    if parity and rank == 0:
        idx = torch.arange(0, pts, max(1, pts // (1 << 20)), device=dev)  # exact integer stride
        Tc, Sc = (x.flatten()[idx].cpu().numpy().astype(np.float64) for x in (T, S))
        m = ~(np.isnan(Tc) | np.isnan(Sc))
        want = ospice.flament_spice(Tc[m], Sc[m])
        got = spice.flatten()[idx].cpu().numpy()
        nan_ok = bool(np.all(np.isnan(got[~m])))
        rel = float(np.max(np.abs(got[m] - want) / np.maximum(np.abs(want), 1.0)))
        chk_sp = {"max_rel_err": rel, "nan_pattern_equal": nan_ok, "points": int(idx.numel()),
                  "wet_points": int(m.sum()), "sample": "2^20 points strided over the whole 4-D field"}
    del spice
    torch.cuda.empty_cache()
    ms_lin_w = max_over_ranks(ms_lin, dev, world)
    ms_sp_w = max_over_ranks(ms_sp, dev, world)
    ms = ms_lin_w + ms_sp_w
    out = {"workload": f"OM4p25 {nx}x{ny}x{nz}, {nt} monthly steps, linear EOS local steric + Flament spiciness "
                       f"(BASELINE configs[4])",
           "points": world * pts, "ms": ms, "value": world * pts / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
           "scaling": "weak", "sharding": "one independent 12-step batch per rank, no collective",
           "steric_linear": {"ms": ms_lin_w, "value": world * pts / (ms_lin_w * 1e-3), "roofline": roofline(lin_bytes, ms_lin),
                             "parity": chk_lin},
           "spice": {"ms": ms_sp_w, "value": world * pts / (ms_sp_w * 1e-3), "roofline": roofline(pts * 16, ms_sp),
                     "parity": chk_sp},
           "roofline": roofline(lin_bytes + pts * 16, ms_lin + ms_sp)}
    return with_efficiency(out, "5", world)


LEGS = {"3": config3, "4": config4, "5": config5}


def run_configs(which, rank, world, dev):
    """``{"3": {...}, "4": {...}, "5": {...}}``; a leg that fails reports its error instead of taking the line down."""
    res = {}
    for key in which:
        t0 = time.perf_counter()
        try:
            res[key] = LEGS[key](rank, world, dev)
            res[key]["wall_s_including_data_generation"] = time.perf_counter() - t0
        except Exception as exc:  # noqa: BLE001
            if world > 1:
                raise  # the other ranks would wait forever in the next collective
            res[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001 -- a sticky device error: report what there is
            res.setdefault(key, {})["error_after"] = f"{type(exc).__name__}: {exc}"[:200]
            break
    return res


def main():
    which = [a for a in sys.argv[1:] if a in LEGS] or ["3", "4", "5"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run_configs(which, rank, world, dev)
    if rank == 0:
        print(json.dumps({"configs": res}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
