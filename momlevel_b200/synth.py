"""Synthetic MOM6-shaped ocean states for tests and benchmarks (no datasets are fetchable).

Shapes and statistics follow SURVEY.md section 8(d): a z* grid whose layer thickness grows
from 2 m at the surface to ~250 m at 6500 m, bathymetry with ~30 % land in coherent blocks
(land columns and cells below the sea floor are NaN, as MOM6 writes them), temperature with
a depth trend, salinity near 35, and time steps that are the first step plus a small
perturbation so the density anomaly is realistic.  ``areacello`` is normalised to the real
ocean area so ``validate_dataset`` accepts it (util.py:669-694).

Everything is generated with ``torch`` on the requested device from explicit seeds; any
shard can regenerate any step (including the reference step 0) without communication.
"""

import numpy as np
import torch

from .labeled import DataArray, Dataset

__all__ = ["vertical_grid", "make_grid", "make_fields", "make_dataset", "dataset_from_fields", "CONFIGS"]

OCEAN_AREA = 3.6111092e14

# name -> (nt, nz, ny, nx); BASELINE.json configs 2-5
CONFIGS = {
    "om4p25": (12, 75, 1080, 1440),
    "spear1deg": (120, 75, 320, 360),
    "om4p125": (365, 75, 2240, 2880),
}


def vertical_grid(nz=75, depth=6500.0):
    """``z_i`` (nz+1 interfaces, 0 -> depth) and ``z_l`` (mid-points), fp64 numpy."""
    k = np.arange(nz, dtype=np.float64)
    dz = 2.0 + 248.0 * (k / max(nz - 1, 1)) ** 2.2
    z_i = np.concatenate([[0.0], np.cumsum(dz)])
    z_i *= depth / z_i[-1]
    return z_i, 0.5 * (z_i[1:] + z_i[:-1])


def make_grid(nz, ny, nx, seed=123, device="cpu", land_fraction=0.3):
    """Static grid: ``z_i, z_l, deptho[ny,nx] (NaN = land), areacello[ny,nx]``, all fp64 torch."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    z_i, z_l = vertical_grid(nz)
    # coherent land blocks: a coarse random field upsampled by repetition
    by, bx = max(ny // 16, 1), max(nx // 16, 1)
    coarse = torch.rand((by, bx), generator=g)
    land = coarse < land_fraction
    ry, rx = -(-ny // by), -(-nx // bx)
    land = land.repeat_interleave(ry, 0).repeat_interleave(rx, 1)[:ny, :nx].clone()
    if by * bx < 8:  # tiny grids: scattered land points instead of blocks, never all land
        land = torch.rand((ny, nx), generator=g) < land_fraction
        land[ny // 2, nx // 2] = False
    depth = torch.rand((ny, nx), generator=g, dtype=torch.float64) * (z_i[-1] - 10.0) + 10.0
    depth = torch.where(land, torch.full_like(depth, float("nan")), depth)
    area = 0.5 + torch.rand((ny, nx), generator=g, dtype=torch.float64)
    area = torch.where(land, torch.zeros_like(area), area)
    area = area / area.sum() * OCEAN_AREA
    return {
        "z_i": torch.from_numpy(z_i).to(device),
        "z_l": torch.from_numpy(z_l).to(device),
        "deptho": depth.to(device),
        "areacello": area.to(device),
    }


def make_fields(grid, nt, seed=123, device=None, dtype=torch.float32, t_first=0):
    """``thetao, so [nt,nz,ny,nx]`` and ``volcello [nz,ny,nx]`` of ``dtype`` on ``device``.

    Step ``t`` depends only on ``(seed, t_first + t)``: the mean state (step-independent,
    seeded by ``seed``) plus a per-step perturbation seeded by ``(seed, t)``; step 0 has no
    perturbation and is the reference state.
    """
    device = grid["deptho"].device if device is None else torch.device(device)
    z_i, z_l, depth = grid["z_i"].to(device), grid["z_l"].to(device), grid["deptho"].to(device)
    nz, (ny, nx) = z_l.numel(), depth.shape
    g = torch.Generator(device=device).manual_seed(seed)
    zfac = torch.exp(-z_l / 1000.0).view(nz, 1, 1).to(torch.float32)
    Tm = (2.0 + 18.0 * zfac + 3.0 * torch.randn((nz, ny, nx), generator=g, device=device)).clamp_(-2.0, 32.0)
    Sm = (35.0 + 1.0 * torch.randn((nz, ny, nx), generator=g, device=device)).clamp_(30.0, 40.0)
    # MOM6-style masking: land columns and cells entirely below the sea floor are missing
    dry = torch.isnan(depth).unsqueeze(0) | (z_i[:-1].view(nz, 1, 1) >= torch.nan_to_num(depth, nan=0.0).unsqueeze(0))
    nan32 = torch.tensor(float("nan"), device=device)
    Tm = torch.where(dry, nan32, Tm)
    Sm = torch.where(dry, nan32, Sm)
    dz = torch.minimum((torch.nan_to_num(depth, nan=0.0).unsqueeze(0) - z_i[:-1].view(nz, 1, 1)).clamp_min(0.0),
                       (z_i[1:] - z_i[:-1]).view(nz, 1, 1))
    V = torch.where(dry, torch.tensor(float("nan"), device=device, dtype=torch.float64),
                    dz * grid["areacello"].to(device).unsqueeze(0))
    T = torch.empty((nt, nz, ny, nx), dtype=dtype, device=device)
    S = torch.empty((nt, nz, ny, nx), dtype=dtype, device=device)
    for t in range(nt):
        tt = t_first + t
        if tt == 0:
            T[t], S[t] = Tm.to(dtype), Sm.to(dtype)
            continue
        gt = torch.Generator(device=device).manual_seed(seed * 1000003 + tt)
        T[t] = (Tm + 0.5 * zfac * torch.randn((nz, ny, nx), generator=gt, device=device)).to(dtype)
        S[t] = (Sm + 0.1 * zfac * torch.randn((nz, ny, nx), generator=gt, device=device)).to(dtype)
    return T, S, V.to(dtype)


def dataset_from_fields(grid, T, S, V):
    """Wrap resident fields in a labelled Dataset shaped like MOM6 output (``volcello`` is a time-expanded view)."""
    nt, nz, ny, nx = T.shape
    dims = ("time", "z_l", "yh", "xh")
    ds = Dataset()
    ds["time"] = DataArray(np.arange(nt, dtype=np.float64), ("time",))
    ds["z_l"] = DataArray(grid["z_l"], ("z_l",))
    ds["z_i"] = DataArray(grid["z_i"], ("z_i",))
    ds["yh"] = DataArray(np.arange(ny, dtype=np.float64), ("yh",))
    ds["xh"] = DataArray(np.arange(nx, dtype=np.float64), ("xh",))
    ds["thetao"] = DataArray(T, dims)
    ds["so"] = DataArray(S, dims)
    ds["volcello"] = DataArray(V.unsqueeze(0).expand(nt, -1, -1, -1), dims)
    ds["deptho"] = DataArray(grid["deptho"], ("yh", "xh"))
    ds["areacello"] = DataArray(grid["areacello"], ("yh", "xh"))
    return ds


def make_dataset(nt, nz, ny, nx, seed=123, device="cpu", dtype=torch.float32):
    """A labelled Dataset shaped like MOM6 output."""
    grid = make_grid(nz, ny, nx, seed=seed, device=device)
    T, S, V = make_fields(grid, nt, seed=seed, dtype=dtype)
    return dataset_from_fields(grid, T, S, V)
