"""oracle.steric -- numpy restatement of momlevel's steric driver (TEST INFRASTRUCTURE).

Array-level: every field is a plain ``numpy`` array in MOM6 order ``[t][z][y][x]``
(x fastest).  xarray's behaviour that the reference relies on is written out:

* name-based broadcasting  -> explicit ``[None, :, None, None]`` reshapes
* ``DataArray.sum``        -> ``np.nansum`` (``skipna=True``; an all-NaN slice sums to 0.0)
* ``xr.where(c, x, nan)``  -> ``np.where``
* result is time-first (``steric.py:154,165``)

Reference lines followed:
  ``src/momlevel/steric.py:96``        pressure from depth
  ``src/momlevel/reference.py:54-83``  reference state
  ``src/momlevel/derived.py:295-323``  partial-cell dz
  ``src/momlevel/derived.py:435-438``  masso, ``:787-789`` volo, ``:661`` rhoga
  ``src/momlevel/steric.py:115-125``   variant select
  ``src/momlevel/steric.py:134-147``   global branch
  ``src/momlevel/steric.py:151-166``   local branch
"""

import numpy as np

from . import eos as _eos

__all__ = [
    "pressure_from_depth",
    "calc_dz",
    "reference_state",
    "steric_local",
    "steric_global",
    "VARIANTS",
]

VARIANTS = ("steric", "thermosteric", "halosteric")


def pressure_from_depth(z_l, patm=101325.0):
    """steric.py:96 / reference.py:54 -- 1 m of depth ~ 1 dbar = 1e4 Pa."""
    return (np.asarray(z_l, dtype=np.float64) * 1.0e4) + patm


def _pres3(z_l, patm):
    """``(z_l * 1e4) + patm`` broadcast by dimension name to ``[z][y][x]``: ``patm`` a scalar or a ``[y][x]`` field."""
    z = np.asarray(z_l, dtype=np.float64) * 1.0e4
    if np.ndim(patm) == 0:
        return (z + patm)[:, None, None]
    return z[:, None, None] + np.asarray(patm, dtype=np.float64)[None, :, :]


def calc_dz(z_l, z_i, deptho, top=0.0, bottom=None, fraction=False):
    """Partial-bottom-cell thickness, shape ``[z][y][x]`` (derived.py:249-325).

    The reference returns dims ``(y, x, z)``; only the axis order differs.
    """
    z_l = np.asarray(z_l, dtype=np.float64)
    z_i = np.asarray(z_i, dtype=np.float64)
    depth = np.asarray(deptho, dtype=np.float64)
    # derived.py:284-292
    assert bool(np.all(np.nan_to_num(depth, nan=0.0) >= 0)), "Depth values must all be positive-definite"
    assert bool(np.all(z_l >= 0)), "Vertical coordinate levels must all be positive-definite"
    assert bool(np.all(z_i >= 0)), "Vertical coordinate interfaces must all be positive-definite"
    # derived.py:295-298
    depth = np.where(np.isnan(depth), 0.0, depth)
    if bottom is not None:
        depth = np.minimum(depth, bottom)
    # derived.py:301-305
    ztop = z_i[:-1][:, None, None]
    zbot = z_i[1:][:, None, None]
    depth = depth[None, :, :]
    # derived.py:308-313
    dz_field = zbot - ztop
    part = depth - ztop
    part = np.where(part < 0.0, 0.0, part)
    result = np.minimum(part, dz_field)
    # derived.py:316-318
    part = zbot - top
    part = np.where(part < 0.0, 0.0, part)
    result = np.minimum(part, result)
    # derived.py:320-323
    if fraction:
        _dz_field = np.where(dz_field == 0, np.nan, dz_field)
        _dz_part = np.where(result == 0, np.nan, result)
        result = _dz_part / _dz_field
    return np.broadcast_to(result, (z_l.size,) + depth.shape[1:]).copy()


def reference_state(thetao, so, volcello, areacello, z_l, patm=101325.0, eos="Wright", time_index=0):
    """reference.py:15-85 -- returns a dict with the 8 reference variables."""
    pres = _pres3(z_l, patm)
    T0 = np.asarray(thetao)[time_index]
    S0 = np.asarray(so)[time_index]
    V0 = np.asarray(volcello)[time_index]
    rho0 = _eos.density(eos, T0, S0, pres)
    volo = np.nansum(V0)
    masso = np.nansum(rho0 * V0)
    return {
        "thetao": T0,
        "so": S0,
        "volcello": V0,
        "rho": rho0,
        "volo": volo,
        "masso": masso,
        "rhoga": masso / volo,
        "areacello": np.asarray(areacello),
    }


def _select(variant, thetao, so, reference):
    """steric.py:115-125."""
    if variant == "thermosteric":
        return np.asarray(thetao), reference["so"][None]
    if variant == "halosteric":
        return reference["thetao"][None], np.asarray(so)
    if variant == "steric":
        return np.asarray(thetao), np.asarray(so)
    raise ValueError(f"Unknown variant '{variant}' passed to `steric`")


def _rho(variant, thetao, so, reference, z_l, patm, eos):
    T, S = _select(variant, thetao, so, reference)
    pres = _pres3(z_l, patm)[None]
    rho = _eos.density(eos, T, S, pres)
    nt = max(np.asarray(thetao).shape[0], np.asarray(so).shape[0])
    return np.broadcast_to(rho, (nt,) + rho.shape[1:])


def steric_local(
    thetao, so, z_l, z_i, deptho, reference, rhozero=1035.0, patm=101325.0, eos="Wright", variant="steric"
):
    """Local branch (steric.py:150-166). Returns ``(eta[t,y,x], delta_rho[t,z,y,x])``."""
    rho = _rho(variant, thetao, so, reference, z_l, patm, eos)
    wet = ~np.isnan(reference["volcello"])
    delta_rho = np.where(wet[None], rho - reference["rho"][None], np.nan)
    dz = calc_dz(z_l, z_i, deptho)
    eta = (-1.0 / rhozero) * np.nansum(dz[None] * delta_rho, axis=1)
    eta = np.where(wet[0][None], eta, np.nan)
    return eta, delta_rho


def steric_global(thetao, so, z_l, reference, patm=101325.0, eos="Wright", variant="steric"):
    """Global branch (steric.py:134-147). Returns ``(eta[t], reference_height, masso[t])``."""
    rho = _rho(variant, thetao, so, reference, z_l, patm, eos)
    masso = np.nansum(rho * reference["volcello"][None], axis=(1, 2, 3))
    expansion = np.log(reference["rhoga"] / (masso / reference["volo"]))
    href = reference["volo"] / np.nansum(reference["areacello"])
    return href * expansion, href, masso
