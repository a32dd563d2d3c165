"""derived.py -- the ``momlevel.derived`` functions on the steric path, labelled arrays in/out.

Mirrors ``src/momlevel/derived.py``: ``calc_rho`` (:597-639), ``calc_dz`` (:249-325),
``calc_masso`` (:414-444), ``calc_volo`` (:769-795), ``calc_rhoga`` (:642-666),
``calc_spice`` (:669-711), plus ``calc_alpha`` / ``calc_beta`` (:74-159) and ``calc_pdens``
(:447-486) which reuse the same elementwise kernel, and the column diagnostics that share the
vertical sweep: ``calc_n2`` at cell centres (:328-411), ``adjust_negative_n2`` (:30-71),
``calc_stability_angle`` (:714-766), ``calc_wave_speed`` (:798-828).  Field arithmetic runs in
libmomlevel_b200; attributes are the reference's CF metadata.
"""

import numpy as np

from . import core, util
from .labeled import DataArray

__all__ = ["adjust_negative_n2", "calc_alpha", "calc_beta", "calc_dz", "calc_masso", "calc_n2", "calc_pdens", "calc_rho",
           "calc_rhoga", "calc_spice", "calc_stability_angle", "calc_volo", "calc_wave_speed"]


def _as_labeled(x):
    return x if isinstance(x, DataArray) else DataArray(x)


def _eos_apply(func_name, thetao, so, pres, eos):
    """``xr.apply_ufunc(eos_func, thetao, so, pres)`` for the operand layouts of the path."""
    util.eos_func_from_str(eos, func_name=func_name)  # ValueError / AssertionError as util.py:243-249
    thetao, so = _as_labeled(thetao), _as_labeled(so)
    t_bcast = so.ndim == thetao.ndim + 1 and so.dims[1:] == thetao.dims
    s_bcast = thetao.ndim == so.ndim + 1 and thetao.dims[1:] == so.dims
    full = so if t_bcast else thetao
    if not (t_bcast or s_bcast) and thetao.dims != so.dims:
        raise ValueError(f"cannot broadcast thetao{thetao.dims} against so{so.dims}")
    z_axis, p = None, pres
    if isinstance(pres, DataArray):
        if pres.ndim == 0:
            p = float(pres)
        elif pres.ndim == 1 and pres.dims[0] in full.dims:
            z_axis, p = full.dims.index(pres.dims[0]), pres.data
        elif pres.dims == full.dims:
            p = pres.data
        else:
            raise ValueError(f"cannot broadcast pres{pres.dims} against {full.dims}")
    if (t_bcast or s_bcast) and z_axis not in (None, 1):
        raise ValueError("with a time-invariant operand the pressure must vary along the level axis")
    out = core.eos_eval(eos, func_name, thetao.data, so.data, p, z_axis=z_axis, t_bcast=t_bcast, s_bcast=s_bcast)
    return DataArray(out, full.dims, coords=full.coords)


def calc_rho(thetao, so, pres, eos="Wright"):
    """In situ density (derived.py:597-639)."""
    rho = _eos_apply("density", thetao, so, pres, eos)
    rho.attrs = {
        "standard_name": "sea_water_density",
        "long_name": "In situ sea water density",
        "comment": f"calculated with the {eos} equation of state",
        "units": "kg m-3",
    }
    return rho


def calc_alpha(thetao, so, pres, eos="Wright"):
    """Thermal expansion coefficient (derived.py:74-115)."""
    alpha = _eos_apply("alpha", thetao, so, pres, eos)
    alpha.attrs = {
        "long_name": "Thermal expansion coefficient",
        "comment": f"calculated with the {eos} equation of state",
        "units": "degC-1",
    }
    return alpha


def calc_beta(thetao, so, pres, eos="Wright"):
    """Haline contraction coefficient (derived.py:118-159)."""
    beta = _eos_apply("beta", thetao, so, pres, eos)
    beta.attrs = {
        "long_name": "Haline contraction coefficient",
        "comment": f"calculated with the {eos} equation of state",
        "units": "PSU-1",
    }
    return beta


def calc_pdens(thetao, so, level=0.0, patm=101325, eos="Wright"):
    """Potential density referenced to ``level`` dbar (derived.py:447-486)."""
    pres = (level * 1.0e4) + patm
    rhopot = _eos_apply("density", thetao, so, float(pres), eos)
    rhopot.attrs = {
        "long_name": f"Potential density referenced to {level} dbar",
        "comment": f"calculated with the {eos} equation of state",
        "units": "kg m-3",
    }
    return rhopot


def calc_spice(thetao, so):
    """Seawater spiciness, Flament 2002 (derived.py:669-711)."""
    thetao, so = _as_labeled(thetao), _as_labeled(so)
    pi = DataArray(core.flament_spice(thetao.data, so.data), thetao.dims, coords=thetao.coords)
    pi.attrs = {
        "long_name": "Sea water spiciness",
        "comment": "calculated based on Flament 2002 methodology",
        "units": "1",
    }
    return pi


def calc_dz(levels, interfaces, depth, top=0.0, bottom=None, fraction=False):
    """dz with partial bottom cells (derived.py:249-325); dims ``(y, x, z)`` as the reference."""
    levels, interfaces, depth = _as_labeled(levels), _as_labeled(interfaces), _as_labeled(depth)
    # derived.py:284-292
    assert bool(np.all(np.nan_to_num(depth.values, nan=0.0) >= 0)), "Depth values must all be positive-definite"
    assert bool(np.all(levels.values >= 0)), "Vertical coordinate levels must all be positive-definite"
    assert bool(np.all(interfaces.values >= 0)), "Vertical coordinate interfaces must all be positive-definite"
    dz = core.calc_dz(interfaces.data, depth.data, top=top, bottom=bottom, fraction=fraction)  # [z][y][x]
    out = DataArray(dz, levels.dims + depth.dims)
    return out.transpose(*(depth.dims + levels.dims))


def calc_volo(volcello):
    """Total ocean volume (derived.py:769-795): skipna sum, two fixed-order stages on the device (``ml_calc_masso``)."""
    volcello = _as_labeled(volcello)
    assert len(volcello.dims) == 3, "Expecting only 3 dimensions for volcello"
    volo = DataArray(core.weighted_nansum(volcello.data)[0], ())
    volo.attrs = {"standard_name": "sea_water_volume", "long_name": "Sea Water Volume", "units": "m3"}
    return volo


def calc_masso(rho, volcello, tcoord="time"):
    """Total ocean mass per time step (derived.py:414-444): skipna sum of ``rho * volcello``, reduced on the device
    without the ``rho * volcello`` temporary (``ml_calc_masso``)."""
    rho, volcello = _as_labeled(rho), _as_labeled(volcello)
    import torch

    if tcoord in rho.dims:
        if rho.dims[0] != tcoord:
            rho = rho.transpose(tcoord, ...)
        r = core.to_device(rho.data)
        v = volcello
        if tcoord in v.dims:  # a time-dependent volume: fall back to the element-wise product, row by row
            v = v.transpose(tcoord, ...) if v.dims[0] != tcoord else v
            vd = core.to_device(v.data)
            rows = [core.weighted_nansum(r[t], vd[t])[0] for t in range(r.shape[0])]
            masso = DataArray(torch.stack(rows), (tcoord,))
        else:
            if tuple(v.dims) != tuple(rho.dims[1:]):
                v = v.transpose(*rho.dims[1:])
            masso = DataArray(core.weighted_nansum(r, v.data, nrows=r.shape[0]), (tcoord,))
    else:
        v = volcello if tuple(volcello.dims) == tuple(rho.dims) else volcello.transpose(*rho.dims)
        masso = DataArray(core.weighted_nansum(rho.data, v.data)[0], ())
    masso.attrs = {"standard_name": "sea_water_mass", "long_name": "Sea Water Mass", "units": "kg"}
    return masso


def calc_rhoga(masso, volo):
    """Global average density (derived.py:642-666)."""
    rhoga = _as_labeled(masso) / _as_labeled(volo)
    rhoga.attrs = {"long_name": "Global Average Sea Water Density", "units": "kg m-3"}
    return rhoga


# ------------------------------------------------------------------ column diagnostics


def _levels(da, zcoord):
    """Values of the level coordinate ``da[zcoord]`` (what ``differentiate(zcoord)`` uses)."""
    if zcoord not in da.dims:
        raise KeyError(zcoord)
    z = da.coords.get(zcoord)
    if z is None:
        raise KeyError(f"{zcoord!r} is a dimension without coordinate values; take the variable out of its Dataset "
                       "or pass coords={...}")
    return z.data if isinstance(z, DataArray) else z


def adjust_negative_n2(n2, zcoord="z_l"):
    """Chelton et al. (1998) adjustment of negative N2 (derived.py:30-71)."""
    n2 = _as_labeled(n2)
    out = DataArray(core.adjust_negative_n2(n2.data, z_axis=n2.dims.index(zcoord)), n2.dims, coords=n2.coords)
    out.attrs = {**n2.attrs, "comment": "adjustment applied for negative values"}
    return out


def calc_n2(thetao, so, eos="Wright", gravity=-9.8, patm=101325.0, zcoord="z_l", interfaces=None, adjust_negative=False):
    """Squared buoyancy frequency at cell centres (derived.py:328-411)."""
    if interfaces is not None:
        raise NotImplementedError("calc_n2 on cell interfaces needs xgcm's vertical transform (derived.py:391-395)")
    util.eos_func_from_str(eos, func_name="alpha")
    thetao, so = _as_labeled(thetao), _as_labeled(so)
    if thetao.dims != so.dims:
        raise ValueError(f"cannot broadcast thetao{thetao.dims} against so{so.dims}")
    if not np.ndim(patm) == 0:
        raise NotImplementedError("a non-scalar patm is not supported")
    out = core.calc_n2(thetao.data, so.data, _levels(thetao, zcoord), eos=eos, gravity=gravity, patm=float(patm),
                       z_axis=thetao.dims.index(zcoord), adjust_negative=adjust_negative)
    n2 = DataArray(out, thetao.dims, coords=thetao.coords)
    n2.attrs = {
        "standard_name": "square_of_brunt_vaisala_frequency_in_sea_water",
        "long_name": "Square of seawater buoyancy frequency",
        "units": "s-2",
    }
    if adjust_negative:
        n2.attrs["comment"] = "adjustment applied for negative values"
    return n2


def calc_stability_angle(thetao, so, pres, eos="Wright", zcoord="z_l"):
    """Stability (Turner) angle in degrees (derived.py:714-766)."""
    util.eos_func_from_str(eos, func_name="alpha")
    thetao, so, pres = _as_labeled(thetao), _as_labeled(so), _as_labeled(pres)
    if thetao.dims != so.dims:
        raise ValueError(f"cannot broadcast thetao{thetao.dims} against so{so.dims}")
    if pres.dims != (zcoord,):
        raise NotImplementedError("calc_stability_angle takes a pressure that varies along the level axis only")
    out = core.stability_angle(thetao.data, so.data, pres.data, _levels(thetao, zcoord), eos=eos,
                               z_axis=thetao.dims.index(zcoord))
    result = DataArray(out, thetao.dims, coords=thetao.coords, name="tu_angle")
    result.attrs = {"long_name": "Stability angle", "units": "degrees"}
    return result


def calc_wave_speed(n2, dz, zcoord="z_l"):
    """Gravity wave speed of the first baroclinic mode (derived.py:798-828).

    As in the reference, the mask ``xr.where(n2[0].isnull(), nan, result)`` pairs index 0 of n2's
    FIRST axis with the column sums, so for a 4-D ``n2`` the result carries the dims
    ``(z, y, x, time)``: the column sums repeated over the levels where the first time step is present.
    """
    import torch

    n2, dz = _as_labeled(n2), _as_labeled(dz)
    zax = n2.dims.index(zcoord)
    hdims = n2.dims[zax + 1:]
    dzt = dz.transpose(zcoord, *hdims)  # the reference's dz is (y, x, z)
    dz_data = core.to_device(dzt.data, torch.float64).contiguous()
    c1 = core.wave_speed(n2.data, dz_data, z_axis=zax)  # dims: n2.dims without zcoord
    sdims = n2.dims[:zax] + hdims
    first = n2[0]
    null = torch.isnan(core.to_device(first.data, torch.float64))
    # xr.where broadcasts by dimension name: the dims of n2[0] first, then the remaining dims of the sums
    rdims = first.dims + tuple(d for d in sdims if d not in first.dims)
    size = dict(zip(n2.dims, n2.shape))
    sums = c1.permute(*[sdims.index(d) for d in rdims if d in sdims])
    sums = sums.reshape([size[d] if d in sdims else 1 for d in rdims])
    null = null.reshape([size[d] if d in first.dims else 1 for d in rdims])
    out = torch.where(null, torch.full((), float("nan"), dtype=torch.float64, device=c1.device), sums)
    result = DataArray(out, rdims)
    result.attrs = {"long name": "Ocean gravity wave speed of the first baroclinic mode", "units": "m s-1"}
    return result
