"""reference.py -- sea level reference state (mirrors ``src/momlevel/reference.py:15-85``).

One fused kernel (``ml_reference_state``) evaluates the reference density and both global
sums in a single pass over the ``time_index`` slab.
"""

import numpy as np
import torch

from . import core, util
from .labeled import DataArray, Dataset

__all__ = ["setup_reference_state"]


def _pressure(dset, zcoord, patm):
    """steric.py:96 / reference.py:54: 1 m ~ 1 dbar = 1e4 Pa, plus the surface pressure.

    ``patm`` is a scalar (a per-level pressure vector comes back) or a 2-D field over the horizontal dims of the
    dataset -- sea-level pressure from a reanalysis, say -- which the reference broadcasts by dimension name into
    ``pres[z, y, x]``; a :class:`core.Pressure` (per-level part + per-column part) comes back then.
    """
    z = dset[zcoord].data
    on_device = isinstance(z, torch.Tensor) and z.is_cuda
    if not isinstance(patm, (int, float, np.floating, np.integer)):
        nd = getattr(patm, "ndim", np.ndim(patm))
        if nd == 0:
            patm = float(patm)
        elif nd == 2:
            hdims = tuple(d for d in dset["thetao"].dims if d not in (zcoord,))[-2:]
            pdims = getattr(patm, "dims", None)
            if isinstance(patm, (DataArray, torch.Tensor)):
                col = patm.data
            else:  # numpy, or an xarray.DataArray (its .values)
                col = np.asarray(patm.values if hasattr(patm, "values") else patm)
            if pdims is not None and tuple(pdims) != hdims:
                if set(pdims) != set(hdims):
                    raise ValueError(f"`patm` has dims {tuple(pdims)}, expecting the horizontal dims {hdims}")
                col = col.T  # name-based broadcasting: the order of the dims does not matter to the reference
            hshape = tuple(dset["thetao"].shape[-2:])
            if tuple(np.shape(col)) != hshape:
                raise ValueError(f"`patm` has shape {tuple(np.shape(col))}, expecting {hshape}")
            level = (z.to(torch.float64) * 1.0e4) if on_device else np.asarray(dset[zcoord].values, dtype=np.float64) * 1.0e4
            return core.Pressure(level, col)
        else:
            raise NotImplementedError("momlevel_b200 takes `patm` as a scalar or as a 2-D (y, x) field")
    if on_device:  # stays on the device: no read-back in front of the launch
        return (z.to(torch.float64) * 1.0e4) + float(patm)
    return (np.asarray(dset[zcoord].values, dtype=np.float64) * 1.0e4) + float(patm)


def setup_reference_state(dset, patm=101325.0, eos="Wright", coord_names=None, time_index=0):
    """Generate the reference dataset (reference.py:15-85).

    Returns a Dataset with ``thetao``, ``so``, ``volcello``, ``rho`` (3-D), ``volo``,
    ``masso``, ``rhoga`` (scalars) and ``areacello``.
    """
    tcoord, zcoord, _ = util.default_coords(coord_names)
    util.eos_func_from_str(eos)
    pres = _pressure(dset, zcoord, patm)

    reference = Dataset()
    for name in ("thetao", "so", "volcello"):
        reference[name] = dset[name].isel({tcoord: time_index}).squeeze().reset_coords(drop=True)

    rho, sums = core.reference_state(reference["thetao"].data, reference["so"].data, reference["volcello"].data,
                                     pres, eos=eos)
    volo, masso = (float(x) for x in sums.cpu())
    reference["rho"] = DataArray(rho, reference["thetao"].dims, attrs={
        "standard_name": "sea_water_density",
        "long_name": "In situ sea water density",
        "comment": f"calculated with the {eos} equation of state",
        "units": "kg m-3",
    })
    reference["volo"] = DataArray(np.float64(volo), (), attrs={
        "standard_name": "sea_water_volume", "long_name": "Sea Water Volume", "units": "m3"})
    reference["masso"] = DataArray(np.float64(masso), (), attrs={
        "standard_name": "sea_water_mass", "long_name": "Sea Water Mass", "units": "kg"})
    reference["rhoga"] = DataArray(np.float64(masso) / np.float64(volo), (), attrs={
        "long_name": "Global Average Sea Water Density", "units": "kg m-3"})
    reference["areacello"] = dset["areacello"]
    return reference
