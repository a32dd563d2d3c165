"""xarray <-> labelled conversion, used only when the caller hands in real xarray objects.

xarray is optional (it is absent from the build image); nothing here is imported unless a
``xarray.Dataset`` reaches ``momlevel_b200.steric``.
"""

import numpy as np

from .labeled import DataArray, Dataset

__all__ = ["from_xarray", "to_xarray"]


def from_xarray(ds):
    out = Dataset(attrs=dict(ds.attrs))
    for name, var in ds.variables.items():
        vals = var.values
        if vals.dtype.kind not in "fiu":
            if var.dims == (name,):  # keep calendar axes as opaque object coordinates
                out[name] = DataArray(vals, var.dims, attrs=dict(var.attrs))
            continue
        out[name] = DataArray(vals, var.dims, attrs=dict(var.attrs))
        out[name].encoding = dict(var.encoding)
    return out


def to_xarray(ds, like=None):
    import xarray as xr

    out = xr.Dataset(attrs=dict(ds.attrs))
    for name, var in ds.variables.items():
        out[name] = xr.DataArray(np.asarray(var.values), dims=var.dims, attrs=dict(var.attrs))
        out[name].encoding.update(var.encoding)
    if like is not None:
        for c in like.coords:
            if c in out.dims and c not in out.variables and like[c].dims == (c,):
                out = out.assign_coords({c: like[c]})
    return out
