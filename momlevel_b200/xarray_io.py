"""xarray <-> labelled conversion, used only when the caller hands in real xarray objects.

xarray is optional (it is absent from the build image); nothing here is imported unless a
``xarray.Dataset`` reaches ``momlevel_b200.steric``.  In this image the adapter is exercised against
``tests/_stubs/xarray`` (a duck-typed stand-in), i.e. it is stub-verified, not verified against xarray itself.

Laziness is kept across the boundary in both directions:

* a dask-backed 4-D variable (``xr.open_mfdataset(..., chunks={"time": 1, ...})``, example.ipynb cell 4) becomes a
  ``labeled.ChunkedArray`` whose blocks are computed one at a time when ``steric()`` streams them
  (``core.HostStream``); it is never asked for as a whole array;
* only the variables the path reads are converted (``names``); the others are left alone;
* a lazy result variable (``delta_rho``, ``reference["rho"]``: each as large as the inputs) goes back as a
  dask-delayed array when dask is importable, and is evaluated at the boundary only when it is not.
"""

import warnings

import numpy as np

from .labeled import ChunkedArray, DataArray, Dataset

__all__ = ["from_xarray", "to_xarray"]


def _chunks_along_first_axis(var):
    """Block lengths along axis 0 of a dask-backed variable, or ``None`` for one that is already in memory."""
    data = getattr(var, "data", None)
    chunks = getattr(data, "chunks", None)
    if chunks is None or getattr(var, "ndim", 0) != 4:
        return None
    first = chunks[0]
    return tuple(int(n) for n in (first if isinstance(first, (tuple, list)) else (first,)))


def _block_source(var, lens):
    def blocks():
        t = 0
        for n in lens:
            yield np.asarray(var[t: t + n].values)  # computes this block only
            t += n

    return blocks


def from_xarray(ds, names=None):
    """``xarray.Dataset`` -> labelled ``Dataset``.  ``names``: the variables to convert (plus the coordinates of
    their dimensions); ``None`` converts every numeric variable."""
    out = Dataset(attrs=dict(ds.attrs))
    variables = ds.variables
    if names is not None:
        wanted = [n for n in names if n in variables]
        dims = []
        for n in wanted:
            dims += [d for d in variables[n].dims if d in variables and d not in wanted and d not in dims]
        wanted += dims
    else:
        wanted = list(variables)
    for name in wanted:
        var = variables[name]
        lens = _chunks_along_first_axis(var)
        if lens is not None:
            arr = ChunkedArray(var.shape, var.dtype, lens, _block_source(var, lens))
            out[name] = DataArray(arr, var.dims, attrs=dict(var.attrs))
            out[name].encoding = dict(var.encoding)
            continue
        vals = var.values
        if vals.dtype.kind not in "fiu":
            if var.dims == (name,):  # calendar axes (cftime objects, datetime64) stay as they are: util.calendar_axis
                out[name] = DataArray(vals, var.dims, attrs=dict(var.attrs))
            continue
        out[name] = DataArray(vals, var.dims, attrs=dict(var.attrs))
        out[name].encoding = dict(var.encoding)
    return out


def _delayed(var):
    """A dask array that evaluates the lazy variable on first use, or ``None`` when dask is not importable."""
    try:
        import dask
        import dask.array as da
    except ImportError:
        return None
    return da.from_delayed(dask.delayed(lambda: np.asarray(var.values))(), shape=tuple(var.shape), dtype=np.float64)


def to_xarray(ds, like=None):
    import xarray as xr

    out = xr.Dataset(attrs=dict(ds.attrs))
    for name, var in ds.variables.items():
        data = None
        if var.is_lazy:
            data = _delayed(var)
            if data is None and int(np.prod(var.shape, dtype=np.int64)) * 8 > (1 << 32):
                warnings.warn(f"'{name}' ({int(np.prod(var.shape, dtype=np.int64)) * 8 >> 30} GiB) is evaluated at the xarray "
                              "boundary because dask is not installed; pass labelled Datasets to keep it lazy")
        if data is None:
            data = np.asarray(var.values)
        out[name] = xr.DataArray(data, dims=var.dims, attrs=dict(var.attrs))
        out[name].encoding.update(var.encoding)
    if like is not None:
        for c in like.coords:
            if c in out.dims and c not in out.variables and like[c].dims == (c,):
                out = out.assign_coords({c: like[c]})
    return out
