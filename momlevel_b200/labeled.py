"""Minimal labelled arrays: the slice of the xarray surface that ``momlevel.steric`` touches.

The reference's API is "xarray Dataset in, xarray Dataset out" (src/momlevel/steric.py:17-31).
xarray is an optional dependency here: when it is importable, ``momlevel_b200.steric`` accepts
and returns real ``xarray.Dataset`` objects (see ``xarray_io``); these two classes carry the
same information -- named dims, ``attrs``, ``encoding``, index coordinates -- without it,
and let a variable be backed by a CUDA tensor so fields can stay resident in HBM.

Only metadata-sized arithmetic happens here (a pressure vector, an area sum); every field
computation goes through ``momlevel_b200.core`` to the CUDA library.
"""

import numpy as np
import torch

__all__ = ["ChunkedArray", "DataArray", "Dataset"]


def _is_tensor(x):
    return isinstance(x, torch.Tensor)


def _to_numpy(x):
    if _is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class ChunkedArray:
    """A field that exists as consecutive blocks along its first axis -- what a dask-backed xarray variable is.

    ``blocks`` is a callable that returns a fresh iterator over the blocks (numpy arrays ``[n_i, ...]`` whose lengths
    are ``chunks``); nothing is read until somebody iterates.  ``steric()`` streams such fields through the device
    block by block (``core.HostStream``) instead of asking for the whole array, which for a daily global series
    (BASELINE config 4: 1.4 TB) does not exist anywhere.  Indexing a step range reads only the blocks it touches;
    ``numpy.asarray`` concatenates everything (the fallback for code that is not block-aware).
    """

    def __init__(self, shape, dtype, chunks, blocks):
        self.shape = tuple(int(n) for n in shape)
        self.dtype = np.dtype(dtype)
        self.chunks = tuple(int(c) for c in chunks)
        assert sum(self.chunks) == self.shape[0], "the blocks must cover the first axis"
        self._blocks = blocks
        self.blocks_read = 0  # accounting for tests: how many blocks have been produced

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def flags(self):
        return {"C_CONTIGUOUS": True}

    def blocks(self):
        for blk in self._blocks():
            self.blocks_read += 1
            yield np.asarray(blk)

    def _steps(self, start, stop):
        """Steps ``[start, stop)`` as one numpy array, reading only the blocks that hold them."""
        out, t0 = [], 0
        for n, blk in zip(self.chunks, self.blocks()):
            lo, hi = max(start, t0), min(stop, t0 + n)
            if lo < hi:
                out.append(np.asarray(blk)[lo - t0: hi - t0])
            t0 += n
            if t0 >= stop:
                break
        return np.concatenate(out) if len(out) != 1 else out[0]

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        k0, rest = key[0], key[1:]
        if isinstance(k0, (int, np.integer)):
            i = int(k0) % self.shape[0]
            first = self._steps(i, i + 1)[0]
            return first[rest] if rest else first
        start, stop, step = k0.indices(self.shape[0])
        assert step == 1, "a chunked field is sliced in whole step ranges"
        sub = self._steps(start, stop)
        return sub[(slice(None),) + rest] if rest else sub

    def __array__(self, dtype=None, copy=None):
        full = np.concatenate(list(self.blocks())) if self.chunks else np.empty(self.shape, self.dtype)
        return full.astype(dtype) if dtype is not None else full

    def __repr__(self):
        return f"<momlevel_b200.ChunkedArray {self.shape} {self.dtype} in {len(self.chunks)} blocks>"


class DataArray:
    """N-d array with dimension names. ``data`` is a numpy array, a torch tensor or a :class:`ChunkedArray`."""

    __array_priority__ = 50

    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        if isinstance(data, DataArray):
            dims = data.dims if dims is None else dims
            attrs = dict(data.attrs) if attrs is None else attrs
            coords = dict(data.coords) if coords is None else coords
            data = data._data
        if not _is_tensor(data) and not callable(data) and not isinstance(data, ChunkedArray):
            data = np.asarray(data)
        self._data = data
        self._lazy_shape = None
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(self._data.ndim))
        if isinstance(dims, str):
            dims = (dims,)
        if isinstance(dims, dict):  # the reference passes {"time": coord, ...} (test_data/__init__.py:67)
            coords = dict(dims) if coords is None else coords
            dims = tuple(dims.keys())
        self.dims = tuple(dims)
        self.coords = {} if coords is None else {k: v for k, v in dict(coords).items()}
        self.attrs = {} if attrs is None else dict(attrs)
        self.encoding = {}
        self.name = name

    # ------------------------------------------------------------------ lazy variables
    @classmethod
    def lazy(cls, compute, shape, dims, attrs=None):
        """A variable whose data is produced by ``compute()`` on first access."""
        out = cls.__new__(cls)
        out._data = compute
        out._lazy_shape = tuple(shape)
        out.dims = tuple(dims)
        out.coords = {}
        out.attrs = {} if attrs is None else dict(attrs)
        out.encoding = {}
        out.name = None
        return out

    @property
    def is_lazy(self):
        return callable(self._data)

    @property
    def data(self):
        if callable(self._data):
            self._data = self._data()
        return self._data

    @property
    def values(self):
        return _to_numpy(self.data)

    def __array__(self, dtype=None, copy=None):
        v = self.values
        return v.astype(dtype) if dtype is not None else v

    @property
    def shape(self):
        return self._lazy_shape if callable(self._data) else tuple(self._data.shape)

    @property
    def ndim(self):
        return len(self.dims)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def sizes(self):
        return dict(zip(self.dims, self.shape))

    def __len__(self):
        return self.shape[0]

    def __float__(self):
        return float(self.values.reshape(()))

    def __bool__(self):
        return bool(self.values)

    def __repr__(self):
        where = "cuda" if _is_tensor(self._data) and self._data.is_cuda else ("lazy" if self.is_lazy else "host")
        return f"<momlevel_b200.DataArray {dict(zip(self.dims, self.shape))} [{where}]>"

    def _new(self, data, dims=None, keep_attrs=False):
        dims = self.dims if dims is None else dims
        coords = {k: v for k, v in self.coords.items() if k in dims}
        return DataArray(data, dims, coords, self.attrs if keep_attrs else None)

    # ---------------------------------------------------------------------- structure
    def copy(self, deep=True):
        d = self.data
        d = (d.clone() if _is_tensor(d) else d.copy()) if deep else d
        out = self._new(d, keep_attrs=True)
        out.encoding = dict(self.encoding)
        return out

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        idx = tuple(indexers.get(d, slice(None)) for d in self.dims)
        dims = tuple(d for d in self.dims if not isinstance(indexers.get(d, slice(None)), (int, np.integer)))
        return self._new(self.data[idx], dims, keep_attrs=True)

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        key = key + (slice(None),) * (self.ndim - len(key))
        dims = tuple(d for d, k in zip(self.dims, key) if not isinstance(k, (int, np.integer)))
        return self._new(self.data[key], dims, keep_attrs=True)

    def __setitem__(self, key, value):
        self.data[key] = value

    def squeeze(self):
        keep = [i for i, n in enumerate(self.shape) if n != 1]
        d = self.data
        d = d.reshape([self.shape[i] for i in keep])
        return self._new(d, tuple(self.dims[i] for i in keep), keep_attrs=True)

    def reset_coords(self, drop=True):
        out = self._new(self.data, keep_attrs=True)
        out.coords = {k: v for k, v in self.coords.items() if k in self.dims}
        return out

    def transpose(self, *dims):
        if Ellipsis in dims:
            i = dims.index(Ellipsis)
            rest = tuple(d for d in self.dims if d not in dims)
            dims = dims[:i] + rest + dims[i + 1:]
        if not dims:
            dims = self.dims[::-1]
        perm = [self.dims.index(d) for d in dims]
        d = self.data
        d = d.permute(*perm) if _is_tensor(d) else np.transpose(d, perm)
        return self._new(d, tuple(dims), keep_attrs=True)

    def astype(self, dtype):
        d = self.data
        return self._new(d.to(dtype) if _is_tensor(d) else d.astype(dtype), keep_attrs=True)

    # --------------------------------------------------------------------- reductions
    def sum(self, dim=None, skipna=True):
        """xarray's ``sum``: NaNs are skipped for float data (an all-NaN slice gives 0)."""
        d = self.data
        if dim is None:
            axes, dims = None, ()
        else:
            names = (dim,) if isinstance(dim, str) else tuple(dim)
            axes = tuple(self.dims.index(n) for n in names)
            dims = tuple(n for n in self.dims if n not in names)
        if _is_tensor(d):
            r = torch.nansum(d) if axes is None else torch.nansum(d, dim=axes)
        else:
            r = np.nansum(d, axis=axes) if skipna else np.sum(d, axis=axes)
        return self._new(r, dims)

    def notnull(self):
        d = self.data
        return self._new(~torch.isnan(d) if _is_tensor(d) else ~np.isnan(d))

    def fillna(self, value):
        d = self.data
        r = torch.nan_to_num(d, nan=value) if _is_tensor(d) else np.where(np.isnan(d), value, d)
        return self._new(r, keep_attrs=True)

    def where(self, cond, other=np.nan):
        c = cond.data if isinstance(cond, DataArray) else cond
        d = self.data
        if _is_tensor(d):
            c = torch.as_tensor(c, device=d.device)
            return self._new(torch.where(c, d, torch.full_like(d, other)), keep_attrs=True)
        return self._new(np.where(_to_numpy(c), d, other), keep_attrs=True)

    # --------------------------------------------------------------------- arithmetic
    def _binary(self, other, op, reflect=False):
        a = self.data
        if isinstance(other, DataArray):
            if other.dims != self.dims and other.ndim != 0 and self.ndim != 0:
                raise ValueError(f"labeled arithmetic needs equal dims, got {self.dims} and {other.dims}")
            b = other.data
            if _is_tensor(a) != _is_tensor(b):
                a, b = _to_numpy(a), _to_numpy(b)
            dims = self.dims if self.ndim else other.dims
        else:
            b, dims = other, self.dims
        r = op(b, a) if reflect else op(a, b)
        return self._new(r, dims)

    def __add__(self, o):
        return self._binary(o, lambda a, b: a + b)

    def __radd__(self, o):
        return self._binary(o, lambda a, b: a + b, True)

    def __sub__(self, o):
        return self._binary(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._binary(o, lambda a, b: a - b, True)

    def __mul__(self, o):
        return self._binary(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return self._binary(o, lambda a, b: a * b, True)

    def __truediv__(self, o):
        return self._binary(o, lambda a, b: a / b)

    def __rtruediv__(self, o):
        return self._binary(o, lambda a, b: a / b, True)

    def __neg__(self):
        return self._new(-self.data)


class Dataset:
    """Ordered mapping name -> DataArray with a shared dimension namespace."""

    def __init__(self, data_vars=None, attrs=None):
        object.__setattr__(self, "_vars", {})
        object.__setattr__(self, "attrs", {} if attrs is None else dict(attrs))
        for k, v in (data_vars or {}).items():
            self[k] = v

    # ------------------------------------------------------------------------ mapping
    def __setitem__(self, name, value):
        if isinstance(value, tuple):  # (dims, data[, attrs])
            value = DataArray(value[1], value[0], attrs=value[2] if len(value) > 2 else None)
        elif not isinstance(value, DataArray):
            value = DataArray(value, ())
        else:
            enc = value.encoding
            lazy = value.is_lazy
            if not lazy:
                value = DataArray(value)
                value.encoding = dict(enc)
        value.name = name
        self._vars[name] = value

    def __getitem__(self, name):
        var = self._vars[name]
        # like xarray, a variable taken out of a Dataset carries the index coordinates of its dims
        # (derived.calc_n2 reads thetao[zcoord], derived.py:396)
        if not var.is_lazy:
            for d in var.dims:
                idx = self._vars.get(d)
                if d != name and d not in var.coords and idx is not None and idx.dims == (d,):
                    var.coords[d] = idx
        return var

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_vars")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __contains__(self, name):
        return name in self._vars

    def __iter__(self):
        return iter(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def __repr__(self):
        return "<momlevel_b200.Dataset " + ", ".join(f"{k}{list(v.dims)}" for k, v in self._vars.items()) + ">"

    @property
    def variables(self):
        return dict(self._vars)

    @property
    def dims(self):
        out = {}
        for v in self._vars.values():
            for d, n in zip(v.dims, v.shape):
                out.setdefault(d, n)
        return out

    @property
    def coords(self):
        dims = self.dims
        return {k: v for k, v in self._vars.items() if v.dims == (k,) and k in dims}

    @property
    def data_vars(self):
        c = self.coords
        return {k: v for k, v in self._vars.items() if k not in c}

    # --------------------------------------------------------------------- operations
    def rename(self, name_dict=None):
        """``Dataset.rename`` (steric.py:84); ``None`` is the identity."""
        if not name_dict:
            return self  # nothing to rename: the callers only read
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            nv = DataArray(v, tuple(name_dict.get(d, d) for d in v.dims)) if not v.is_lazy else v
            nv.encoding = dict(v.encoding)
            out[name_dict.get(k, k)] = nv
        return out

    def drop_vars(self, names):
        names = [names] if isinstance(names, str) else list(names)
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            if k not in names:
                out[k] = v
        return out

    def copy(self, deep=False):
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            out[k] = v.copy(deep=deep) if not v.is_lazy else v
        return out

    def sum(self):
        """Every data variable summed over all its dims (coordinates are dropped)."""
        out = Dataset()
        for k, v in self.data_vars.items():
            if np.issubdtype(np.asarray(v.values).dtype, np.number):
                out[k] = v.sum()
        return out

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            out[k] = v.isel({d: i for d, i in indexers.items() if d in v.dims})
        return out
