"""A duck-typed stand-in for the sliver of xarray that momlevel_b200.xarray_io touches (TEST ONLY).

xarray is not installable in this image; this stub lets the adapter's control flow -- Dataset in,
Dataset out, coordinates carried over -- run in the tests.  It is put on sys.path by the test
itself and only when the real package is absent.
"""

import numpy as np

__version__ = "0.0-stub"


class ChunkedNumpy:
    """Stands in for a dask array: knows its ``chunks``, computes only what is sliced out of it, and keeps
    count of what it was asked to materialise (``largest_compute``: the biggest single request, in elements)."""

    largest_compute = 0
    total_computed = 0

    def __init__(self, array, chunks0):
        self._a = array
        self.chunks = (tuple(chunks0),) + tuple((n,) for n in array.shape[1:])
        assert sum(chunks0) == array.shape[0]

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)
    ndim = property(lambda self: self._a.ndim)

    def __getitem__(self, key):
        sub = self._a[key]
        return ChunkedNumpy(sub, (sub.shape[0],)) if sub.ndim == self._a.ndim else sub

    def compute(self):
        ChunkedNumpy.largest_compute = max(ChunkedNumpy.largest_compute, self._a.size)
        ChunkedNumpy.total_computed += self._a.size
        return np.array(self._a)

    def __array__(self, dtype=None, copy=None):
        out = self.compute()
        return out.astype(dtype) if dtype is not None else out


class Variable:
    def __init__(self, values, dims, attrs=None):
        self._data = values if isinstance(values, ChunkedNumpy) else np.asarray(values)
        self.dims = tuple(dims)
        self.attrs = dict(attrs or {})
        self.encoding = {}

    @property
    def data(self):
        return self._data

    @property
    def values(self):
        return np.asarray(self._data)

    @property
    def shape(self):
        return self._data.shape

    @property
    def ndim(self):
        return self._data.ndim

    @property
    def dtype(self):
        return self._data.dtype

    def __getitem__(self, key):
        sub = self._data[key]
        keys = key if isinstance(key, tuple) else (key,)
        dims = tuple(d for d, k in zip(self.dims, keys + (slice(None),) * (len(self.dims) - len(keys)))
                     if not isinstance(k, (int, np.integer)))
        return Variable(sub, dims, self.attrs)


class DataArray(Variable):
    def __init__(self, data, dims=None, attrs=None, coords=None, name=None):
        data = data if isinstance(data, ChunkedNumpy) else np.asarray(data)
        super().__init__(data, dims if dims is not None else tuple(f"dim_{i}" for i in range(data.ndim)), attrs)
        self.name = name

    def sum(self):
        return DataArray(np.nansum(self.values), ())

    def __float__(self):
        return float(self.values)


class Dataset:
    def __init__(self, data_vars=None, attrs=None):
        self._vars = {}
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, name, value):
        if isinstance(value, tuple):
            value = DataArray(value[1], dims=value[0], attrs=value[2] if len(value) > 2 else None)
        value.name = name
        self._vars[name] = value

    def __getitem__(self, name):
        return self._vars[name]

    def __contains__(self, name):
        return name in self._vars

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
        return {k: v for k, v in self._vars.items() if v.dims == (k,)}

    def assign_coords(self, mapping):
        for k, v in mapping.items():
            self[k] = v
        return self

    def sum(self):
        return Dataset({k: v.sum() for k, v in self._vars.items() if v.values.dtype.kind in "fiu"})
