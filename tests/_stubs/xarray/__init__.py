"""A duck-typed stand-in for the sliver of xarray that momlevel_b200.xarray_io touches (TEST ONLY).

xarray is not installable in this image; this stub lets the adapter's control flow -- Dataset in,
Dataset out, coordinates carried over -- run in the tests.  It is put on sys.path by the test
itself and only when the real package is absent.
"""

import numpy as np

__version__ = "0.0-stub"


class Variable:
    def __init__(self, values, dims, attrs=None):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        self.attrs = dict(attrs or {})
        self.encoding = {}

    @property
    def shape(self):
        return self.values.shape


class DataArray(Variable):
    def __init__(self, data, dims=None, attrs=None, coords=None, name=None):
        data = np.asarray(data)
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
