"""Host-side helpers of the steric path: coordinate names, EOS dispatch, dataset validation.

Same names, arguments and error behaviour as the reference functions they mirror
(``src/momlevel/util.py``): ``default_coords`` (:199-224), ``eos_func_from_str``
(:227-249), ``validate_areacello`` (:669-694), ``validate_dataset`` (:697-814).
Validation works on metadata (names, ranks, one area sum) and stays on the host so that
messages, warnings and exception types are drop-in.
"""

import warnings

import numpy as np

__all__ = ["default_coords", "eos_func_from_str", "validate_areacello", "validate_dataset", "annual_average",
           "calendar_axis", "whole_years_in_order"]


def default_coords(coord_names=None):
    """util.py:199-224 -> ``(tcoord, zcoord, zbounds)``."""
    coord_names = {} if coord_names is None else coord_names
    assert isinstance(coord_names, dict), "Coordinate mapping must be a dictionary."
    zcoord = coord_names["z"] if "z" in coord_names.keys() else "z_l"
    zbounds = coord_names["zbounds"] if "zbounds" in coord_names.keys() else "z_i"
    tcoord = coord_names["t"] if "t" in coord_names.keys() else "time"
    return (tcoord, zcoord, zbounds)


def eos_func_from_str(eos_str, func_name="density"):
    """util.py:227-249: resolve ``momlevel_b200.eos.<eos_str>.<func_name>``."""
    from . import eos

    assert isinstance(eos_str, str), "Expecting string for equation of state"
    eos_str = eos_str.lower()
    avail_eos = [k for k, v in eos.__dict__.items() if not k.startswith("_")]
    if eos_str not in avail_eos:
        raise ValueError(f"Unknown equation of state: {eos_str}")
    return eos.__dict__[eos_str].__dict__[func_name]


def _total(arr):
    """``DataArray.sum()`` as a python float, for xarray, labelled or plain arrays."""
    s = arr.sum()
    return float(s.values) if hasattr(s, "values") else float(s)


def validate_areacello(areacello, reference=3.6111092e14, tolerance=0.02, total=None):
    """util.py:669-694: ocean area within ``tolerance`` of the real-world total.

    ``total`` hands in ``areacello.sum()`` when the caller has already read it back from the device.
    """
    error = ((_total(areacello) if total is None else float(total)) - reference) / reference
    return bool(np.abs(error) < tolerance)


def validate_dataset(dset, reference=False, strict=True, additional_vars=None, area_total=None):
    """util.py:697-814: presence and rank of the required variables.

    All problems are collected, printed, and reported as one ``ValueError``; a bad
    ``areacello`` is only a warning when ``strict`` is False.  ``area_total`` (not in the reference) lets a
    caller that reads device values back in one go supply ``areacello.sum()``; ``False`` leaves the area
    check to a later call.
    """
    dset_varlist = list(dset.variables)
    exceptions = []

    # util.py:726-734 -- the reference compares bound methods (`x.lower` without the call),
    # so this check can never fire there; kept inert for drop-in behaviour.

    expected_varlist = ["thetao", "so", "volcello", "areacello"]
    if additional_vars is not None:
        additional_vars = [additional_vars] if not isinstance(additional_vars, list) else additional_vars
    else:
        additional_vars = []
    expected_varlist = expected_varlist + additional_vars
    reference_varlist = ["rho", "volo", "masso", "rhoga"]
    expected_varlist = expected_varlist + reference_varlist if reference else expected_varlist

    missing = list(set(expected_varlist) - set(dset_varlist))

    def _collect(cond, message):
        if not cond:
            exceptions.append(AssertionError(message))

    _collect(len(missing) == 0, f"Reference dataset is missing variables: {missing}")

    ranks = (3, "(z,y,x)") if reference else (4, ("t,z,y,x"))
    for var in ["thetao", "so", "volcello"]:
        if var in dset_varlist:
            _collect(len(dset[var].dims) == ranks[0], f"Variable {var} must have exactly {ranks[0]} dimensions {ranks[1]}")

    for var in ["areacello", "deptho"]:
        if var in dset_varlist:
            _collect(len(dset[var].dims) == 2, f"Variable {var} must have exactly 2 dimensions (y,x)")

    if "areacello" in dset_varlist and area_total is not False:
        if not validate_areacello(dset["areacello"], total=area_total):
            message = "Variable `areacello` field is out of range. It may not be masked."
            if not strict:
                warnings.warn(message)
            else:
                exceptions.append(AssertionError(message))

    if reference:
        if "rho" not in missing:
            _collect(len(dset["rho"].dims) == 3, "Variable areacello must have exactly 3 dimensions (z,y,x)")
        for var in ["masso", "volo", "rhoga"]:
            if var not in missing:
                _collect(len(dset[var].dims) == 0, f"Variable {var} must be a scalar")

    if len(exceptions) > 0:
        for e in exceptions:
            print(e)
        raise ValueError("Errors found in dataset.")


def calendar_axis(tvals):
    """``(years, days_in_month)`` per step of a calendar time axis, or ``None`` when the axis carries no calendar.

    What ``util.py:79-87`` reads through ``time.dt.year`` / ``time.dt.days_in_month``: cftime objects (and
    ``cftime_lite.Datetime``) have ``.year`` and ``.daysinmonth``; numpy ``datetime64`` is handled arithmetically.
    """
    vals = np.asarray(tvals)
    if vals.ndim != 1 or vals.size == 0:
        return None
    if vals.dtype.kind == "M":
        months = vals.astype("datetime64[M]")
        dim = ((months + 1).astype("datetime64[D]") - months.astype("datetime64[D]")).astype(np.int64)
        years = vals.astype("datetime64[Y]").astype(np.int64) + 1970
        return years, dim.astype(np.float64)
    if vals.dtype == object and all(hasattr(v, "year") and hasattr(v, "daysinmonth") for v in vals):
        return (np.array([int(v.year) for v in vals], dtype=np.int64),
                np.array([float(v.daysinmonth) for v in vals], dtype=np.float64))
    return None


def _mid_year(sample, year):
    """Mid-point of ``year`` in the calendar of ``sample`` (util.py:96-102: ``bounds[0] + (bounds[1] - bounds[0]) / 2``)."""
    if isinstance(sample, np.datetime64):
        b0, b1 = np.datetime64(f"{int(year):04d}-01-01", "s"), np.datetime64(f"{int(year) + 1:04d}-01-01", "s")
        return b0 + (b1 - b0) // 2
    b0 = sample.replace(year=int(year), month=1, day=1, hour=0, minute=0, second=0, microsecond=0)
    b1 = b0.replace(year=int(year) + 1)
    return b0 + (b1 - b0) / 2


def whole_years_in_order(years):
    """True when the steps come as consecutive blocks of twelve, one calendar year each, in ascending order."""
    years = np.asarray(years)
    if years.size % 12:
        return False
    blocks = years.reshape(-1, 12)
    return bool(np.all(blocks == blocks[:, :1]) and np.all(np.diff(blocks[:, 0]) > 0))


def annual_average(xobj, tcoord="time", days_in_month=None):
    """Days-in-month weighted annual means (util.py:49-119), for labelled Datasets.

    As in the reference the steps are grouped by calendar year (``groupby("time.year")``, every group must hold
    twelve steps), weighted by the days in their month (``time.dt.days_in_month``) with NaNs skipped and the
    weights renormalised per cell, and each mean is labelled with the mid-point of its year.  Year and weights are
    read from the time axis when it holds calendar objects (cftime, ``cftime_lite.Datetime``, ``datetime64``).
    ``days_in_month`` (length nt, a multiple of 12; not in the reference) supplies the weights for a time axis
    without a calendar: the steps are then taken as consecutive years of twelve.
    """
    from .labeled import DataArray, Dataset

    if isinstance(xobj, DataArray):
        tvals = np.asarray(xobj.coords[tcoord].values) if tcoord in xobj.coords else None
    else:
        tvals = np.asarray(xobj[tcoord].values) if tcoord in xobj.variables else None
    cal = calendar_axis(tvals) if tvals is not None else None
    order = None
    year_labels = None
    if days_in_month is None:
        if cal is None:
            raise ValueError("annual_average needs `days_in_month` when the time axis carries no calendar")
        years, days_in_month = cal
    elif cal is not None:
        years = cal[0]
    else:
        years = None
    w = np.asarray(days_in_month, dtype=np.float64)
    if years is not None:
        assert years.size == w.size, "one weight per time step"
        uniq, counts = np.unique(years, return_counts=True)
        assert np.all(counts == 12), "annual averaging needs twelve steps in every year"  # util.py:82
        if not whole_years_in_order(years):
            order = np.argsort(years, kind="stable")  # groupby gathers a year's steps wherever they are
            w = w[order]
        year_labels = uniq
    assert w.size % 12 == 0, "annual averaging needs whole years of monthly data"
    nyears = w.size // 12
    w = w.reshape(nyears, 12)

    def _one(da):
        if tcoord not in da.dims:
            return da
        import torch

        d = da.data
        if not isinstance(d, torch.Tensor) and np.asarray(d).dtype.kind not in "fiu":
            return None  # util.py:72-73: non-numeric variables are skipped
        ax = da.dims.index(tcoord)
        lib = torch if isinstance(d, torch.Tensor) else np
        d = lib.movedim(d, ax, 0) if ax else d
        if order is not None:
            d = d[torch.as_tensor(order, device=d.device)] if lib is torch else d[order]
        d = d.reshape((nyears, 12) + tuple(d.shape[1:]))
        ww = w.reshape((nyears, 12) + (1,) * (d.ndim - 2))
        if lib is torch:
            ww = torch.as_tensor(ww, device=d.device)
        # xarray's weighted mean: NaNs are skipped and the weights renormalised per cell
        valid = ~lib.isnan(d)
        num = lib.where(valid, d, lib.zeros_like(d) if lib is torch else 0.0) * ww
        den = valid * ww
        out = num.sum(1) / den.sum(1)
        dims = (tcoord,) + tuple(x for x in da.dims if x != tcoord)
        res = DataArray(out, dims, attrs=da.attrs)
        res.encoding = dict(da.encoding)
        return res

    if isinstance(xobj, DataArray):
        return _one(xobj)
    out = Dataset(attrs=xobj.attrs)
    for k, v in xobj.data_vars.items():
        r = _one(v)
        if r is not None:
            out[k] = r
    for k, v in xobj.coords.items():
        if k != tcoord:
            out[k] = v
    # each mean is labelled with the mid-point of its year (util.py:96-107); a time axis without a calendar keeps
    # the mean of the twelve original time values
    if tvals is not None:
        if year_labels is not None:
            mids = np.empty(nyears, dtype=object if tvals.dtype == object else tvals.dtype)
            for i, y in enumerate(year_labels):
                mids[i] = _mid_year(tvals[0], y)
            out[tcoord] = DataArray(mids, (tcoord,), attrs=xobj[tcoord].attrs)
        elif np.issubdtype(tvals.dtype, np.number):
            tv = np.asarray(tvals, dtype=np.float64).reshape(nyears, 12).mean(1)
            out[tcoord] = DataArray(tv, (tcoord,), attrs=xobj[tcoord].attrs)
    return out
