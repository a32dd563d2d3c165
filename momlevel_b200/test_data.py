"""test_data.py -- the reference's 5x5x5 synthetic MOM6 dataset as a labelled Dataset.

Mirrors ``momlevel.test_data.generate_test_data`` / ``generate_test_data_dz``
(src/momlevel/test_data/__init__.py:16-140): the same ``numpy.random.default_rng(seed)``
draws, so every known-answer value of the reference's tests applies unchanged.  The
``nyears >= 1`` variant carries the reference's monthly calendar time axis
(test_data/time.py:44-99: the mid-points of consecutive month starts) as ``cftime`` objects
when cftime is importable and as ``cftime_lite.Datetime`` otherwise, plus a ``days_in_month``
variable (not in the reference) for callers that want the weights as numbers.
"""

import numpy as np

from .labeled import DataArray, Dataset

__all__ = ["generate_test_data", "generate_test_data_dz"]

_NOLEAP = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])


def _monthly_time_axis(start_year, ntimes, calendar):
    """Mid-points of ``ntimes`` consecutive months (test_data/time.py:66-89), an object array of calendar dates."""
    try:
        import cftime

        bounds = [cftime.datetime(start_year + k // 12, k % 12 + 1, 1, calendar=calendar) for k in range(ntimes + 1)]
    except ImportError:
        from .cftime_lite import month_starts

        bounds = month_starts(start_year, ntimes + 1, calendar)
    out = np.empty(ntimes, dtype=object)
    for k in range(ntimes):
        out[k] = bounds[k] + (bounds[k + 1] - bounds[k]) / 2
    return out


def generate_test_data(start_year=1981, nyears=0, calendar="noleap", seed=123):
    """ntimes x 5 x 5 x 5 dataset for unit testing (test_data/__init__.py:16-105)."""
    dset = Dataset()
    if nyears >= 1:
        ntimes = 12 * nyears
        dset["time"] = DataArray(_monthly_time_axis(start_year, ntimes, calendar), ("time",), attrs={
            "long_name": "time", "cartesian_axis": "T", "calendar_type": calendar, "bounds": "time_bnds"})
        dim = _NOLEAP.copy()
        years = start_year + np.arange(nyears)
        leap = np.zeros(nyears, dtype=bool)
        if calendar in ("julian",):
            leap = years % 4 == 0
        elif calendar in ("standard", "gregorian", "proleptic_gregorian"):
            leap = (years % 4 == 0) & ((years % 100 != 0) | (years % 400 == 0))
        dims = np.tile(dim, (nyears, 1))
        dims[leap, 1] = 29
        if calendar == "360_day":
            dims[:] = 30
        dset["days_in_month"] = DataArray(dims.reshape(-1).astype(np.float64), ("time",))
    else:
        ntimes = 5
        dset["time"] = DataArray(np.array([1.0, 2.0, 3.0, 4.0, 5.0]), ("time",), attrs={
            "long_name": "time", "cartesian_axis": "T", "calendar_type": calendar, "bounds": "time_bnds"})

    # tripolar/horizontal.py:62-122
    for name, long_name, units, axis in (("xh", "h point nominal longitude", "degrees_east", "X"),
                                         ("yh", "h point nominal latitude", "degrees_north", "Y")):
        dset[name] = DataArray(np.array([1.0, 2.0, 3.0, 4.0, 5.0]), (name,), attrs={
            "long_name": long_name, "units": units, "axis": axis, "cartesian_axis": axis})
    area = np.random.default_rng(seed).normal(100.0, 10.0, (5, 5))
    area = area / area.sum()
    dset["areacello"] = DataArray(area * 3.6111092e14, ("yh", "xh"), attrs={
        "long_name": "Ocean Grid-Cell Area", "units": "m2", "standard_name": "cell_area"})

    # tripolar/vertical.py:37-84
    dset["z_i"] = DataArray(np.array([0.0, 5.0, 15.0, 185.0, 1815.0, 6185.0]), ("z_i",), attrs={
        "long_name": "Depth at interface", "units": "meters", "axis": "Z", "positive": "down"})
    dset["z_l"] = DataArray(np.array([2.5, 10.0, 100.0, 1000.0, 4000.0]), ("z_l",), attrs={
        "long_name": "Depth at cell center", "units": "meters", "axis": "Z", "positive": "down", "edges": "z_i"})
    deptho = np.array([np.random.default_rng(seed).uniform(0.0, hi, 5) for hi in (5.0, 15.0, 185.0, 1815.0, 6185.0)])
    dset["deptho"] = DataArray(deptho, ("yh", "xh"), attrs={
        "long_name": "Sea Floor Depth", "units": "m", "standard_name": "sea_floor_depth_below_geoid"})

    # test_data/__init__.py:66-103: a fresh generator with the same seed for every field
    dims = ("time", "z_l", "yh", "xh")
    shape = (ntimes, 5, 5, 5)
    dset["thetao"] = DataArray(np.random.default_rng(seed).normal(15.0, 5.0, shape), dims, attrs={
        "long_name": "Sea Water Potential Temperature", "units": "degC",
        "standard_name": "sea_water_potential_temperature"})
    dset["so"] = DataArray(np.random.default_rng(seed).normal(35.0, 1.5, shape), dims, attrs={
        "long_name": "Sea Water Salinity", "units": "psu", "standard_name": "sea_water_salinity"})
    dset["volcello"] = DataArray(np.random.default_rng(seed).normal(1000.0, 100.0, shape), dims, attrs={
        "long_name": "Ocean grid-cell volume", "units": "m3", "standard_name": "ocean_volume"})
    return dset


def generate_test_data_dz(seed=123):
    """Partial-bottom-cell fixture (test_data/__init__.py:108-140)."""
    deptho = np.random.default_rng(seed).uniform(0.0, 100.0, (5, 5))
    deptho[2, 2] = np.nan
    deptho[2, 3] = np.nan
    z_i = np.array([0.0, 5.0, 10.0, 20.0, 50.0, 100.0])
    z_l = (z_i[1:] + z_i[:-1]) / 2.0
    dset = Dataset()
    dset["xh"] = DataArray(np.arange(1, 6), ("xh",))
    dset["yh"] = DataArray(np.arange(10, 60, 10), ("yh",))
    dset["deptho"] = DataArray(deptho, ("yh", "xh"))
    dset["z_l"] = DataArray(z_l, ("z_l",))
    dset["z_i"] = DataArray(z_i, ("z_i",))
    return dset
