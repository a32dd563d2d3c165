"""The xarray boundary of ``steric()``: an ``xarray.Dataset`` in gives ``xarray.Dataset`` objects out.

The real xarray is used when it is importable; this image does not have it, so a duck-typed stub
(tests/_stubs/xarray) stands in and the test covers the adapter's control flow and metadata handling.
"""

import importlib.util
import pathlib
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def xr(monkeypatch):
    if importlib.util.find_spec("xarray") is None:
        monkeypatch.syspath_prepend(str(pathlib.Path(__file__).resolve().parent / "_stubs"))
        monkeypatch.delitem(sys.modules, "xarray", raising=False)
    import xarray

    yield xarray
    if getattr(xarray, "__version__", "").endswith("stub"):
        sys.modules.pop("xarray", None)


def test_xarray_dataset_in_and_out(xr):
    import momlevel_b200 as ml

    lab = ml.test_data.generate_test_data()
    ds = xr.Dataset(attrs={"title": "config 1"})
    for name, var in lab.variables.items():
        ds[name] = xr.DataArray(np.asarray(var.values), dims=var.dims, attrs=dict(var.attrs))
    result, reference = ml.steric(ds)
    assert type(result).__module__.startswith("xarray") and type(reference).__module__.startswith("xarray")
    want, wref = ml.steric(lab)
    assert result["steric"].dims == ("time", "yh", "xh")
    assert np.array_equal(np.asarray(result["steric"].values), want["steric"].values, equal_nan=True)
    assert np.array_equal(np.asarray(result["delta_rho"].values), want["delta_rho"].values, equal_nan=True)
    assert result["steric"].attrs == {"long_name": "Steric height adjustment", "units": "m"}
    assert result["steric"].encoding["dtype"] == "float32"  # steric.py:174
    assert float(np.nansum(result["steric"].values)) == pytest.approx(1.38250197, abs=5e-9)  # tests/test_steric.py:64
    for name in ("thetao", "so", "volcello", "rho", "volo", "masso", "rhoga", "areacello"):  # tests/test_reference.py:7-19
        assert name in reference.variables
    assert float(reference["rhoga"].values) == pytest.approx(float(wref["rhoga"]), rel=1e-15)
    # the reference Dataset goes back in (steric.py:98-103) and the coordinates of the input come back out
    again, _ = ml.thermosteric(ds, reference=reference)
    assert float(np.nansum(again["thermosteric"].values)) == pytest.approx(-4.14327109, abs=5e-9)  # :76
    assert np.array_equal(np.asarray(again["time"].values), np.asarray(ds["time"].values))
    with pytest.raises(AssertionError):
        ml.steric(ds, reference=wref)  # a non-xarray reference next to an xarray dataset (steric.py:99-101)


def test_xarray_annual_average_with_a_calendar_axis(xr):
    """steric(xr_dset, annual=True) without extra arguments (tests/test_steric.py:158-163): the calendar objects of the
    time axis cross the adapter and supply years and weights."""
    import momlevel_b200 as ml

    lab = ml.test_data.generate_test_data(start_year=1983, nyears=2, calendar="julian")
    ds = xr.Dataset()
    for name, var in lab.variables.items():
        if name != "days_in_month":
            ds[name] = xr.DataArray(np.asarray(var.values), dims=var.dims, attrs=dict(var.attrs))
    result, _ = ml.steric(ds, annual=True)
    assert type(result).__module__.startswith("xarray")
    assert np.asarray(result["time"].values).shape == (2,)
    assert float(np.nansum(result["steric"].values)) == pytest.approx(1.07892738, abs=5e-9)
    assert float(np.nansum(result["delta_rho"].values)) == pytest.approx(-4.15906613, abs=5e-9)


def test_dask_backed_fields_are_never_loaded_whole(xr):
    """A Dataset whose 4-D variables are dask-like (chunked along time, as `xr.open_mfdataset(..., chunks={"time": 1})`
    gives them): the adapter keeps them in blocks, steric() streams them, and no request ever asks for more than one
    block.  Variables the path does not read are not touched at all."""
    import momlevel_b200 as ml
    from momlevel_b200 import synth

    if not hasattr(xr, "ChunkedNumpy"):
        pytest.skip("needs the stub's dask stand-in (real xarray + dask is exercised the same way by its own chunks)")
    lab = synth.make_dataset(6, 10, 16, 64, seed=12, device="cpu", dtype=__import__("torch").float32)
    ds = xr.Dataset()
    for name, var in lab.variables.items():
        vals = np.asarray(var.values)
        data = xr.ChunkedNumpy(vals, (1,) * 6) if vals.ndim == 4 else vals
        ds[name] = xr.DataArray(data, dims=var.dims, attrs=dict(var.attrs))

    class Untouchable(xr.ChunkedNumpy):
        def compute(self):
            raise AssertionError("a variable the path does not read was loaded")

    ds["uo"] = xr.DataArray(Untouchable(np.zeros((6, 10, 16, 64), np.float32), (1,) * 6), dims=lab["thetao"].dims)
    xr.ChunkedNumpy.largest_compute = 0
    result, reference = ml.steric(ds)
    one_block = 10 * 16 * 64
    assert 0 < xr.ChunkedNumpy.largest_compute <= one_block
    want, wref = ml.steric(lab)
    assert np.array_equal(np.asarray(result["steric"].values), want["steric"].values, equal_nan=True)
    assert float(reference["masso"].values) == float(wref["masso"])
    g, _ = ml.steric(ds, domain="global", variant="halosteric")
    gw, _ = ml.steric(lab, domain="global", variant="halosteric")
    assert np.array_equal(np.asarray(g["halosteric"].values), gw["halosteric"].values)
    assert xr.ChunkedNumpy.largest_compute <= one_block
