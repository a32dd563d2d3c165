"""CPU tests of the host-side mirror of the reference interface (no kernels involved).

Cases follow the reference's own tests: tests/test_util.py:48-122 (coords, area and dataset
validation) and the error paths of src/momlevel/steric.py that fire before any compute.
"""

import numpy as np
import pytest

import momlevel_b200 as momlevel
from momlevel_b200 import util
from momlevel_b200.labeled import DataArray, Dataset
from momlevel_b200.test_data import generate_test_data, generate_test_data_dz

dset = generate_test_data()


def test_default_coords_1():
    assert util.default_coords() == ("time", "z_l", "z_i")


def test_default_coords_2():
    assert util.default_coords(coord_names={"z": "lev", "t": "TIME"}) == ("TIME", "lev", "z_i")
    with pytest.raises(AssertionError):
        util.default_coords(coord_names=["z"])


def test_validate_areacello():
    assert util.validate_areacello(dset.areacello)
    assert not util.validate_areacello(dset.areacello * 1.3)


def test_validate_dataset_ok():
    util.validate_dataset(dset)
    util.validate_dataset(dset, additional_vars=["z_i", "deptho"])


def test_validate_dataset_missing_variable(capsys):
    with pytest.raises(ValueError, match="Errors found in dataset."):
        util.validate_dataset(dset.drop_vars(["thetao"]))
    assert "missing variables" in capsys.readouterr().out


def test_validate_dataset_bad_area_strict_and_lenient():
    bad = dset.copy()
    bad["areacello"] = bad["areacello"] * 1.3
    with pytest.raises(ValueError):
        util.validate_dataset(bad)
    with pytest.warns(UserWarning):
        util.validate_dataset(bad, strict=False)


def test_validate_dataset_reference_rank():
    # a 4-D dataset is not a reference state (tests/test_util.py:96-100)
    with pytest.raises(ValueError):
        util.validate_dataset(dset.copy(), reference=True)


def test_validate_dataset_reference_fields():
    ref = Dataset()
    for k in ("thetao", "so", "volcello"):
        ref[k] = dset[k].isel(time=0)
    ref["rho"] = dset["thetao"].isel(time=0)
    for k in ("volo", "masso", "rhoga"):
        ref[k] = DataArray(np.float64(1.0), ())
    ref["areacello"] = dset["areacello"]
    util.validate_dataset(ref, reference=True)
    with pytest.raises(ValueError):
        util.validate_dataset(ref.drop_vars(["rhoga"]), reference=True)
    ref["volo"] = DataArray(np.ones(3), ("x",))
    with pytest.raises(ValueError):
        util.validate_dataset(ref, reference=True)


def test_validate_dataset_additional_vars():
    with pytest.raises(ValueError):
        util.validate_dataset(dset.copy(), additional_vars=["foo", "bar"])


def test_eos_func_from_str():
    from momlevel_b200.eos import linear, wright

    assert util.eos_func_from_str("Wright") is wright.density
    assert util.eos_func_from_str("LINEAR", func_name="alpha") is linear.alpha
    with pytest.raises(ValueError, match="Unknown equation of state"):
        util.eos_func_from_str("teos10")
    with pytest.raises(AssertionError):
        util.eos_func_from_str(10)


def test_steric_rejects_bad_area_before_compute():
    # tests/test_steric.py:25-29
    bad = dset.copy()
    bad["areacello"] = bad["areacello"] * 1.3
    with pytest.raises(Exception):
        momlevel.steric(bad)


def test_wrapper_duplicate_variant_kwarg():
    # steric.py:187-196: the wrappers pass variant= themselves
    with pytest.raises(TypeError):
        momlevel.thermosteric(dset, variant="steric")


def test_labeled_basics():
    t = dset["thetao"]
    assert t.dims == ("time", "z_l", "yh", "xh") and t.shape == (5, 5, 5, 5)
    assert t.isel(time=0).dims == ("z_l", "yh", "xh")
    assert float(t[0, 1, 2, 3]) == t.values[0, 1, 2, 3]
    assert t.transpose("z_l", ...).dims == ("z_l", "time", "yh", "xh")
    a = DataArray(np.array([1.0, np.nan, 3.0]), ("x",))
    assert float(a.sum()) == 4.0 and a.notnull().values.tolist() == [True, False, True]
    ds = dset.rename({"thetao": "temp"})
    assert "temp" in ds and "thetao" not in ds
    assert dset.rename(None)["so"].dims == dset["so"].dims
    assert set(dset.coords) == {"time", "xh", "yh", "z_i", "z_l"}
    s = dset.sum()
    assert float(s["thetao"]) == pytest.approx(dset["thetao"].values.sum())
    lazy = DataArray.lazy(lambda: np.ones((2, 2)), (2, 2), ("a", "b"))
    assert lazy.is_lazy and lazy.shape == (2, 2)
    assert float(lazy.sum()) == 4.0 and not lazy.is_lazy


def test_test_data_matches_reference_draws():
    # reference sums of the t=0 slab, tests/test_steric.py:32-36
    assert dset["thetao"].values[0].sum() == pytest.approx(1921.05772939, abs=5e-9)
    assert dset["so"].values[0].sum() == pytest.approx(4388.81731882, abs=5e-9)
    assert dset["volcello"].values[0].sum() == pytest.approx(125921.15458782, abs=5e-9)
    assert float(dset["areacello"].sum()) == pytest.approx(3.6111092e14, rel=1e-12)
    dz = generate_test_data_dz()
    assert np.isnan(dz["deptho"].values[2, 2]) and dz["z_l"].values.tolist() == [2.5, 7.5, 15.0, 35.0, 75.0]
    d3 = generate_test_data(start_year=1983, nyears=2, calendar="julian")
    assert d3["thetao"].shape == (24, 5, 5, 5)
    assert d3["days_in_month"].values[[1, 13]].tolist() == [28.0, 29.0]  # 1984 is a julian leap year


def test_annual_average_weights():
    d = Dataset()
    vals = np.arange(24, dtype=np.float64)
    d["time"] = DataArray(np.arange(24.0), ("time",))
    d["x"] = DataArray(vals.reshape(24, 1) * np.ones((1, 3)), ("time", "p"))
    w = np.tile([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], 2).astype(float)
    out = util.annual_average(d, days_in_month=w)
    expect = [(vals[:12] * w[:12]).sum() / w[:12].sum(), (vals[12:] * w[12:]).sum() / w[12:].sum()]
    assert out["x"].shape == (2, 3)
    assert np.allclose(out["x"].values[:, 0], expect, rtol=1e-15)


def test_dataset_hands_out_index_coordinates_and_rename_identity():
    # like xarray, a variable taken out of a Dataset knows the values of its dimension coordinates
    # (derived.calc_n2 reads thetao[zcoord], derived.py:396)
    t = dset["thetao"]
    assert set(t.coords) >= {"time", "z_l", "yh", "xh"}
    assert np.array_equal(np.asarray(t.coords["z_l"].values), np.asarray(dset["z_l"].values))
    assert "z_l" not in dset["z_l"].coords  # an index variable does not carry itself
    assert dset.rename(None) is dset and dset.rename({}) is dset
    renamed = dset.rename({"thetao": "temp"})
    assert "temp" in renamed.variables and "thetao" not in renamed.variables and "thetao" in dset.variables


def test_column_diagnostics_argument_errors_fire_before_any_compute():
    from momlevel_b200 import derived

    bare = DataArray(np.zeros((2, 3, 4)), ("time", "z_l", "xh"))  # no level values attached
    with pytest.raises(KeyError):
        derived.calc_n2(bare, bare)
    with pytest.raises(NotImplementedError):
        derived.calc_n2(dset["thetao"], dset["so"], interfaces=dset["z_i"])  # needs xgcm (derived.py:391-395)
    with pytest.raises(ValueError):
        derived.calc_n2(dset["thetao"], dset["so"], eos="teos10")  # util.py:243-249
    with pytest.raises(ValueError):
        derived.calc_n2(dset["thetao"], dset["so"][0])


def test_steric_variants_validates_like_steric():
    with pytest.raises(ValueError, match="Errors found in dataset."):
        momlevel.steric_variants(dset.drop_vars(["deptho"]))
    bad = dset.copy()
    bad["areacello"] = dset["areacello"] * 2.0
    with pytest.raises(ValueError, match="Errors found in dataset."):
        momlevel.steric_variants(bad)
    with pytest.raises(ValueError, match="Unknown equation of state"):
        momlevel.steric_variants(dset, equation_of_state="teos10")


def test_bind_host_to_device_is_harmless_without_a_gpu():
    import os

    from momlevel_b200 import distributed

    before = os.sched_getaffinity(0)
    n = distributed.bind_host_to_device(0)
    assert n == 0 or n == len(os.sched_getaffinity(0))
    if n == 0:
        assert os.sched_getaffinity(0) == before


def test_steric_routes_host_resident_fields_to_the_host_entry(monkeypatch):
    """steric() hands numpy / CPU-tensor fields to ``core.steric_local_host`` (the streaming entry point) with
    host arrays for the grid, asks for the one extra height a variant needs, and assembles the same result and
    reference Datasets around what comes back.  The library call itself is replaced: no device here."""
    import importlib

    import torch

    import momlevel_b200 as ml
    from momlevel_b200 import core, synth

    steric_mod = importlib.import_module("momlevel_b200.steric")
    small = synth.make_dataset(5, 12, 20, 32, seed=5, device="cpu", dtype=torch.float32)
    assert not steric_mod._host_resident(small, "time", "z_l", "z_i")  # too small to be worth streaming
    monkeypatch.setattr(steric_mod, "HOST_ROUTE_MIN_BYTES", 0)
    assert steric_mod._host_resident(small, "time", "z_l", "z_i")
    calls = []

    def fake(T, S, V0, z_i, depth, pres, rhozero=1035.0, eos="Wright", steps_per_window=1, want_rho_ref=False,
             eta_out=None, variants=False):
        calls.append((tuple(V0.shape), type(z_i).__name__, type(depth).__name__, type(pres).__name__,
                      steps_per_window, variants, eos, rhozero))
        eta = torch.zeros((T.shape[0],) + tuple(T.shape[2:]), dtype=torch.float64)
        if not variants:
            return eta, None, (2.0, 2070.0)
        names = ("thermosteric", "halosteric") if variants is True else variants
        return {"steric": eta, **{v: eta + 1.0 for v in names}}, None, (2.0, 2070.0)

    monkeypatch.setattr(core, "steric_local_host", fake)
    for variant, level in (("steric", 0.0), ("thermosteric", 1.0), ("halosteric", 1.0)):
        res, ref = ml.steric(small, variant=variant, rhozero=1030.0)
        assert res[variant].shape == (5, 20, 32) and float(res[variant].values.max()) == level
        assert res[variant].attrs["long_name"] == f"{variant.capitalize()} height adjustment"
        assert res["delta_rho"].shape == (5, 12, 20, 32)  # lazy: not evaluated here
        assert float(ref["volo"]) == 2.0 and float(ref["masso"]) == 2070.0 and float(ref["rhoga"]) == 1035.0
        assert set(ref.variables) >= {"thetao", "so", "volcello", "rho", "volo", "masso", "rhoga", "areacello"}
    assert [c[5] for c in calls] == [False, ("thermosteric",), ("halosteric",)]
    assert all(c[0] == (12, 20, 32) and c[1:4] == ("ndarray",) * 3 and c[4] == 5 and c[6] == "Wright"
               and c[7] == 1030.0 for c in calls)
    # steric_variants: one host call for the three heights
    calls.clear()
    res, ref = ml.steric_variants(small)
    assert [c[5] for c in calls] == [True] and calls[0][4] == 5
    assert float(res["steric"].values.max()) == 0.0 and float(res["halosteric"].values.min()) == 1.0
    assert float(ref["rhoga"]) == 1035.0
    # the global domain: the masses of host-resident fields come from the streaming entry point as well
    from oracle import steric as osteric

    f64 = lambda k: small[k].values.astype(np.float64)  # noqa: E731
    o = osteric.reference_state(f64("thetao"), f64("so"), f64("volcello"), f64("areacello"), f64("z_l"))
    ref_in = ml.Dataset()
    for k in ("thetao", "so", "volcello", "rho"):
        ref_in[k] = ml.DataArray(o[k], ("z_l", "yh", "xh"))
    for k in ("volo", "masso", "rhoga"):
        ref_in[k] = ml.DataArray(np.float64(o[k]), ())
    ref_in["areacello"] = small["areacello"]
    gcalls = []

    class FakeStream:  # stands in for core.HostStream (ml_host_stream_*): records what steric() hands it
        def __init__(self, domain, v_ref, p_level, variants=("steric",), reference=None, eos="Wright", max_block_steps=1,
                     dtype=torch.float32, **kw):
            self.variants = variants
            gcalls.append(["begin", domain, tuple(np.shape(v_ref)), type(p_level).__name__, variants,
                           None if reference is None else sorted(reference), eos, max_block_steps])

        def push(self, T, S):
            gcalls.append(["push", tuple(T.shape)])
            return {v: torch.full((T.shape[0],), float(o["masso"]), dtype=torch.float64) for v in self.variants}

        def finish(self):
            gcalls.append(["finish"])
            return None, None

        def abort(self):
            gcalls.append(["abort"])

    monkeypatch.setattr(core, "HostStream", FakeStream)
    calls.clear()
    res, _ = ml.steric(small, domain="global", reference=ref_in)
    assert gcalls == [["begin", "global", (12, 20, 32), "ndarray", ("steric",), None, "Wright", 5],
                      ["push", (5, 12, 20, 32)], ["finish"]] and calls == []
    assert res["steric"].shape == (5,) and np.allclose(res["steric"].values, 0.0, atol=1e-12)  # M(t) = M_ref
    assert float(res["reference_height"]) == pytest.approx(o["volo"] / small["areacello"].values.sum())
    # thermosteric / halosteric in the global domain hold one field at its reference slab, which the stream keeps on
    # the device for the whole series (steric.py:115-121)
    gcalls.clear()
    res, _ = ml.steric(small, domain="global", reference=ref_in, variant="thermosteric")
    assert gcalls[0] == ["begin", "global", (12, 20, 32), "ndarray", ("thermosteric",), ["so", "thetao"], "Wright", 5]
    assert res["thermosteric"].shape == (5,)


def test_annual_average_reads_a_calendar_time_axis():
    """util.py:49-119 as the reference calls it (tests/test_steric.py:158-163): no weights argument -- year and
    days-in-month of every step come from the calendar objects on the time axis, each mean is labelled with the
    mid-point of its year, non-numeric variables are dropped."""
    import datetime

    from momlevel_b200 import test_data
    from momlevel_b200.cftime_lite import Datetime

    d3 = test_data.generate_test_data(start_year=1983, nyears=2, calendar="julian")
    t = d3["time"].values
    assert t.dtype == object and t[0].year == 1983 and t[0].month == 1 and t[13].daysinmonth == 29
    years, dim = util.calendar_axis(t)
    assert years.tolist() == [1983] * 12 + [1984] * 12 and util.whole_years_in_order(years)
    assert np.array_equal(dim, d3["days_in_month"].values)
    d = Dataset()
    d["time"] = d3["time"]
    vals = np.arange(24, dtype=np.float64)
    d["x"] = DataArray(vals.reshape(24, 1) * np.ones((1, 3)), ("time", "p"))
    d["label"] = DataArray(np.array(["m%d" % i for i in range(24)], dtype=object), ("time",))
    out = util.annual_average(d)
    expect = [(vals[:12] * dim[:12]).sum() / dim[:12].sum(), (vals[12:] * dim[12:]).sum() / dim[12:].sum()]
    assert np.allclose(out["x"].values[:, 0], expect, rtol=1e-15)
    assert "label" not in out.variables  # util.py:72-73
    mids = out["time"].values
    assert len(mids) == 2 and (mids[0].year, mids[0].month, mids[0].day, mids[0].hour) == (1983, 7, 2, 12)
    assert (mids[1].year, mids[1].month, mids[1].day, mids[1].hour) == (1984, 7, 2, 0)  # 366 days: mid-point at midnight
    # the reference groups by year wherever the steps are: a shuffled axis gives the same means
    perm = np.random.default_rng(0).permutation(24)
    ds = Dataset()
    ds["time"] = DataArray(t[perm], ("time",))
    ds["x"] = DataArray(d["x"].values[perm], ("time", "p"))
    assert np.allclose(util.annual_average(ds)["x"].values[:, 0], expect, rtol=1e-15)
    # a year with fewer than twelve steps is an error, as in the reference (util.py:82)
    short = Dataset()
    short["time"] = DataArray(t[:23], ("time",))
    short["x"] = DataArray(np.zeros((23, 1)), ("time", "p"))
    with pytest.raises(AssertionError):
        util.annual_average(short)
    # no calendar and no weights: nothing to weight with
    bare = Dataset()
    bare["time"] = DataArray(np.arange(24.0), ("time",))
    bare["x"] = DataArray(np.zeros((24, 1)), ("time", "p"))
    with pytest.raises(ValueError):
        util.annual_average(bare)
    # numpy datetime64 axes carry a (proleptic Gregorian) calendar too
    t64 = np.array([f"{y}-{m:02d}-15" for y in (2023, 2024) for m in range(1, 13)], dtype="datetime64[D]")
    y64, d64 = util.calendar_axis(t64)
    assert y64[0] == 2023 and d64[1] == 28 and d64[13] == 29
    # calendar arithmetic of the stand-in class
    a = Datetime(1900, 2, 28, calendar="gregorian") + datetime.timedelta(days=1)
    assert (a.month, a.day) == (3, 1)
    b = Datetime(1900, 2, 28, calendar="julian") + datetime.timedelta(days=1)
    assert (b.month, b.day) == (2, 29)
    assert Datetime(2001, 1, 1, calendar="360_day") - Datetime(2000, 1, 1, calendar="360_day") == datetime.timedelta(days=360)


# ---------------------------------------------------------------- the table the ranks of a host share (csrc/ml_hostpath.cu)

def _share_worker(rank, ranks, port, tables, out):
    import ctypes

    from momlevel_b200 import _lib

    L = _lib.lib()
    ms = (ctypes.c_double * 4)(*tables[rank])
    got = (ctypes.c_double * 4)()
    seen = L.ml_host_tuner_share_selftest(port.encode(), rank, ranks, 75, 1555200, ctypes.cast(ms, ctypes.c_void_p),
                                          ctypes.cast(got, ctypes.c_void_p), 20000)
    out.put((rank, seen, list(got)))


def test_ranks_of_a_host_see_the_same_combined_table():
    """Every rank picks its packing threads from the element-wise maximum of all ranks' timings, so they pick the same:
    three processes publish different tables through the POSIX shared-memory segment and must read identical maxima;
    a choice one of them has not timed yet stays unknown for all."""
    import multiprocessing as mp
    import os

    ctx = mp.get_context("spawn")
    ranks = 3
    tables = [[20.0, 25.0, 19.0, 22.0], [21.0, 24.0, 30.0, 22.5], [18.0, 26.0, 17.0, -1.0]]
    out = ctx.Queue()
    port = f"selftest-{os.getpid()}"
    procs = [ctx.Process(target=_share_worker, args=(r, ranks, port, tables, out)) for r in range(ranks)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in range(ranks))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, seen, got in results:
        assert seen == ranks, (rank, seen)
        assert got == [21.0, 26.0, 30.0, -1.0], (rank, got)


def test_checked_grid_entries_stand_for_the_tensors_they_were_made_on():
    """steric._checked_before: an entry is void after an in-place write to one of its tensors, for another tensor
    object in its place (equal values or not), and once a tensor is gone (its id may be handed out again)."""
    import gc
    import importlib

    import torch

    st = importlib.import_module("momlevel_b200.steric")  # the package exports the function under the same name

    st._CHECKED_GRIDS.clear()
    arrs = tuple(torch.arange(4, dtype=torch.float64) + i for i in range(4))
    assert st._checked_before(arrs) is None
    st._remember_checked(arrs, 3.6e14)
    assert st._checked_before(arrs) == 3.6e14
    assert st._checked_before(arrs[:3] + (arrs[3].clone(),)) is None  # another object, same values
    arrs[1][0] = -1.0                                                 # in-place write: _version moves
    assert st._checked_before(arrs) is None
    st._remember_checked(arrs, 1.0)
    assert st._checked_before(arrs) == 1.0
    key = tuple(id(a) for a in arrs)
    kept = arrs[1:]
    del arrs
    gc.collect()
    refs, _, _ = st._CHECKED_GRIDS[key]
    assert refs[0]() is None  # the entry cannot be mistaken for a new tensor that lands on the same id
    for _ in range(40):  # the table does not grow without bound
        st._remember_checked(tuple(torch.zeros(1) for _ in range(4)), 0.0)
    assert len(st._CHECKED_GRIDS) <= 16
    del kept
