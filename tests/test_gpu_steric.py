"""GPU parity of the fused steric path against the oracle and the reference's goldens.

Tolerances (BASELINE.json north_star): 1e-10 relative on density, 1e-9 m absolute on steric
height, against the reference's numpy path on the same fp64-upcast inputs.
Every computation goes through libmomlevel_b200's C ABI.
"""

import numpy as np
import pytest
import torch

from oracle import steric as osteric
from oracle import testdata

pytestmark = pytest.mark.gpu

ETA_ATOL = 1e-9  # m
RHO_RTOL = 1e-10


@pytest.fixture(scope="module")
def ml():
    import momlevel_b200

    assert torch.cuda.is_available(), "GPU tests need a CUDA device: there is no CPU path"
    return momlevel_b200


@pytest.fixture(scope="module")
def dset(ml):
    return ml.test_data.generate_test_data()


def _close_nan(a, b, atol=0.0, rtol=0.0):
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    m = ~np.isnan(b)
    if m.any():
        err = np.abs(a[m] - b[m])
        assert np.all(err <= atol + rtol * np.abs(b[m])), f"max err {err.max():.3e}"


# ------------------------------------------------------ config 1: the reference's own tests


def test_steric_broadcast(ml, dset):
    # tests/test_steric.py:13-22
    result, reference = ml.steric(dset)
    ref = float(reference["rho"][1, 2, 3])
    rho = ml.eos.wright.density(float(dset["thetao"][0, 1, 2, 3]), float(dset["so"][0, 1, 2, 3]),
                                (float(dset["z_l"][1]) * 1.0e4) + 101325.0)
    assert np.allclose(ref, rho)
    assert ref == pytest.approx(float(rho), rel=1e-15)


REFERENCE_RESULTS = {  # tests/test_steric.py:32-41
    "reference_thetao": 1921.05772939,
    "reference_so": 4388.81731882,
    "reference_vol": 125921.15458782,
    "reference_rho": 128781.63975736,
    "global_reference_vol": 125921.15458782,
    "global_reference_rho": 1030.2309221,
}


def _check_reference(reference):
    reference = reference.sum()
    assert float(reference["thetao"]) == pytest.approx(REFERENCE_RESULTS["reference_thetao"], abs=5e-9)
    assert float(reference["so"]) == pytest.approx(REFERENCE_RESULTS["reference_so"], abs=5e-9)
    assert float(reference["volcello"]) == pytest.approx(REFERENCE_RESULTS["reference_vol"], abs=5e-9)
    assert float(reference["rho"]) == pytest.approx(REFERENCE_RESULTS["reference_rho"], abs=5e-9)
    assert float(reference["volo"]) == pytest.approx(REFERENCE_RESULTS["global_reference_vol"], abs=5e-9)
    assert float(reference["rhoga"]) == pytest.approx(REFERENCE_RESULTS["global_reference_rho"], abs=5e-8)


@pytest.mark.parametrize(
    "func,variant,eta_sum,drho_sum",
    [
        ("steric", "steric", 1.38250197, -11.33133173),  # tests/test_steric.py:56-65
        ("thermosteric", "thermosteric", -4.14327109, 33.83631611),  # :68-77
        ("halosteric", "halosteric", 4.39398075, -32.07946717),  # :44-53
    ],
)
def test_local_values(ml, dset, func, variant, eta_sum, drho_sum):
    result, reference = getattr(ml, func)(dset)
    assert result[variant].dims == ("time", "yh", "xh")
    assert result["delta_rho"].dims == ("time", "z_l", "yh", "xh")
    assert result[variant].attrs == {"long_name": f"{variant.capitalize()} height adjustment", "units": "m"}
    assert result["delta_rho"].attrs["units"] == "kg m-3"
    _check_reference(reference)
    summed = result.sum()
    assert float(summed[variant]) == pytest.approx(eta_sum, abs=5e-9)
    assert float(summed["delta_rho"]) == pytest.approx(drho_sum, abs=5e-9)
    # and field by field against the oracle, at the stated tolerances
    o = testdata.generate_test_data()
    oref = osteric.reference_state(o["thetao"], o["so"], o["volcello"], o["areacello"], o["z_l"])
    eta, drho = osteric.steric_local(o["thetao"], o["so"], o["z_l"], o["z_i"], o["deptho"], oref, variant=variant)
    _close_nan(result[variant].values, eta, atol=ETA_ATOL)
    _close_nan(result["delta_rho"].values, drho, atol=RHO_RTOL * 1030.0)
    _close_nan(reference["rho"].values, oref["rho"], rtol=RHO_RTOL)


@pytest.mark.parametrize("func,variant", [("steric", "steric"), ("thermosteric", "thermosteric"),
                                          ("halosteric", "halosteric")])
def test_global_values(ml, dset, func, variant):
    # tests/test_steric.py:80-125.  The constants there are below the reference's own atol
    # and pin nothing; the oracle (= the reference code, steric.py:134-147) is the check.
    result, reference = getattr(ml, func)(dset, domain="global")
    _check_reference(reference)
    assert result[variant].dims == ("time",) and result["reference_height"].dims == ()
    o = testdata.generate_test_data()
    oref = osteric.reference_state(o["thetao"], o["so"], o["volcello"], o["areacello"], o["z_l"])
    eta, href, masso = osteric.steric_global(o["thetao"], o["so"], o["z_l"], oref, variant=variant)
    assert float(result["reference_height"]) == pytest.approx(href, rel=1e-13)
    # eta_global = href * ln(.) with href ~ 3.5e-10 m here: compare the log argument instead
    got_ratio = np.exp(result[variant].values / float(result["reference_height"]))
    assert np.allclose(got_ratio, np.exp(eta / href), rtol=1e-13, atol=0)
    assert np.allclose(result[variant].values, eta, rtol=0, atol=ETA_ATOL)
    # the reference's vacuous checks still pass
    stale = {"steric": 6.29048941e-14, "thermosteric": -1.38053154e-13, "halosteric": 1.98293992e-13}[variant]
    assert np.allclose(float(result.sum()[variant]), stale)
    assert np.allclose(float(result["reference_height"]), 3.4726688e-10)


def test_steric_read_reference(ml, dset, capsys):
    # tests/test_steric.py:128-137
    dset2 = ml.test_data.generate_test_data(seed=999)
    _, reference = ml.steric(dset2)
    result, reference = ml.steric(dset, verbose=True, reference=reference)
    assert "Using supplied reference state" in capsys.readouterr().out
    rs = reference.sum()
    assert float(rs["thetao"]) == pytest.approx(1917.31113456, abs=5e-9)
    assert float(rs["so"]) == pytest.approx(4387.69334037, abs=5e-9)
    assert float(rs["volcello"]) == pytest.approx(125846.22269117, abs=5e-9)
    assert float(rs["rho"]) == pytest.approx(128780.12974804, abs=5e-9)
    assert float(result.sum()["steric"]) == pytest.approx(1.25554742, abs=5e-9)
    with pytest.raises(AssertionError):
        ml.steric(dset, reference={"rho": 1})


def test_encoding(ml, dset):
    # tests/test_steric.py:140-155
    result, _ = ml.steric(dset)
    assert result["delta_rho"].encoding["dtype"] == "float32" and result["steric"].encoding["dtype"] == "float32"
    result, _ = ml.steric(dset, dtype="float64")
    assert result["delta_rho"].encoding["dtype"] == "float64" and result["steric"].encoding["dtype"] == "float64"
    result, _ = ml.steric(dset, domain="global")
    assert result["reference_height"].encoding["dtype"] == "float32"
    result, _ = ml.steric(dset, domain="global", dtype="float64")
    assert result["steric"].encoding["dtype"] == "float64"


def test_steric_annual_average(ml):
    # tests/test_steric.py:158-163, called as the reference calls it: julian calendar, 1983-1984, the weights and the
    # year groups are read from the calendar objects on the time axis (util.py:79-87)
    dset3 = ml.test_data.generate_test_data(start_year=1983, nyears=2, calendar="julian")
    result, _ = ml.steric(dset3, annual=True)
    assert len(result["time"]) == 2
    assert [(t.year, t.month, t.day) for t in result["time"].values] == [(1983, 7, 2), (1984, 7, 2)]  # util.py:96-102
    assert result["delta_rho"].shape == (2, 5, 5, 5) and result["steric"].shape == (2, 5, 5)
    summed = result.sum()
    assert float(summed["steric"]) == pytest.approx(1.07892738, abs=5e-9)
    assert float(summed["delta_rho"]) == pytest.approx(-4.15906613, abs=5e-9)
    # explicit weights (the extension for time axes without a calendar) give the same numbers
    again, _ = ml.steric(dset3, annual=True, days_in_month=dset3["days_in_month"].values)
    assert np.array_equal(again["steric"].values, result["steric"].values, equal_nan=True)
    assert np.array_equal(again["delta_rho"].values, result["delta_rho"].values, equal_nan=True)
    # the global series is averaged the same way (steric.py:181-182)
    g, _ = ml.steric(dset3, annual=True, domain="global")
    assert g["steric"].shape == (2,) and len(g["time"]) == 2


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_annual_delta_rho_is_the_weighted_mean_of_the_monthly_field(ml, dtype):
    """ml_delta_rho_annual == xarray's weighted(days_in_month).mean over each year of ml_delta_rho's field
    (util.py:84-87): missing months are skipped and the weights renormalised per cell."""
    from momlevel_b200 import core, synth

    shape = (24, 6, 9, 32) if dtype == torch.float32 else (12, 5, 7, 13)
    ds = synth.make_dataset(*shape, seed=8, device="cuda", dtype=dtype)
    T, S = ds["thetao"].data.clone(), ds["so"].data.clone()
    V = ds["volcello"].data[0]
    pres = ds["z_l"].values * 1.0e4 + 101325.0
    rho_ref, _ = core.reference_state(T[0], S[0], V, pres)
    wet = torch.nonzero(torch.isfinite(T).all(0) & torch.isfinite(S).all(0) & torch.isfinite(V))
    (z1, y1, x1), (z2, y2, x2) = wet[0].tolist(), wet[-1].tolist()
    T[3, z1, y1, x1] = float("nan")       # one month missing in a wet cell
    S[:12, z2, y2, x2] = float("nan")     # a whole year missing in another
    w = np.tile(np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], dtype=np.float64), shape[0] // 12)
    got = core.delta_rho_annual(T, S, rho_ref, V, pres, w).cpu().numpy()
    monthly = core.delta_rho(T, S, rho_ref, V, pres).cpu().numpy().reshape((shape[0] // 12, 12) + shape[1:])
    ww = w.reshape(-1, 12)[:, :, None, None, None]
    valid = ~np.isnan(monthly)
    with np.errstate(invalid="ignore"):
        want = np.where(valid, monthly, 0.0).__mul__(ww).sum(1) / (valid * ww).sum(1)
    assert got.shape == want.shape
    _close_nan(got, want, atol=1e-12)
    assert np.isfinite(got[0, z1, y1, x1]) and np.isnan(got[0, z2, y2, x2])
    assert shape[0] == 12 or np.isfinite(got[1, z2, y2, x2])


def test_errors(ml, dset):
    with pytest.raises(ValueError, match="Unknown variant"):
        ml.steric(dset, variant="barosteric")
    with pytest.raises(ValueError, match="Unknown equation of state"):
        ml.steric(dset, equation_of_state="teos10")
    bad = dset.copy()
    bad["areacello"] = bad["areacello"] * 1.3
    with pytest.raises(Exception):
        ml.steric(bad)
    with pytest.warns(UserWarning):
        ml.steric(bad, strict=False)
    neg = dset.copy()
    neg["deptho"] = neg["deptho"] * -1.0
    with pytest.raises(AssertionError):
        ml.steric(neg)


def test_renaming(ml, dset):
    d = dset.rename({"thetao": "temp", "so": "salt", "z_l": "lev", "z_i": "ilev", "time": "T"})
    result, _ = ml.steric(d, varname_map={"temp": "thetao", "salt": "so"},
                          coord_names={"z": "lev", "zbounds": "ilev", "t": "T"})
    assert result["steric"].dims == ("T", "yh", "xh")
    assert float(result.sum()["steric"]) == pytest.approx(1.38250197, abs=5e-9)


# -------------------------------------------------------- synthetic ocean-like states


def _oracle_case(ds, variant, eos):
    f64 = lambda k: ds[k].values.astype(np.float64)  # noqa: E731 -- the parity definition: fp64 upcast
    ref = osteric.reference_state(f64("thetao"), f64("so"), f64("volcello"), f64("areacello"), f64("z_l"), eos=eos)
    eta, drho = osteric.steric_local(f64("thetao"), f64("so"), f64("z_l"), f64("z_i"), f64("deptho"), ref,
                                     eos=eos, variant=variant)
    g, href, masso = osteric.steric_global(f64("thetao"), f64("so"), f64("z_l"), ref, eos=eos, variant=variant)
    return ref, eta, drho, g, href, masso


SHAPES = [
    (3, 10, 37, 53),    # ragged: ncol = 1961, not a multiple of 4
    (14, 9, 16, 64),    # aligned, more steps than one register chunk
    (2, 75, 24, 128),   # full 75-level column
    (1, 4, 3, 5),       # single step, tiny
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant", ["steric", "thermosteric", "halosteric"])
@pytest.mark.parametrize("eos", ["Wright", "linear"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("force_direct", [False, True])
def test_synthetic_parity(ml, shape, variant, eos, dtype, force_direct):
    from momlevel_b200 import synth

    if dtype == torch.float64 and (variant != "steric" or eos == "linear") and shape != SHAPES[0]:
        pytest.skip("fp64 storage is covered on one shape for the non-default variants")
    ds = synth.make_dataset(*shape, seed=11, device="cuda", dtype=dtype)
    prev = ml.core.force_direct(force_direct)
    try:
        result, reference = ml.steric(ds, variant=variant, equation_of_state=eos)
        gres, _ = ml.steric(ds, variant=variant, equation_of_state=eos, domain="global", reference=reference)
        drho = result["delta_rho"].values
    finally:
        ml.core.force_direct(prev)
    ref, eta, odrho, g, href, masso = _oracle_case(ds, variant, eos)
    _close_nan(reference["rho"].values, ref["rho"], rtol=RHO_RTOL)
    assert float(reference["volo"]) == pytest.approx(ref["volo"], rel=1e-12)
    assert float(reference["masso"]) == pytest.approx(ref["masso"], rel=1e-12)
    _close_nan(result[variant].values, eta, atol=ETA_ATOL)
    _close_nan(drho, odrho, atol=RHO_RTOL * 1030.0)
    assert float(gres["reference_height"]) == pytest.approx(href, rel=1e-12)
    assert np.allclose(gres[variant].values, g, rtol=0, atol=ETA_ATOL)
    # M(t) itself to fp64 reduction accuracy
    got_m = ref["volo"] * ref["rhoga"] / np.exp(gres[variant].values / href)
    assert np.allclose(got_m, masso, rtol=1e-12, atol=0)


def test_nan_semantics(ml):
    """Land, partial columns and transient holes: follows steric.py:151-166 literally."""
    o = testdata.generate_test_data()
    T, S, V = o["thetao"].copy(), o["so"].copy(), o["volcello"].copy()
    V[:, :, 0, 0] = T[:, :, 0, 0] = S[:, :, 0, 0] = np.nan          # land column
    V[:, 3:, 1, 1] = T[:, 3:, 1, 1] = S[:, 3:, 1, 1] = np.nan       # shallow column
    T[2, 1, 2, 2] = np.nan                                           # hole in T only, one step
    V[:, 2, 3, 3] = np.nan                                           # volume missing, T/S present
    V[:, 0, 4, 4] = np.nan                                           # surface volume missing only
    d = ml.test_data.generate_test_data()
    d["thetao"], d["so"], d["volcello"] = (ml.DataArray(x, d["thetao"].dims) for x in (T, S, V))
    oref = osteric.reference_state(T, S, V, o["areacello"], o["z_l"])
    for variant in ("steric", "thermosteric", "halosteric"):
        eta, drho = osteric.steric_local(T, S, o["z_l"], o["z_i"], o["deptho"], oref, variant=variant)
        g, href, masso = osteric.steric_global(T, S, o["z_l"], oref, variant=variant)
        for direct in (False, True):
            prev = ml.core.force_direct(direct)
            try:
                result, reference = ml.steric(d, variant=variant)
                gres, _ = ml.steric(d, variant=variant, domain="global")
                _close_nan(result[variant].values, eta, atol=ETA_ATOL)
                _close_nan(result["delta_rho"].values, drho, atol=1e-7)
                assert np.allclose(gres[variant].values, g, rtol=0, atol=ETA_ATOL)
            finally:
                ml.core.force_direct(prev)
    assert np.all(np.isnan(eta[:, 0, 0])) and np.all(np.isnan(eta[:, 4, 4])) and np.all(np.isfinite(eta[:, 1, 1]))


def test_nan_semantics_tma_family(ml):
    """The same literal reading of steric.py:151-166 on a grid the TMA family takes, with masks that
    disagree with the bathymetry, so the depth-sorted tile cannot rely on either being consistent."""
    from momlevel_b200 import core, synth

    shape = (13, 12, 16, 64)  # ncol = 1024: four tiles; 13 steps: a 12-step chunk and a 1-step tail
    ds = synth.make_dataset(*shape, seed=5, device="cpu", dtype=torch.float32)
    T, S = ds["thetao"].data.clone(), ds["so"].data.clone()
    V = ds["volcello"].data[0].clone()
    depth = ds["deptho"].data.clone()
    g = torch.Generator().manual_seed(7)
    ny, nx = shape[2], shape[3]
    for _ in range(40):  # holes in T or S at single cells and steps (wet or not)
        t, z, y, x = (int(torch.randint(0, n, (1,), generator=g)) for n in shape)
        (T if _ % 2 else S)[t, z, y, x] = float("nan")
    for _ in range(20):  # reference volume missing where there is water, present where there is none
        z, y, x = (int(torch.randint(0, n, (1,), generator=g)) for n in shape[1:])
        V[z, y, x] = float("nan") if torch.isfinite(V[z, y, x]) else 1.0e9
    for _ in range(12):  # bathymetry that disagrees with the masks: land with data, deep water over missing data
        y, x = int(torch.randint(0, ny, (1,), generator=g)), int(torch.randint(0, nx, (1,), generator=g))
        depth[y, x] = float("nan") if torch.isfinite(depth[y, x]) else 3000.0
    V[0, 3, 7] = float("nan")  # surface volume missing: eta is masked (steric.py:166)
    d = ml.Dataset()
    for k in ("time", "z_l", "z_i", "yh", "xh", "areacello"):
        d[k] = ds[k]
    dims = ("time", "z_l", "yh", "xh")
    d["thetao"], d["so"] = ml.DataArray(T.cuda(), dims), ml.DataArray(S.cuda(), dims)
    d["volcello"] = ml.DataArray(V.cuda().unsqueeze(0).expand(shape[0], -1, -1, -1), dims)
    d["deptho"] = ml.DataArray(depth.cuda(), ("yh", "xh"))
    T64, S64, V64 = T.double().numpy(), S.double().numpy(), V.double().numpy()
    V4 = np.broadcast_to(V64, T64.shape)
    z_l, z_i = ds["z_l"].values, ds["z_i"].values
    oref = osteric.reference_state(T64, S64, V4, ds["areacello"].values, z_l)
    for variant in ("steric", "thermosteric", "halosteric"):
        eta, drho = osteric.steric_local(T64, S64, z_l, z_i, depth.numpy(), oref, variant=variant)
        gser, href, masso = osteric.steric_global(T64, S64, z_l, oref, variant=variant)
        result, reference = ml.steric(d, variant=variant)
        assert core.last_path() == 2, "expected the TMA family"
        _close_nan(reference["rho"].values, oref["rho"], rtol=RHO_RTOL)
        assert float(reference["masso"]) == pytest.approx(oref["masso"], rel=1e-12)
        _close_nan(result[variant].values, eta, atol=ETA_ATOL)
        again, _ = ml.steric(d, variant=variant, reference=reference)
        assert core.last_path() == 2
        _close_nan(again[variant].values, eta, atol=ETA_ATOL)
        gres, _ = ml.steric(d, variant=variant, domain="global", reference=reference)
        assert core.last_path() == 2
        assert np.allclose(gres[variant].values, gser, rtol=0, atol=ETA_ATOL)


@pytest.mark.parametrize("nt", [1, 4, 5, 8, 9, 12, 13, 16, 21, 24, 25, 37])
def test_time_axis_partition(ml, nt):
    """Every way the TMA family cuts the time axis (12-step chunks + one remainder chunk of 4, 8 or 12)
    gives what the direct family gives for the same call."""
    from momlevel_b200 import core, synth

    grid = synth.make_grid(6, 8, 64, seed=2, device="cuda")  # 512 columns: two tiles
    T, S, V = synth.make_fields(grid, nt, seed=40 + nt, dtype=torch.float32)
    pres = grid["z_l"] * 1.0e4 + 101325.0
    z_i, depth = grid["z_i"], grid["deptho"]
    out = {}
    for direct in (False, True):
        prev = core.force_direct(direct)
        try:
            eta, rho, sums = core.steric_local_selfref(T, S, V, z_i, depth, pres)
            assert core.last_path() == (1 if direct else 2)
            eta2, _ = core.steric_local(T, S, rho, V, z_i, depth, pres)
            masso = core.steric_global(T, S, V, pres)
            out[direct] = (eta, rho, sums, eta2, masso)
        finally:
            core.force_direct(prev)
    for a, b in zip(out[False], out[True]):
        assert a.shape == b.shape
        assert torch.equal(torch.isnan(a), torch.isnan(b))
        err = float(torch.nan_to_num(a - b).abs().max())
        scale = float(torch.nan_to_num(b).abs().max())
        # heights: the two families round rho - rho_ref differently (one FMA against two operations), which
        # shows as ~1e-12 m where the height itself is a rounding residue; everything else to fp64 accuracy
        assert err <= 1e-11 + 1e-13 * scale


@pytest.mark.parametrize("shape,dtype", [((13, 12, 16, 64), torch.float32),   # TMA family, 4-step chunks + a short one
                                         ((5, 75, 8, 96), torch.float32),     # full 75-level column
                                         ((3, 10, 37, 53), torch.float32),    # ragged rows: one rank-1-map launch per height
                                         ((3, 10, 7, 31), torch.float32),     # fewer columns than a tile: the direct family
                                         ((6, 9, 16, 64), torch.float64),     # fp64 storage: one TMA launch per height
                                         ((9, 9, 5, 59), torch.float64)])     # fp64, odd ncol: rank-1 maps
@pytest.mark.parametrize("eos", ["Wright", "linear"])
def test_all_variants_in_one_pass(ml, shape, dtype, eos):
    """steric_variants == steric + thermosteric + halosteric called one by one, and the oracle."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(*shape, seed=31, device="cuda", dtype=dtype)
    ds["thetao"].data[shape[0] // 2, 1, 3, 5] = float("nan")  # a hole at one step
    res, ref = ml.steric_variants(ds, equation_of_state=eos)
    path = core.last_path()
    ncol = shape[2] * shape[3]
    assert path == (2 if ncol >= 256 else 1)
    one, ref1 = ml.steric(ds, equation_of_state=eos)
    _close_nan(ref["rho"].values, ref1["rho"].values, rtol=1e-15)
    assert float(ref["volo"]) == pytest.approx(float(ref1["volo"]), rel=1e-13)
    assert float(ref["masso"]) == pytest.approx(float(ref1["masso"]), rel=1e-13)
    singles = {"steric": one}
    singles["thermosteric"], _ = ml.thermosteric(ds, equation_of_state=eos, reference=ref1)
    singles["halosteric"], _ = ml.halosteric(ds, equation_of_state=eos, reference=ref1)
    wet = ~torch.isnan(ref["volcello"].data[0])
    for variant in ("steric", "thermosteric", "halosteric"):
        got = res[variant]
        assert got.dims == ("time", "yh", "xh") and got.attrs["long_name"] == f"{variant.capitalize()} height adjustment"
        # the single-variant kernels fold a pinned operand into the coefficients: same polynomial, other association
        _close_nan(got.values, singles[variant][variant].values, atol=1e-11)
        if path == 2:
            assert torch.all(got.data[0][wet] == 0.0)  # the reference step, exactly as in the reference
        else:  # the fallback subtracts a rounded rho_ref: the residue of one rounding, integrated
            assert float(got.data[0][wet].abs().max()) < 1e-12
        ref_o, eta_o, _, _, _, _ = _oracle_case(ds, variant, eos)
        _close_nan(got.values, eta_o, atol=ETA_ATOL)
    # a supplied reference takes the same kernel with rho_ref read instead of evaluated
    again, _ = ml.steric_variants(ds, equation_of_state=eos, reference=ref)
    assert core.last_path() == path
    for variant in ("steric", "thermosteric", "halosteric"):
        _close_nan(again[variant].values, res[variant].values, atol=1e-11)


@pytest.mark.parametrize("nt", [1, 4, 5, 7, 9, 12, 13, 25])
@pytest.mark.parametrize("tc", [0, 4, 6, 8, 12, 104, 106, 108, 112])
def test_one_pass_variants_equal_single_launches_bit_for_bit(ml, nt, tc):
    """csrc/ml_tma3.cu: the one-pass kernel evaluates each height with the single-height kernel's instructions, so
    the three fields, the reference density and volo / masso are bit-identical to separate launches -- for every
    chunk width, with remainder chunks, holes at wet cells (repair pass) and masks that disagree with the bathymetry."""
    from momlevel_b200 import core, synth

    shape = (nt, 20, 8, 160)  # 1280 columns: five tiles
    grid = synth.make_grid(*shape[1:], seed=17, device="cuda")
    pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
    T, S, V = synth.make_fields(grid, nt, seed=17, dtype=torch.float32)
    z_i, depth = grid["z_i"], grid["deptho"]
    if nt > 1:
        T[nt // 2, 2, 3, 7] = float("nan")      # a hole at a wet cell of one step
        S[nt - 1, 1, 5, 100] = float("nan")
    V[3, 2, 40:60] = float("nan")               # volume missing where the bathymetry says water
    T_ref = (T[0] + 0.25).contiguous()          # a supplied reference that is not step 0
    S_ref = (S[0] - 0.05).contiguous()
    rho_given, _ = core.reference_state(T_ref, S_ref, V, pres)

    def run(mode, **kw):
        prev, prev_tc = core.force_direct(mode), core.variants_chunk(tc)
        try:
            out = core.steric_local_variants(T, S, V, z_i, depth, pres, **kw)
            return out, core.last_path()
        finally:
            core.force_direct(prev)
            core.variants_chunk(prev_tc)

    same = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))  # noqa: E731
    for kw in ({}, {"T_ref": T_ref, "S_ref": S_ref, "rho_ref": rho_given}, {"T_ref": T_ref, "S_ref": S_ref}):
        (eta1, rho1, sums1), path1 = run(0, **kw)
        (eta3, rho3, sums3), path3 = run(2, **kw)
        assert path1 == 2 and path3 == 2
        for v in ("steric", "thermosteric", "halosteric"):
            assert same(eta1[v], eta3[v]), (v, kw.keys(), float(torch.nan_to_num(eta1[v] - eta3[v]).abs().max()))
        assert same(rho1, rho3)
        if sums1 is not None:
            assert torch.equal(sums1, sums3)
    # the reference density is optional for a self-reference call; the heights do not depend on it
    prev_tc = core.variants_chunk(tc)
    try:
        eta_n, rho_n, sums_n = core.steric_local_variants(T, S, V, z_i, depth, pres, want_rho_ref=False)
    finally:
        core.variants_chunk(prev_tc)
    (eta1, _, sums1), _ = run(0)
    assert rho_n is None and torch.equal(sums_n, sums1)
    for v in ("steric", "thermosteric", "halosteric"):
        assert same(eta_n[v], eta1[v])
        wet = ~torch.isnan(V[0])
        assert torch.all(eta_n[v][0][wet] == 0.0)  # the reference step: exactly zero, as in the reference


@pytest.mark.parametrize("variant", ["steric", "thermosteric", "halosteric"])
@pytest.mark.parametrize("shape,dtype", [((5, 12, 16, 64), torch.float32), ((3, 7, 9, 13), torch.float64)])
def test_patm_as_a_2d_field(ml, variant, shape, dtype):
    """``patm`` given as a sea-level-pressure field (steric.py:96 broadcasts ``z_l * 1e4 + patm`` by dimension name):
    local and global heights, the reference density and delta_rho against the oracle, for fields that would
    otherwise take the TMA family and for fp64 storage."""
    from momlevel_b200 import core, synth
    from momlevel_b200.labeled import DataArray

    ds = synth.make_dataset(*shape, seed=5, device="cuda", dtype=dtype)
    ny, nx = shape[2:]
    rng = np.random.default_rng(42)
    patm = 101325.0 + rng.normal(0.0, 1500.0, (ny, nx))  # +-15 hPa of weather
    f64 = lambda k: ds[k].values.astype(np.float64)  # noqa: E731
    oref = osteric.reference_state(f64("thetao"), f64("so"), f64("volcello"), f64("areacello"), f64("z_l"), patm=patm)
    eta_o, drho_o = osteric.steric_local(f64("thetao"), f64("so"), f64("z_l"), f64("z_i"), f64("deptho"), oref,
                                         patm=patm, variant=variant)
    g_o, href_o, _ = osteric.steric_global(f64("thetao"), f64("so"), f64("z_l"), oref, patm=patm, variant=variant)
    for arg in (patm, DataArray(patm, ("yh", "xh")), DataArray(patm.T.copy(), ("xh", "yh")), torch.from_numpy(patm).cuda()):
        res, ref = ml.steric(ds, variant=variant, patm=arg)
        assert core.last_path() == 1  # the plain-load family carries the per-column offset
        _close_nan(ref["rho"].values, oref["rho"], rtol=RHO_RTOL)
        _close_nan(res[variant].values, eta_o, atol=ETA_ATOL)
    _close_nan(res["delta_rho"].values, drho_o, atol=1e-9)
    assert float(ref["masso"]) == pytest.approx(float(oref["masso"]), rel=1e-13)
    gres, _ = ml.steric(ds, variant=variant, domain="global", patm=patm, reference=ref)
    assert np.max(np.abs(gres[variant].values - g_o)) < ETA_ATOL
    # the offset is gone after the call: the next one is the plain scalar case again, on the usual family
    res0, _ = ml.steric(ds, variant=variant)
    ref0 = osteric.reference_state(f64("thetao"), f64("so"), f64("volcello"), f64("areacello"), f64("z_l"))
    eta0, _ = osteric.steric_local(f64("thetao"), f64("so"), f64("z_l"), f64("z_i"), f64("deptho"), ref0, variant=variant)
    _close_nan(res0[variant].values, eta0, atol=ETA_ATOL)
    assert float(np.nanmax(np.abs(eta0 - eta_o))) > 1e-7  # the field did change the answer
    with pytest.raises(NotImplementedError):
        ml.steric(ds, patm=np.zeros((shape[0], ny, nx)))
    with pytest.raises(ValueError):
        ml.steric(ds, patm=np.zeros((ny + 1, nx)))


# ------------------------------------------------------- size-independent properties


@pytest.fixture(scope="module")
def big(ml):
    from momlevel_b200 import synth

    # one OM4p25-sized level set would be 116 M points per step; a 1/9 slab keeps the test
    # to seconds while exercising >L2-sized fields and every tile/tail path
    return synth.make_dataset(5, 75, 360, 480, seed=3, device="cuda", dtype=torch.float32)


def test_property_reference_step_is_zero(ml, big):
    """Step 0 is the reference state: eta(t=0) vanishes wherever the column is wet.

    With the reference built from the dataset itself (the fused self-reference pass) the zero
    is exact, as in the reference.  With a *supplied* reference, rho_ref arrives rounded to
    fp64 while the kernel subtracts it from the unrounded product inside one FMA, so eta(0) is
    the rounding residue of rho_ref integrated over the column: ~1e-14 m.
    """
    result, reference = ml.steric(big)
    eta0 = result["steric"].data[0]
    wet = ~torch.isnan(reference["volcello"].data[0])
    assert torch.all(eta0[wet] == 0.0) and torch.all(torch.isnan(eta0[~wet]))
    again, _ = ml.steric(big, reference=reference)
    eta0 = again["steric"].data[0]
    assert float(eta0[wet].abs().max()) < 1e-12 and torch.all(torch.isnan(eta0[~wet]))
    a, b = result["steric"].data, again["steric"].data
    assert float(torch.nan_to_num(a - b).abs().max()) < 1e-12
    g, _ = ml.steric(big, domain="global", reference=reference)
    assert abs(float(g["steric"].values[0])) < 1e-11


@pytest.mark.parametrize("shape", [(14, 9, 16, 64), (3, 10, 37, 53), (5, 75, 8, 96)])
@pytest.mark.parametrize("bcast", ["none", "t", "s"])
@pytest.mark.parametrize("eos", ["Wright", "linear"])
def test_selfref_equals_two_pass(ml, shape, bcast, eos):
    """ml_steric_local_selfref == ml_reference_state followed by ml_steric_local, both families."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(*shape, seed=21, device="cuda", dtype=torch.float32)
    T, S, V = ds["thetao"].data, ds["so"].data, ds["volcello"].data[0]
    pres = ds["z_l"].values * 1e4 + 101325.0
    Tin = T[0].contiguous() if bcast == "t" else T
    Sin = S[0].contiguous() if bcast == "s" else S
    kw = dict(eos=eos, t_bcast=bcast == "t", s_bcast=bcast == "s")
    rho2, sums2 = core.reference_state(T[0], S[0], V, pres, eos=eos)
    eta2, _ = core.steric_local(Tin, Sin, rho2, V, ds["z_i"].data, ds["deptho"].data, pres, **kw)
    for direct in (False, True):
        prev = core.force_direct(direct)
        try:
            eta1, rho1, sums1 = core.steric_local_selfref(Tin, Sin, V, ds["z_i"].data, ds["deptho"].data, pres, **kw)
        finally:
            core.force_direct(prev)
        assert torch.equal(torch.isnan(rho1), torch.isnan(rho2))
        assert torch.equal(torch.nan_to_num(rho1), torch.nan_to_num(rho2))  # bit-identical reference density
        assert torch.allclose(sums1, sums2, rtol=1e-13, atol=0)
        assert torch.equal(torch.isnan(eta1), torch.isnan(eta2))
        assert float(torch.nan_to_num(eta1 - eta2).abs().max()) < 1e-12


def test_property_linear_eos_is_additive(ml, big):
    """Under the linear EOS thermosteric + halosteric == steric (to rounding)."""
    s, ref = ml.steric(big, equation_of_state="linear")
    t, _ = ml.thermosteric(big, equation_of_state="linear", reference=ref)
    h, _ = ml.halosteric(big, equation_of_state="linear", reference=ref)
    a, b = s["steric"].data, t["thermosteric"].data + h["halosteric"].data
    m = ~torch.isnan(a)
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    assert float((a[m] - b[m]).abs().max()) < 1e-9


def test_property_kernel_families_agree_and_are_deterministic(ml, big):
    r0, ref = ml.steric(big)
    r0b, _ = ml.steric(big)
    assert torch.equal(torch.nan_to_num(r0["steric"].data), torch.nan_to_num(r0b["steric"].data))
    r1, _ = ml.steric(big, reference=ref)
    r2, _ = ml.steric(big, reference=ref)
    assert torch.equal(torch.nan_to_num(r1["steric"].data), torch.nan_to_num(r2["steric"].data))
    g1, _ = ml.steric(big, domain="global", reference=ref)
    g2, _ = ml.steric(big, domain="global", reference=ref)
    assert np.array_equal(g1["steric"].values, g2["steric"].values)
    prev = ml.core.force_direct(True)
    try:
        r3, _ = ml.steric(big, reference=ref)
        g3, _ = ml.steric(big, domain="global", reference=ref)
    finally:
        ml.core.force_direct(prev)
    a, b = r1["steric"].data, r3["steric"].data
    m = ~torch.isnan(a)
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    assert float((a[m] - b[m]).abs().max()) < 1e-12
    assert np.allclose(g1["steric"].values, g3["steric"].values, rtol=0, atol=1e-12)


def test_property_eta_equals_integral_of_delta_rho(ml, big):
    """eta == -1/rho0 * sum_z dz * delta_rho, with both fields produced by the library."""
    result, reference = ml.steric(big)
    dz = ml.derived.calc_dz(big["z_l"], big["z_i"], big["deptho"]).transpose("z_l", "yh", "xh").data
    drho = result["delta_rho"].data
    eta = (-1.0 / 1035.0) * torch.nansum(dz.unsqueeze(0) * drho, dim=1)
    a = result["steric"].data
    m = ~torch.isnan(a)
    assert float((a[m] - eta[m]).abs().max()) < 1e-11


def test_subsample_parity_at_size(ml, big):
    """Oracle on a slab of the big case (what bench.py's cpu_baseline leg also does)."""
    result, reference = ml.steric(big)
    ys = slice(100, 112)
    sub = {k: big[k].values for k in ("z_l", "z_i")}
    f64 = lambda k: big[k].data[..., ys, :].cpu().numpy().astype(np.float64)  # noqa: E731
    ref = osteric.reference_state(f64("thetao"), f64("so"), f64("volcello"), big["areacello"].values[ys],
                                  sub["z_l"])
    eta, _ = osteric.steric_local(f64("thetao"), f64("so"), sub["z_l"], sub["z_i"], big["deptho"].values[ys], ref)
    _close_nan(result["steric"].data[:, ys, :].cpu().numpy(), eta, atol=ETA_ATOL)


def _sampled_columns(ncol, n_stride=96, tile=256):
    """Columns from every part of the grid: both ends of the first, second, middle and last 256-column tile of the
    TMA family and a stride across everything in between (not one slab of rows from the middle)."""
    tiles = (ncol + tile - 1) // tile
    picks = {0, 1, tile - 1, tile, tile + 1, 2 * tile - 1, (tiles // 2) * tile - 1, (tiles // 2) * tile,
             (tiles - 1) * tile - 1, (tiles - 1) * tile, (tiles - 1) * tile + 1, ncol - 2, ncol - 1}
    picks |= {int(x) for x in np.linspace(0, ncol - 1, n_stride)}
    return np.array(sorted(p for p in picks if 0 <= p < ncol), dtype=np.int64)


def _oracle_on_columns(T, S, V, grid, cols, eos="Wright", variants=("steric",), T_ref=None, S_ref=None):
    """The oracle (fp64-upcast inputs) on the sampled columns of device-resident fields ``[t][z][y][x]``.
    Returns ``(reference dict, {variant: eta[t][col]}, masso[t] of these columns)``."""
    idx = torch.as_tensor(cols, device=T.device)
    f64 = lambda x: x.double().cpu().numpy()  # noqa: E731
    Tc = f64(T.flatten(2)[:, :, idx])[:, :, None, :]
    Sc = f64(S.flatten(2)[:, :, idx])[:, :, None, :]
    Vc = f64(V.flatten(1)[:, idx])[None, :, None, :]
    depth = f64(grid["deptho"].flatten()[idx])[None, :]
    area = f64(grid["areacello"].flatten()[idx])[None, :]
    z_l, z_i = f64(grid["z_l"]), f64(grid["z_i"])
    if T_ref is None:
        oref = osteric.reference_state(Tc, Sc, Vc, area, z_l, eos=eos)
    else:
        T0 = f64(T_ref.flatten(1)[:, idx])[None, :, None, :]
        S0 = f64(S_ref.flatten(1)[:, idx])[None, :, None, :]
        oref = osteric.reference_state(T0, S0, Vc, area, z_l, eos=eos)
    etas = {v: osteric.steric_local(Tc, Sc, z_l, z_i, depth, oref, eos=eos, variant=v)[0][:, 0, :] for v in variants}
    _, _, masso = osteric.steric_global(Tc, Sc, z_l, oref, eos=eos)
    return oref, etas, masso


def test_full_size_om4p25_properties(ml):
    """BASELINE config 2 at full size (1440x1080x75 x 12): size-independent properties + a slab vs the oracle."""
    from momlevel_b200 import synth

    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~40 GB of free HBM")
    ds = synth.make_dataset(12, 75, 1080, 1440, seed=123, device="cuda", dtype=torch.float32)
    result, reference = ml.steric(ds)
    assert ml.core.last_path() == 2  # the TMA family carried it
    eta = result["steric"].data
    wet = ~torch.isnan(reference["volcello"].data[0])
    assert torch.all(eta[0][wet] == 0.0)                       # reference step
    assert torch.equal(torch.isnan(eta[5]), ~wet)              # mask = surface volcello
    assert float(eta[:, wet].abs().max()) < 5.0                # metres, sane
    # thermosteric + halosteric ~ steric to first order; exact additivity under the linear EOS
    lin, lref = ml.steric(ds, equation_of_state="linear")
    th, _ = ml.thermosteric(ds, equation_of_state="linear", reference=lref)
    ha, _ = ml.halosteric(ds, equation_of_state="linear", reference=lref)
    d = lin["steric"].data - th["thermosteric"].data - ha["halosteric"].data
    assert float(torch.nan_to_num(d).abs().max()) < 1e-9
    # checksum of checksums: the global mass series from the 4-D field equals the kernel's
    g, _ = ml.steric(ds, domain="global", reference=reference)
    assert abs(float(g["steric"].values[0])) < 1e-10 and np.all(np.isfinite(g["steric"].values))
    # columns from every part of the grid against the oracle: the three heights (BASELINE config 2 names them
    # together; the one-pass kernel serves the call) and the reference density
    grid = {k: ds[k].data for k in ("deptho", "areacello", "z_l", "z_i")}
    T, S, V = ds["thetao"].data, ds["so"].data, ds["volcello"].data[0]
    cols = _sampled_columns(1080 * 1440)
    oref, oetas, _ = _oracle_on_columns(T, S, V, grid, cols, variants=("steric", "thermosteric", "halosteric"))
    idx = torch.as_tensor(cols, device="cuda")
    _close_nan(eta.flatten(1)[:, idx].cpu().numpy(), oetas["steric"], atol=ETA_ATOL)
    _close_nan(reference["rho"].data.flatten(1)[:, idx].cpu().numpy(), oref["rho"][:, 0, :], rtol=RHO_RTOL)
    all3, _ = ml.steric_variants(ds)
    for v in ("steric", "thermosteric", "halosteric"):
        _close_nan(all3[v].data.flatten(1)[:, idx].cpu().numpy(), oetas[v], atol=ETA_ATOL)
    assert torch.equal(torch.nan_to_num(all3["steric"].data), torch.nan_to_num(eta))
    # BASELINE config 5 on the same fields: the linear EOS height and Flament spiciness against the oracle
    lidx = idx
    _, letas, _ = _oracle_on_columns(T, S, V, grid, cols, eos="linear")
    _close_nan(lin["steric"].data.flatten(1)[:, lidx].cpu().numpy(), letas["steric"], atol=ETA_ATOL)
    half = ml.core.flament_spice(T[:6], S[:6])
    pts = torch.arange(0, half.numel(), 997, device="cuda")
    Tp, Sp = (x[:6].flatten()[pts].double().cpu().numpy() for x in (T, S))
    m = ~(np.isnan(Tp) | np.isnan(Sp))
    from oracle import spice as ospice

    got = half.flatten()[pts].cpu().numpy()
    assert np.all(np.isnan(got[~m]))
    want = ospice.flament_spice(Tp[m], Sp[m])
    assert np.max(np.abs(got[m] - want) / np.maximum(np.abs(want), 1.0)) < 1e-12
    del half


def test_full_size_spear_member(ml):
    """BASELINE config 3, one ensemble member at full size (360x320x75, 120 months): ten register chunks,
    a slab against the oracle, step 0 exactly zero, and the member loop on streams against plain calls."""
    from momlevel_b200 import core, synth
    from momlevel_b200 import distributed as mld

    nt, nz, ny, nx = synth.CONFIGS["spear1deg"]
    grid = synth.make_grid(nz, ny, nx, seed=7, device="cuda")
    T, S, V = synth.make_fields(grid, nt, seed=1000, dtype=torch.float32)
    pres = grid["z_l"] * 1.0e4 + 101325.0
    eta, rho, sums = core.steric_local_selfref(T, S, V, grid["z_i"], grid["deptho"], pres)
    assert core.last_path() == 2
    wet = ~torch.isnan(V[0])
    assert torch.all(eta[0][wet] == 0.0) and torch.equal(torch.isnan(eta[77]), ~wet)
    cols = _sampled_columns(ny * nx)
    idx = torch.as_tensor(cols, device="cuda")
    oref, oetas, _ = _oracle_on_columns(T, S, V, grid, cols)
    _close_nan(eta.flatten(1)[:, idx].cpu().numpy(), oetas["steric"], atol=ETA_ATOL)
    _close_nan(rho.flatten(1)[:, idx].cpu().numpy(), oref["rho"][:, 0, :], rtol=RHO_RTOL)
    # a block of the member that does not start at its step 0 (distributed.assign_member_blocks): reference from
    # the step-0 slabs, then the block against it
    (e3, r3, s3), = mld.steric_local_pieces([(T[36:].contiguous(), S[36:].contiguous(), V, (T[0].contiguous(), S[0].contiguous()))],
                                            grid["z_i"], grid["deptho"], pres)
    assert torch.equal(torch.nan_to_num(r3), torch.nan_to_num(rho))
    assert float(torch.nan_to_num(e3 - eta[36:]).abs().max()) < 1e-12
    (e2, r2, s2), = mld.steric_local_members([(T, S, V)], grid["z_i"], grid["deptho"], pres)
    assert torch.equal(torch.nan_to_num(e2), torch.nan_to_num(eta)) and torch.equal(s2, sums)


def test_full_size_om4p125_window(ml):
    """BASELINE config 4 at full grid size (2880x2240x75), a 10-step window of the daily series: the global
    mass series of the fused kernel against the 4-D route (ml_eos_eval + a torch reduction) and, on a slab
    of rows, against the oracle's masses."""
    from momlevel_b200 import core, synth

    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~70 GB of free HBM")
    nz, ny, nx = synth.CONFIGS["om4p125"][1:]
    nt = 10  # 12 + remainder logic: one 12-wide chunk with two idle rows
    grid = synth.make_grid(nz, ny, nx, seed=11, device="cuda")
    T, S, V = synth.make_fields(grid, nt, seed=55, dtype=torch.float32)
    pres = grid["z_l"] * 1.0e4 + 101325.0
    masso = core.steric_global(T, S, V, pres)
    assert core.last_path() == 2 and masso.shape == (nt,)
    # the reference's own route for one step: rho as a field, times volcello, skipna sum (derived.py:435-438)
    for t in (0, nt - 1):
        rho = core.eos_eval("Wright", "density", T[t], S[t], pres, z_axis=0)
        want = torch.nansum(rho * V.double())
        assert float((masso[t] - want).abs() / want) < 1e-12
        del rho
    # the masses of a sub-grid made of columns from every part of the grid, against the oracle
    cols = _sampled_columns(ny * nx, n_stride=1024)
    cols = cols[: len(cols) // 4 * 4]  # rows of whole 16-byte units: the sub-grid takes the TMA family too
    idx = torch.as_tensor(cols, device="cuda")
    _, _, omass = _oracle_on_columns(T, S, V, grid, cols)
    got = core.steric_global(T.flatten(2)[:, :, idx].contiguous(), S.flatten(2)[:, :, idx].contiguous(),
                             V.flatten(1)[:, idx].contiguous(), pres)
    assert np.allclose(got.cpu().numpy(), omass, rtol=1e-12, atol=0)


@pytest.mark.parametrize("shape", [(3, 10, 37, 53), (5, 9, 16, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_delta_rho_entry_equals_fused_output(ml, shape, dtype):
    """ml_delta_rho (vectorised and scalar paths) == the delta_rho the direct column kernel writes."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(*shape, seed=4, device="cuda", dtype=dtype)
    T, S, V = ds["thetao"].data, ds["so"].data, ds["volcello"].data[0]
    pres = ds["z_l"].values * 1e4 + 101325.0
    rho_ref, _ = core.reference_state(T[0], S[0], V, pres)
    for kw in ({}, {"s_bcast": True}, {"t_bcast": True}):
        Tin = T[0].contiguous() if kw.get("t_bcast") else T
        Sin = S[0].contiguous() if kw.get("s_bcast") else S
        a = core.delta_rho(Tin, Sin, rho_ref, V, pres, **kw)
        _, b = core.steric_local(Tin, Sin, rho_ref, V, ds["z_i"].data, ds["deptho"].data, pres, want_delta_rho=True, **kw)
        assert torch.equal(torch.isnan(a), torch.isnan(b))
        assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))


# -------------------------------------------------------------------- host (e2e) entry


def test_host_entry_matches_device_path(ml):
    from momlevel_b200 import synth

    ds = synth.make_dataset(5, 12, 20, 32, seed=5, device="cpu", dtype=torch.float32)
    pres = ds["z_l"].values * 1e4 + 101325.0
    for spw in (1, 2, 5, 8):
        eta, rho, (volo, masso) = ml.core.steric_local_host(
            ds["thetao"].data, ds["so"].data, ds["volcello"].data[0], ds["z_i"].values, ds["deptho"].values, pres,
            steps_per_window=spw, want_rho_ref=True)
        result, reference = ml.steric(ds)
        _close_nan(eta.numpy(), result["steric"].values, atol=1e-13)
        _close_nan(rho.numpy(), reference["rho"].values, rtol=1e-15)
        assert volo == pytest.approx(float(reference["volo"]), rel=1e-14)
        assert masso == pytest.approx(float(reference["masso"]), rel=1e-14)


def test_global_host_entry_matches_device_path(ml):
    """ml_steric_global_host: the masses of a host-resident series, any window width, against the device call."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(7, 12, 20, 32, seed=6, device="cpu", dtype=torch.float32)
    pres = ds["z_l"].values * 1e4 + 101325.0
    T, S, V = ds["thetao"].data, ds["so"].data, ds["volcello"].data[0].contiguous()
    want = core.steric_global(T.cuda(), S.cuda(), V.cuda(), pres).cpu()
    for spw in (1, 3, 7, 12):
        got = core.steric_global_host(T, S, V, pres, steps_per_window=spw)
        assert got.shape == (7,) and torch.allclose(got, want, rtol=1e-14, atol=0)
    # the public route: steric(domain="global") on the same data
    res, ref = ml.steric(ds, domain="global")
    eta, href = ml.distributed.global_sea_level(got.numpy(), float(ref["volo"]), float(ref["rhoga"]),
                                                float(ref["areacello"].sum()))
    assert np.allclose(eta, res["steric"].values, rtol=0, atol=1e-12)


def test_host_entry_all_variants_from_one_transfer(ml):
    """ml_steric_local_variants_host: three heights from one pass of the fields over PCIe == three device calls."""
    from momlevel_b200 import synth

    ds = synth.make_dataset(7, 12, 20, 32, seed=5, device="cpu", dtype=torch.float32)
    pres = ds["z_l"].values * 1e4 + 101325.0
    want = {}
    want["steric"], ref = ml.steric(ds)
    want["thermosteric"], _ = ml.thermosteric(ds)
    want["halosteric"], _ = ml.halosteric(ds)
    for spw in (1, 3, 7):
        etas, rho, (volo, masso) = ml.core.steric_local_host(
            ds["thetao"].data, ds["so"].data, ds["volcello"].data[0], ds["z_i"].values, ds["deptho"].values, pres,
            steps_per_window=spw, want_rho_ref=True, variants=True)
        for variant in ("steric", "thermosteric", "halosteric"):
            _close_nan(etas[variant].numpy(), want[variant][variant].values, atol=1e-12)
            wet = ~np.isnan(ref["volcello"].values[0])
            assert np.all(etas[variant].numpy()[0][wet] == 0.0)  # step 0 is the reference state, for every variant
        _close_nan(rho.numpy(), ref["rho"].values, rtol=1e-15)
        assert masso == pytest.approx(float(ref["masso"]), rel=1e-14)


def _disagreeing_masks(ds, seed):
    """Host fields whose volcello mask, bathymetry and data holes disagree (as in test_nan_semantics_tma_family)."""
    T, S = ds["thetao"].data.clone(), ds["so"].data.clone()
    V = ds["volcello"].data[0].clone()
    depth = ds["deptho"].data.clone()
    shape = tuple(T.shape)
    g = torch.Generator().manual_seed(seed)
    for k in range(40):
        t, z, y, x = (int(torch.randint(0, n, (1,), generator=g)) for n in shape)
        (T if k % 2 else S)[t, z, y, x] = float("nan")
    for _ in range(30):
        z, y, x = (int(torch.randint(0, n, (1,), generator=g)) for n in shape[1:])
        V[z, y, x] = float("nan") if torch.isfinite(V[z, y, x]) else 1.0e9
    for _ in range(12):
        y, x = (int(torch.randint(0, n, (1,), generator=g)) for n in shape[2:])
        depth[y, x] = float("nan") if torch.isfinite(depth[y, x]) else 3000.0
    # data where the reference volume is missing: the reference never reads it, the packed rows never carry it
    junk = torch.isnan(V).unsqueeze(0) & (torch.rand(shape, generator=g) < 0.3)
    T[junk] = 25.0
    S[junk] = 31.0
    return T, S, V, depth


@pytest.mark.parametrize("shape", [(5, 12, 20, 32), (4, 7, 5, 7), (3, 8, 256, 520)])
def test_host_packed_transfer_is_bit_identical(ml, shape):
    """Level rows that cross PCIe as their present cells give the heights of rows that cross as they are,
    bit for bit, and both agree with the oracle -- whatever the masks, the window width or the split."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(*shape, seed=11, device="cpu", dtype=torch.float32)
    T, S, V, depth = _disagreeing_masks(ds, 3)
    z_l, z_i = ds["z_l"].values, ds["z_i"].values
    pres = z_l * 1e4 + 101325.0
    T64, S64, V64 = T.double().numpy(), S.double().numpy(), V.double().numpy()
    oref = osteric.reference_state(T64, S64, np.broadcast_to(V64, T64.shape), ds["areacello"].values, z_l)
    want, _ = osteric.steric_local(T64, S64, z_l, z_i, depth.numpy(), oref, variant="steric")
    dense_bytes = 2 * T.numel() * 4
    try:
        core.host_packing(0)
        eta0, _, sums0 = core.steric_local_host(T, S, V, z_i, depth.numpy(), pres, steps_per_window=2)
        bytes0, frac0 = core.host_last_transfer()
        assert frac0 == 0.0 and bytes0 >= dense_bytes
        _close_nan(eta0.numpy(), want, atol=ETA_ATOL)
        assert sums0[1] == pytest.approx(oref["masso"], rel=1e-12) and sums0[0] == pytest.approx(oref["volo"], rel=1e-12)
        for mode, threads, spw in ((2, 0, 1), (2, 3, 2), (2, 1, shape[0]), (1, 0, 1), (1, 2, 3), (3, 0, 1), (3, 3, 2)):
            core.host_packing(mode, threads)
            eta, rho, sums = core.steric_local_host(T, S, V, z_i, depth.numpy(), pres, steps_per_window=spw)
            nbytes, frac = core.host_last_transfer()
            assert torch.equal(eta.view(torch.int64), eta0.view(torch.int64)), (mode, threads, spw)
            assert sums == sums0
            if mode == 2:
                assert frac > 0.0 and nbytes < bytes0
        # a call that wants rho_ref back moves every row as it is (rho_ref is defined on absent cells too)
        core.host_packing(2)
        eta, rho, _ = core.steric_local_host(T, S, V, z_i, depth.numpy(), pres, want_rho_ref=True)
        assert core.host_last_transfer()[1] == 0.0
        assert torch.equal(eta.view(torch.int64), eta0.view(torch.int64))
        _close_nan(rho.numpy(), oref["rho"], rtol=RHO_RTOL)
        # all three heights from one packed transfer
        core.host_packing(0)
        etas0, _, _ = core.steric_local_host(T, S, V, z_i, depth.numpy(), pres, variants=True)
        core.host_packing(2)
        etas2, _, _ = core.steric_local_host(T, S, V, z_i, depth.numpy(), pres, variants=True)
        assert core.host_last_transfer()[1] > 0.0
        for variant in ("steric", "thermosteric", "halosteric"):
            assert torch.equal(etas2[variant].view(torch.int64), etas0[variant].view(torch.int64)), variant
            o, _ = osteric.steric_local(T64, S64, z_l, z_i, depth.numpy(), oref, variant=variant)
            _close_nan(etas2[variant].numpy(), o, atol=ETA_ATOL)
        # the global masses through the same transfer
        core.host_packing(0)
        m0 = core.steric_global_host(T, S, V, pres, steps_per_window=2)
        for mode in (2, 3, 1):
            core.host_packing(mode)
            m = core.steric_global_host(T, S, V, pres, steps_per_window=1)
            assert torch.allclose(m, m0, rtol=1e-14, atol=0)
        _, _, omass = osteric.steric_global(T64, S64, z_l, oref, variant="steric")
        assert np.allclose(m.numpy(), omass, rtol=1e-12, atol=0)
    finally:
        core.host_packing(1)


def test_host_packed_transfer_of_a_dense_field(ml):
    """No absent cells (the reference's own 5x5x5 test dataset): from pinned memory nothing is worth compressing
    and every row is copied as it is; from pageable memory every row goes through the pinned staging."""
    from momlevel_b200 import core

    ds = ml.test_data.generate_test_data()
    T = torch.from_numpy(ds["thetao"].values).float()
    S = torch.from_numpy(ds["so"].values).float()
    V = torch.from_numpy(ds["volcello"].values[0]).float()
    pres = ds["z_l"].values * 1e4 + 101325.0
    args = (ds["z_i"].values, ds["deptho"].values, pres)
    try:
        core.host_packing(0)
        eta0, _, _ = core.steric_local_host(T, S, V, *args)
        for mode in (1, 2):
            core.host_packing(mode)
            eta, _, _ = core.steric_local_host(T.pin_memory(), S.pin_memory(), V.pin_memory(), *args)
            assert core.host_last_transfer()[1] == 0.0
            assert torch.equal(eta.view(torch.int64), eta0.view(torch.int64))
            eta, _, _ = core.steric_local_host(T, S, V, *args)  # pageable
            assert core.host_last_transfer()[1] == 1.0
            assert torch.equal(eta.view(torch.int64), eta0.view(torch.int64))
    finally:
        core.host_packing(1)


def test_host_packed_transfer_from_pinned_memory(ml):
    """Pinned fields: the fullest rows are copied straight from the caller's buffer, the emptiest ones are packed,
    and the split (which depends on timing) does not change a bit of the result."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(4, 10, 128, 520, seed=12, device="cpu", dtype=torch.float32)
    T, S, V, depth = _disagreeing_masks(ds, 5)
    V[:3] = 1.0e9  # three levels with every volume present: rows that are never worth packing
    pres = ds["z_l"].values * 1e4 + 101325.0
    args = (ds["z_i"].values, depth.numpy(), pres)
    Tp, Sp, Vp = T.pin_memory(), S.pin_memory(), V.pin_memory()
    try:
        core.host_packing(0)
        eta0, _, sums0 = core.steric_local_host(Tp, Sp, Vp, *args)
        seen = set()
        for mode, threads in ((1, 0), (1, 1), (2, 0), (1, 4), (2, 2), (3, 0), (3, 5)):
            core.host_packing(mode, threads)
            for spw in (1, 3):
                eta, _, sums = core.steric_local_host(Tp, Sp, Vp, *args, steps_per_window=spw)
                frac = core.host_last_transfer()[1]
                seen.add(round(frac, 3))
                assert 0.0 <= frac <= 0.7  # three of the ten levels are fully present: never packed
                assert torch.equal(eta.view(torch.int64), eta0.view(torch.int64)) and sums == sums0
                if mode == 2:
                    assert frac == pytest.approx(0.7)
        assert max(seen) > 0.0
    finally:
        core.host_packing(1)


def test_steric_takes_the_host_route_for_host_resident_fields(ml, monkeypatch):
    """``steric(dset)`` on numpy-backed fields (what xarray hands over) streams them through the host entry
    point; the heights agree with the device-resident call and the oracle for every variant."""
    import importlib

    from momlevel_b200 import core, synth

    steric_mod = importlib.import_module("momlevel_b200.steric")  # the package attribute is the function
    shape = (6, 20, 256, 520)  # 128 MB of T and S: above HOST_ROUTE_MIN_BYTES
    ds = synth.make_dataset(*shape, seed=21, device="cpu", dtype=torch.float32)
    dims = ("time", "z_l", "yh", "xh")
    host = ml.Dataset()
    dev = ml.Dataset()
    for k in ds.variables:
        host[k] = ds[k]
        dev[k] = ds[k]
    for k in ("thetao", "so"):
        host[k] = ml.DataArray(ds[k].data.numpy(), dims)  # plain numpy: pageable memory
        dev[k] = ml.DataArray(ds[k].data.cuda(), dims)
    host["volcello"] = ml.DataArray(np.ascontiguousarray(ds["volcello"].values), dims)
    dev["volcello"] = ml.DataArray(ds["volcello"].data.cuda(), dims)
    for k in ("deptho", "areacello", "z_l", "z_i"):
        dev[k] = ml.DataArray(torch.as_tensor(ds[k].values).cuda(), ds[k].dims)
    assert steric_mod._host_resident(host, "time", "z_l", "z_i") and not steric_mod._host_resident(dev, "time", "z_l", "z_i")
    T64, S64 = ds["thetao"].values.astype(np.float64), ds["so"].values.astype(np.float64)
    V64 = ds["volcello"].values.astype(np.float64)
    z_l, z_i = ds["z_l"].values, ds["z_i"].values
    oref = osteric.reference_state(T64, S64, V64, ds["areacello"].values, z_l)
    for variant in ("steric", "thermosteric", "halosteric"):
        core.host_packing(1)
        got, ref = ml.steric(host, variant=variant)
        nbytes, frac = core.host_last_transfer()
        assert frac == 1.0 and nbytes < 2 * T64.size * 4, "expected the packed host route"
        assert not got[variant].data.is_cuda
        want, wref = ml.steric(dev, variant=variant)
        _close_nan(got[variant].values, want[variant].values, atol=1e-12)
        assert float(ref["masso"]) == pytest.approx(float(wref["masso"]), rel=1e-14)
        assert float(ref["volo"]) == pytest.approx(float(wref["volo"]), rel=1e-14)
        o, _ = osteric.steric_local(T64, S64, z_l, z_i, ds["deptho"].values, oref, variant=variant)
        _close_nan(got[variant].values, o, atol=ETA_ATOL)
        if variant == "steric":  # the lazy members still work from host fields
            _close_nan(ref["rho"].values, oref["rho"], rtol=RHO_RTOL)
            assert got["delta_rho"].shape == shape
        # the global domain from host memory, every variant (ml_host_stream_*: the pinned operand's reference slab
        # stays on the device): against the device-resident call and the oracle
        before = core.launch_count()
        gh, _ = ml.steric(host, variant=variant, domain="global", reference=ref)
        assert core.host_last_transfer()[0] > 0 and core.launch_count() > before
        gd, _ = ml.steric(dev, variant=variant, domain="global", reference=wref)
        assert np.max(np.abs(gh[variant].values - gd[variant].values)) < 1e-12  # (the two references' scalars differ by an ulp)
        g_o, _, _ = osteric.steric_global(T64, S64, z_l, oref, variant=variant)
        assert np.max(np.abs(gh[variant].values - g_o)) < ETA_ATOL
    # small datasets are simply copied to the device; forcing the route shows the same numbers
    small = synth.make_dataset(5, 12, 20, 32, seed=5, device="cpu", dtype=torch.float32)
    assert not steric_mod._host_resident(small, "time", "z_l", "z_i")
    want, _ = ml.steric(small)
    monkeypatch.setattr(steric_mod, "HOST_ROUTE_MIN_BYTES", 0)
    assert steric_mod._host_resident(small, "time", "z_l", "z_i")
    got, _ = ml.steric(small)
    _close_nan(got["steric"].values, want["steric"].values, atol=1e-13)
    res, _ = ml.steric(small, annual=False, domain="global")  # the global branch is untouched by the route
    assert res["steric"].shape == (5,)


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_random_shapes_tma_against_direct(ml, seed, dtype):
    """Shapes drawn at random around the TMA family's edges -- a last tile with a few columns, one or two
    levels, a single step, more levels than the ring is deep -- must agree with the direct family, for fields
    stored as fp32 (12-step chunks) and as fp64 (6-step chunks, rows a multiple of two values)."""
    from momlevel_b200 import core, synth

    rng = np.random.default_rng(1000 + seed)
    nt = int(rng.choice([1, 2, 3, 5, 11, 12, 13, 17, 29]))
    nz = int(rng.choice([1, 2, 3, 4, 5, 9, 33, 75, 130]))
    ny = int(rng.integers(1, 9))
    per16 = 4 if dtype == torch.float32 else 2
    nx = per16 * int(rng.integers(256 // (per16 * ny) + 1, 800 // per16))  # rows of whole 16-byte units, ncol >= 256
    if seed % 3 == 2:
        nx += 1 + seed % 2  # rows that are NOT whole 16-byte units: rank-1 tensor maps, one 1-D box per row
    grid = synth.make_grid(nz, ny, nx, seed=seed, device="cuda")
    T, S, V = synth.make_fields(grid, nt, seed=seed, dtype=dtype)
    if nt > 1 and nz > 1:
        T[nt - 1, nz // 2, 0, nx // 2] = float("nan")
    pres = grid["z_l"] * 1.0e4 + 101325.0
    out = {}
    for direct in (False, True):
        prev = core.force_direct(direct)
        try:
            eta, rho, sums = core.steric_local_selfref(T, S, V, grid["z_i"], grid["deptho"], pres)
            assert core.last_path() == (1 if direct else 2), (nt, nz, ny, nx)
            eta_t, _ = core.steric_local(T, S[0], rho, V, grid["z_i"], grid["deptho"], pres, s_bcast=True)
            eta_h, _ = core.steric_local(T[0], S, rho, V, grid["z_i"], grid["deptho"], pres, t_bcast=True)
            masso = core.steric_global(T, S, V, pres)
            out[direct] = (eta, rho, sums, eta_t, eta_h, masso)
        finally:
            core.force_direct(prev)
    for a, b in zip(out[False], out[True]):
        assert a.shape == b.shape and torch.equal(torch.isnan(a), torch.isnan(b)), (nt, nz, ny, nx)
        err = float(torch.nan_to_num(a - b).abs().max())
        scale = float(torch.nan_to_num(b).abs().max())
        assert err <= 1e-11 + 1e-13 * scale, (nt, nz, ny, nx, err)


def test_deferred_value_checks_raise_what_the_eager_ones_do(ml):
    """Device-resident datasets take the fused path whose value checks are read back after the launch:
    same exceptions, same warning (util.py:783-792, derived.py:284-292)."""
    from momlevel_b200 import synth

    def fresh():
        return synth.make_dataset(3, 6, 16, 64, seed=12, device="cuda", dtype=torch.float32)

    ds = fresh()
    ds["areacello"] = ml.DataArray(ds["areacello"].data * 2.0, ("yh", "xh"))
    with pytest.raises(ValueError, match="Errors found in dataset."):
        ml.steric(ds)
    with pytest.warns(UserWarning, match="areacello"):
        res, _ = ml.steric(ds, strict=False)
    assert res["steric"].shape == (3, 16, 64)
    ds = fresh()
    depth = ds["deptho"].data.clone()
    depth[3, 5] = -10.0
    ds["deptho"] = ml.DataArray(depth, ("yh", "xh"))
    with pytest.raises(AssertionError, match="Depth values"):
        ml.steric(ds)
    ds = fresh()
    z_i = ds["z_i"].data.clone()
    z_i[0] = -1.0
    ds["z_i"] = ml.DataArray(z_i, ("z_i",))
    with pytest.raises(AssertionError, match="interfaces"):
        ml.thermosteric(ds)


def test_reference_density_is_produced_on_demand(ml):
    """The fused default call does not store rho_ref (ml_steric_local_selfref with rho_ref = NULL); the
    reference Dataset evaluates it on first access, bit-identical to the stored one, and longer series or
    layouts that need the field internally still get it."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(12, 10, 16, 64, seed=14, device="cuda", dtype=torch.float32)
    T, S, V = ds["thetao"].data, ds["so"].data, ds["volcello"].data[0]
    pres = ds["z_l"].data * 1.0e4 + 101325.0
    eta1, rho1, sums1 = core.steric_local_selfref(T, S, V, ds["z_i"].data, ds["deptho"].data, pres)
    eta0, rho0, sums0 = core.steric_local_selfref(T, S, V, ds["z_i"].data, ds["deptho"].data, pres, want_rho_ref=False)
    assert rho0 is None and rho1 is not None and core.last_path() == 2
    assert torch.equal(torch.nan_to_num(eta0), torch.nan_to_num(eta1)) and torch.equal(sums0, sums1)
    result, reference = ml.steric(ds)
    assert reference["rho"].is_lazy and reference["rho"].dims == ("z_l", "yh", "xh")
    assert torch.equal(torch.nan_to_num(reference["rho"].data), torch.nan_to_num(rho1))  # evaluated here
    again, _ = ml.thermosteric(ds, reference=reference)  # and usable as a supplied reference
    want, _ = ml.thermosteric(ds)
    _close_nan(again["thermosteric"].values, want["thermosteric"].values, atol=1e-12)
    # 13 steps: the later chunk reads rho_ref, so it is produced whether asked for or not
    ds13 = synth.make_dataset(13, 10, 16, 64, seed=14, device="cuda", dtype=torch.float32)
    out = core.steric_local_selfref(ds13["thetao"].data, ds13["so"].data, ds13["volcello"].data[0], ds13["z_i"].data,
                                    ds13["deptho"].data, pres, want_rho_ref=False)
    assert out[1] is not None
    # a grid with fewer columns than a tile takes the direct family, which needs the field as well
    dsr = synth.make_dataset(3, 10, 7, 31, seed=14, device="cuda", dtype=torch.float32)
    out = core.steric_local_selfref(dsr["thetao"].data, dsr["so"].data, dsr["volcello"].data[0], dsr["z_i"].data,
                                    dsr["deptho"].data, pres, want_rho_ref=False)
    assert out[1] is not None and core.last_path() == 1


# ------------------------------------------------- fields that arrive block by block (core.HostStream, ChunkedArray)


def _chunked_dataset(ds, chunks):
    """The same labelled Dataset with thetao / so / volcello held as blocks along time (what dask-backed variables are)."""
    from momlevel_b200.labeled import ChunkedArray, DataArray

    out = ds.copy()
    made = {}
    for name in ("thetao", "so", "volcello"):
        full = ds[name].values
        cuts = np.cumsum((0,) + tuple(chunks))

        def blocks(full=full, cuts=cuts):
            for a, b in zip(cuts[:-1], cuts[1:]):
                yield np.array(full[a:b])

        made[name] = ChunkedArray(full.shape, full.dtype, chunks, blocks)
        out[name] = DataArray(made[name], ds[name].dims, attrs=ds[name].attrs)
    return out, made


@pytest.mark.parametrize("variant", ["steric", "thermosteric", "halosteric"])
@pytest.mark.parametrize("domain", ["local", "global"])
@pytest.mark.parametrize("shape,chunks", [((7, 12, 16, 64), (1, 2, 3, 1)), ((5, 6, 9, 13), (5,)), ((6, 75, 8, 96), (2, 2, 2))])
def test_chunked_fields_are_streamed_block_by_block(ml, variant, domain, shape, chunks):
    """steric(dset) on fields that exist only as blocks along time equals the call on whole arrays bit for bit --
    heights, masses, reference scalars -- for a self-reference and for a supplied reference, and reads every block
    of T and S exactly once (volcello: its first block only)."""
    from momlevel_b200 import synth

    ds = synth.make_dataset(*shape, seed=8, device="cpu", dtype=torch.float32)
    want, wref = ml.steric(ds, variant=variant, domain=domain)
    cds, made = _chunked_dataset(ds, chunks)
    got, gref = ml.steric(cds, variant=variant, domain=domain)
    assert made["thetao"].blocks_read == len(chunks) and made["so"].blocks_read == len(chunks)
    assert made["volcello"].blocks_read == 1
    assert np.array_equal(got[variant].values, want[variant].values, equal_nan=True)
    for k in ("volo", "masso", "rhoga"):
        assert float(gref[k]) == float(wref[k])
    assert np.array_equal(gref["rho"].values, wref["rho"].values, equal_nan=True)
    if domain == "local":
        assert np.array_equal(got["delta_rho"].values, want["delta_rho"].values, equal_nan=True)
    else:
        assert float(got["reference_height"]) == float(want["reference_height"])
    # a supplied reference (steric.py:98-103): another pass over the blocks, same numbers
    again, _ = ml.steric(cds, variant=variant, domain=domain, reference=wref)
    again_w, _ = ml.steric(ds, variant=variant, domain=domain, reference=wref)
    assert np.array_equal(again[variant].values, again_w[variant].values, equal_nan=True)
    # and the oracle
    ref_o, eta_o, _, g_o, _, _ = _oracle_case(ds, variant, "Wright")
    if domain == "local":
        _close_nan(got[variant].values, eta_o, atol=ETA_ATOL)
    else:
        assert np.max(np.abs(got[variant].values - g_o)) < ETA_ATOL


def test_host_stream_global_series_longer_than_memory_budget(ml):
    """BASELINE config 4 in miniature through core.HostStream: a daily global series pushed one day at a time from a
    generator, three variants at once, with at most two blocks alive; equals the resident call bit for bit."""
    from momlevel_b200 import core, synth

    nt, nz, ny, nx = 40, 10, 16, 64
    grid = synth.make_grid(nz, ny, nx, seed=3, device="cpu")
    pres = (grid["z_l"] * 1.0e4 + 101325.0).numpy()
    T, S, V = synth.make_fields(grid, nt, seed=3, dtype=torch.float32)
    want = {v: core.steric_global(T.cuda() if v != "halosteric" else T[0].cuda(), S.cuda() if v != "thermosteric" else S[0].cuda(),
                                  V.cuda(), pres, t_bcast=v == "halosteric", s_bcast=v == "thermosteric").cpu()
            for v in ("steric", "thermosteric", "halosteric")}
    rho_w, sums_w = core.reference_state(T[0].cuda(), S[0].cuda(), V.cuda(), pres)
    hs = core.HostStream("global", V, pres, variants=("steric", "thermosteric", "halosteric"), max_block_steps=3,
                         want_sums=True)
    outs = []
    t = 0
    import weakref

    alive = []
    for n in [1, 3, 2] * 7:
        n = min(n, nt - t)
        if n <= 0:
            break
        blk = (T[t:t + n].clone().numpy(), S[t:t + n].clone().numpy())  # a fresh block, as a reader would produce it
        alive.append(weakref.ref(blk[0]))
        outs.append(hs.push(*blk))
        del blk
        t += n
        assert sum(r() is not None for r in alive) <= 2  # the stream keeps the current and the previous block only
    rho, sums = hs.finish()
    for v in want:
        assert torch.equal(torch.cat([o[v] for o in outs]), want[v])
    assert sums == (float(sums_w[0]), float(sums_w[1]))
    with pytest.raises(RuntimeError):
        hs.push(T[:1].numpy(), S[:1].numpy())  # closed


def test_public_call_on_a_checked_grid_does_not_synchronise(ml):
    """The second ``steric(dset)`` on device-resident fields queues its work and returns: the value checks of the grid
    arrays are remembered for the tensors they were made on, volo / masso are read back when they are looked at."""
    from momlevel_b200 import synth

    ds = synth.make_dataset(5, 10, 16, 64, seed=21, device="cuda", dtype=torch.float32)
    res0, ref0 = ml.steric(ds)
    want_eta = res0["steric"].data.clone()
    want = [float(ref0[k]) for k in ("volo", "masso", "rhoga")]
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")  # any torch-side synchronisation raises
    try:
        res, ref = ml.steric(ds)
        assert ref["volo"].is_lazy and ref["masso"].is_lazy and ref["rhoga"].is_lazy
        assert ref["volo"].shape == () and ref["volo"].dims == ()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert torch.equal(torch.nan_to_num(res["steric"].data, nan=-1.0), torch.nan_to_num(want_eta, nan=-1.0))
    assert [float(ref[k]) for k in ("volo", "masso", "rhoga")] == want
    assert want[2] == want[1] / want[0]

    # the three-height call shares the remembered checks and hands back the same lazy scalars
    first, _ = ml.steric_variants(ds)
    again, vref = ml.steric_variants(ds)
    assert vref["volo"].is_lazy  # the sums of the one-pass kernel are added in another order: equal to rounding
    assert float(vref["volo"]) == pytest.approx(want[0], rel=1e-12) and float(vref["masso"]) == pytest.approx(want[1], rel=1e-12)
    for v in ("steric", "thermosteric", "halosteric"):
        assert torch.equal(torch.nan_to_num(again[v].data, nan=-1.0), torch.nan_to_num(first[v].data, nan=-1.0))
    assert torch.equal(torch.nan_to_num(again["steric"].data, nan=-1.0), torch.nan_to_num(want_eta, nan=-1.0))

    # an in-place write to a grid array voids the entry: the check runs again and fires
    ds["deptho"].data[0, 0] = -1.0
    with pytest.raises(AssertionError, match="Depth values"):
        ml.steric(ds)
    with pytest.raises(AssertionError, match="Depth values"):
        ml.steric_variants(ds)
    ds["deptho"].data[0, 0] = 1.0
    ml.steric(ds)
    # so does another tensor in its place, equal or not
    ds["areacello"] = type(ds["areacello"])(ds["areacello"].data * 10.0, ds["areacello"].dims)
    with pytest.raises(ValueError, match="Errors found"):
        ml.steric(ds)
