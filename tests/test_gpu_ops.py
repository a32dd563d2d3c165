"""GPU parity of the elementwise operators against the oracle and the reference's goldens.

Tolerances (BASELINE.json north_star): density within 1e-10 relative of the reference's
numpy path on the same fp64-upcast inputs.  Everything goes through the C ABI.
"""

import numpy as np
import pytest
import torch

from oracle import eos as oeos
from oracle import spice as ospice
from oracle import steric as osteric
from oracle import testdata

pytestmark = pytest.mark.gpu

RHO_RTOL = 1e-10  # north_star: 1e-10 relative on density


@pytest.fixture(scope="module")
def ml():
    import momlevel_b200

    assert torch.cuda.is_available(), "GPU tests need a CUDA device: there is no CPU path"
    return momlevel_b200


def _relerr(a, b):
    m = ~np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    return np.max(np.abs(a[m] - b[m]) / np.abs(b[m])) if m.any() else 0.0


# ------------------------------------------------------------------------------ Wright


def test_wright_scalar_kats(ml):
    # tests/test_wright.py:12,31,51,71,121
    w = ml.eos.wright
    assert w.density(18.0, 35.0, 200000.0) == pytest.approx(1025.359957453976, rel=RHO_RTOL)
    assert w.drho_dtemp(18.0, 35.0, 200000.0) == pytest.approx(-0.24680005918175105, rel=1e-12)
    assert w.drho_dsal(18.0, 35.0, 200000.0) == pytest.approx(0.7652676800174607, rel=1e-12)
    assert w.alpha(18.0, 35.0, 200000.0) == pytest.approx(0.0002406960183958898, rel=1e-12)
    assert w.beta(18.0, 35.0, 200000.0) == pytest.approx(0.0007463405162784603, rel=1e-12)


def test_wright_array_kat(ml):
    # tests/test_wright.py:4-27
    rng = np.random.default_rng(123)
    T, S, p = rng.normal(15.0, 5.0, (5, 5)), rng.normal(35.0, 1.5, (5, 5)), rng.normal(2000.0, 500.0, (5, 5))
    rho = ml.eos.wright.density(T, S, p)
    assert rho.shape == (5, 5) and rho.dtype == np.float64
    assert np.allclose(rho[0], [1026.77225958, 1027.8498461, 1025.60122596, 1026.20882763, 1024.87391971],
                       rtol=0, atol=6e-9)
    assert _relerr(rho, oeos.wright_density(T, S, p)) < RHO_RTOL


@pytest.mark.parametrize("func", ["density", "drho_dtemp", "drho_dsal", "alpha", "beta"])
def test_wright_golden(ml, golden, func):
    g = golden("eos_wright.npz")
    got = getattr(ml.eos.wright, func)(g["T"], g["S"], g["p"])
    assert _relerr(got, g[func]) < (RHO_RTOL if func == "density" else 1e-11)


def test_wright_fp32_storage_equals_fp64_upcast(ml, golden):
    """fp32 T/S are widened exactly on load: same result as handing in the fp64 upcast."""
    g = golden("eos_wright.npz")
    T32, S32 = g["T"].astype(np.float32), g["S"].astype(np.float32)
    a = ml.eos.wright.density(T32, S32, g["p"])
    b = ml.eos.wright.density(T32.astype(np.float64), S32.astype(np.float64), g["p"])
    np.testing.assert_array_equal(a, b)
    assert _relerr(a, oeos.wright_density(T32.astype(np.float64), S32.astype(np.float64), g["p"])) < RHO_RTOL


def test_wright_wide_range_and_broadcast(ml):
    rng = np.random.default_rng(7)
    T = rng.uniform(-2, 40, (3, 7, 11)).astype(np.float32).astype(np.float64)
    S = rng.uniform(0, 42, (3, 7, 11)).astype(np.float32).astype(np.float64)
    p = rng.uniform(1e5, 7e7, (7, 1))  # numpy broadcasting against (3,7,11)
    assert _relerr(ml.eos.wright.density(T, S, p), oeos.wright_density(T, S, p)) < RHO_RTOL
    # scalar T against array S
    assert _relerr(ml.eos.wright.density(10.0, S, 2e5), oeos.wright_density(10.0, S, 2e5)) < RHO_RTOL


def test_wright_degenerate_denominator_matches_ieee(ml):
    """Far outside the ocean range the lean reciprocal hands over to IEEE division."""
    # T = 1e78 puts the denominator at ~7e307, where 1/d is subnormal
    T = np.array([1e78, 9e77, 1e120, 18.0])
    S = np.array([35.0, 35.0, 35.0, 35.0])
    with np.errstate(all="ignore"):
        want = oeos.wright_density(T, S, 2e5)
    got = ml.eos.wright.density(T, S, 2e5)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = np.isfinite(want)
    assert m[:2].all() and np.all(want[:2] > 0)
    assert np.allclose(got[m], want[m], rtol=1e-10, atol=0)


def test_empty_input(ml):
    out = ml.eos.wright.density(np.zeros((0, 4)), np.zeros((0, 4)), 1e5)
    assert out.shape == (0, 4)
    assert ml.spice.flament.spice(np.zeros(0), np.zeros(0)).shape == (0,)


def test_device_tensors_stay_on_device(ml):
    T = torch.full((4, 6), 18.0, device="cuda", dtype=torch.float32)
    S = torch.full((4, 6), 35.0, device="cuda", dtype=torch.float32)
    rho = ml.eos.wright.density(T, S, 200000.0)
    assert rho.is_cuda and rho.dtype == torch.float64
    assert float(rho[0, 0]) == pytest.approx(1025.359957453976, rel=RHO_RTOL)


# ------------------------------------------------------------------------------ linear


def test_linear_kats_and_golden(ml, golden):
    # tests/test_linear.py:12
    lin = ml.eos.linear
    assert lin.density(18.0, 35.0, 200000.0) == pytest.approx(1024.4, rel=1e-15)
    assert lin.drho_dtemp() == -0.2 and lin.drho_dsal() == 0.8
    g = golden("eos_linear.npz")
    assert _relerr(lin.density(g["T"], g["S"], g["p"]), g["density"]) < 1e-14
    assert _relerr(lin.alpha(g["T"], g["S"], g["p"]), g["alpha"]) < 1e-14
    assert _relerr(lin.beta(g["T"], g["S"], g["p"]), g["beta"]) < 1e-14
    m = ~np.isnan(g["density_rho_ref"])
    assert np.max(np.abs(lin.density(g["T"], g["S"], None, rho_ref=1035.0)[m] - g["density_rho_ref"][m])) < 1e-12


# ------------------------------------------------------------------------------- spice


def test_flament_kat(ml):
    # tests/test_flament.py:4-13
    S = np.arange(33.0, 37.1, 0.1)
    T = np.arange(0.0, 31.0, 1.0)
    SS = np.tile(S[None, :], (len(T), 1))
    TT = np.tile(T[:, None], (1, len(S)))
    pi = ml.spice.flament.spice(TT, SS)
    assert pi.shape == TT.shape
    assert pi.sum() == pytest.approx(3283.680384169385, rel=1e-13)


def test_flament_golden_and_scalar(ml, golden):
    g = golden("spice.npz")
    got = ml.spice.flament.spice(g["T"], g["S"])
    assert np.array_equal(np.isnan(got), np.isnan(g["spice"]))
    m = ~np.isnan(g["spice"])
    assert np.max(np.abs(got[m] - g["spice"][m])) < 1e-13
    assert ml.spice.flament.spice(10.0, 35.0).shape == (1,)
    with pytest.raises(AssertionError):
        ml.spice.flament.spice(np.zeros(3), np.zeros(4))


def test_calc_spice_kat(ml):
    # tests/test_derived.py:135-137
    d = ml.test_data.generate_test_data()
    pi = ml.derived.calc_spice(d["thetao"], d["so"])
    assert pi.dims == d["thetao"].dims and pi.attrs["long_name"] == "Sea water spiciness"
    assert float(pi.sum()) == pytest.approx(1412.03593361, abs=5e-9)


def test_spice_large_fp32(ml):
    rng = np.random.default_rng(5)
    T = rng.uniform(-2, 32, 1_000_003).astype(np.float32)
    S = rng.uniform(30, 40, 1_000_003).astype(np.float32)
    T[17] = np.nan
    got = ml.spice.flament.spice(T, S)
    want = ospice.flament_spice(T.astype(np.float64), S.astype(np.float64))
    assert np.isnan(got[17])
    m = ~np.isnan(want)
    assert np.max(np.abs(got[m] - want[m])) < 1e-12


# ---------------------------------------------------------------------------------- dz


def test_calc_dz_kats(ml):
    # tests/test_derived.py:26-45
    d = ml.test_data.generate_test_data_dz()
    dz = ml.derived.calc_dz(d.z_l, d.z_i, d.deptho)
    assert dz.dims == ("yh", "xh", "z_l")
    assert float(dz.sum()) == pytest.approx(1130.67307641, abs=5e-9)
    assert float(ml.derived.calc_dz(d.z_l, d.z_i, d.deptho, fraction=True).sum()) == pytest.approx(85.53726628, abs=5e-9)
    assert float(ml.derived.calc_dz(d.z_l, d.z_i, d.deptho, top=12.0, bottom=33.0).sum()) == pytest.approx(
        363.71725794, abs=5e-9)
    o = testdata.generate_test_data_dz()
    want = osteric.calc_dz(o["z_l"], o["z_i"], o["deptho"])
    np.testing.assert_array_equal(dz.transpose("z_l", "yh", "xh").values, want)
    bad = d.deptho.copy()
    bad[4, 4] = -200.0
    with pytest.raises(AssertionError):
        ml.derived.calc_dz(d.z_l, d.z_i, bad)


# ------------------------------------------------------------------ derived.calc_rho etc


def test_calc_rho_and_friends(ml):
    d = ml.test_data.generate_test_data()
    o = testdata.generate_test_data()
    pres = d["z_l"] * 1.0e4
    rho = ml.derived.calc_rho(d["thetao"], d["so"], pres, eos="Wright")
    want = oeos.wright_density(o["thetao"], o["so"], (o["z_l"] * 1e4)[None, :, None, None])
    assert rho.dims == d["thetao"].dims and rho.attrs["units"] == "kg m-3"
    assert _relerr(rho.values, want) < RHO_RTOL
    # tests/test_derived.py:48-51 holds a stale constant (643872.597, rel 4.7e-6 from the code);
    # the reference's own tolerance still accepts the code's value
    assert np.allclose(float(rho.sum()), 643872.59725673)
    # tests/test_derived.py:73-81
    assert float(ml.derived.calc_alpha(d["thetao"], d["so"], pres).sum()) == pytest.approx(0.14302587, abs=5e-9)
    assert float(ml.derived.calc_beta(d["thetao"], d["so"], pres).sum()) == pytest.approx(0.4639801, abs=5e-8)
    # tests/test_derived.py:84-103
    masso = ml.derived.calc_masso(rho, d["volcello"])
    assert masso.dims == ("time",)
    assert float(masso.sum()) == pytest.approx(6.45215577e08, rel=2e-9)
    assert float(ml.derived.calc_volo(d["volcello"].isel(time=0))) == pytest.approx(125921.15458782, abs=5e-9)
    with pytest.raises(AssertionError):
        ml.derived.calc_volo(d["volcello"])
    with pytest.raises(ValueError):
        ml.derived.calc_rho(d["thetao"], d["so"], pres, eos="nope")
    # potential density = density at a fixed pressure (derived.py:477)
    pd = ml.derived.calc_pdens(d["thetao"], d["so"], level=2000.0)
    assert _relerr(pd.values, oeos.wright_density(o["thetao"], o["so"], 2000.0 * 1e4 + 101325)) < RHO_RTOL


def test_inverse_barometer_kat(ml):
    # tests/test_dynamic.py:6-11
    d = ml.test_data.generate_test_data().isel(z_l=0)
    result = ml.inverse_barometer(d["thetao"], d["so"], 101325.0)
    assert result.attrs == {"long_name": "Inverse Barometer Height", "units": "m"}
    assert float(result.sum()) == pytest.approx(-1259.79345168, abs=5e-9)


# ------------------------------------------------------ ring-staged streaming kernels (csrc/ml_stream.cu)


@pytest.mark.parametrize("n", [4, 2044, 2048, 2052, 5 * 2048 + 8, 148 * 2 * 4 * 2048 + 4 * 377])
def test_stream_spice_against_oracle_and_plain_kernel(ml, n):
    """fp32, n % 4 == 0: the ring-staged kernel -- whole tiles, a short last tile, fewer tiles than CTAs, and more
    tiles than one pass of the ring -- against the oracle and bit for bit against the plain kernel."""
    from momlevel_b200 import core

    rng = np.random.default_rng(n)
    T = rng.uniform(-2, 32, n).astype(np.float32)
    S = rng.uniform(30, 40, n).astype(np.float32)
    T[n // 2] = np.nan
    S[n - 1] = np.nan
    Td, Sd = torch.from_numpy(T).cuda(), torch.from_numpy(S).cuda()
    before = core.launch_count()
    got = core.flament_spice(Td, Sd)
    assert core.launch_count() == before + 1
    prev = core.force_direct(1)
    try:
        plain = core.flament_spice(Td, Sd)
    finally:
        core.force_direct(prev)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(plain, nan=-7.0))
    m = slice(None) if n <= 1 << 20 else slice(0, None, 97)
    want = ospice.flament_spice(T[m].astype(np.float64), S[m].astype(np.float64))
    g = got.cpu().numpy()[m]
    assert np.array_equal(np.isnan(g), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.max(np.abs(g[ok] - want[ok])) < 1e-12


@pytest.mark.parametrize("eos", ["Wright", "linear"])
@pytest.mark.parametrize("shape,bcast", [((3, 7, 4, 516), None), ((2, 5, 8, 1024), "t"), ((2, 5, 8, 1024), "s"),
                                         ((1, 75, 16, 160), None), ((4, 3, 1, 8), None)])
def test_stream_density_rows_pressure_and_broadcast(ml, eos, shape, bcast):
    """eos.<name>.density over [t][z][y][x] with one pressure per level: rows that are not a whole number of tiles,
    a broadcast operand (steric.py:115-121), against the oracle (1e-10 relative) and the plain kernel (bit for bit)."""
    from momlevel_b200 import core
    from oracle import eos as oeos

    nt, nz, ny, nx = shape
    rng = np.random.default_rng(nt * 1000 + nz)
    T = rng.uniform(-2, 32, shape).astype(np.float32)
    S = rng.uniform(30, 40, shape).astype(np.float32)
    T[0, 0, 0, 1] = np.nan
    p = np.linspace(1e5, 6e7, nz)
    Td, Sd = torch.from_numpy(T).cuda(), torch.from_numpy(S).cuda()
    Tin = Td[0].contiguous() if bcast == "t" else Td
    Sin = Sd[0].contiguous() if bcast == "s" else Sd
    kw = dict(z_axis=1, t_bcast=bcast == "t", s_bcast=bcast == "s")
    got = core.eos_eval(eos, "density", Tin, Sin, p, **kw)
    prev = core.force_direct(1)
    try:
        plain = core.eos_eval(eos, "density", Tin, Sin, p, **kw)
    finally:
        core.force_direct(prev)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(plain, nan=-7.0))
    T64 = (T[0:1] if bcast == "t" else T).astype(np.float64)
    S64 = (S[0:1] if bcast == "s" else S).astype(np.float64)
    want = np.broadcast_to(oeos.density(eos, T64, S64, p[None, :, None, None]), shape)
    g = got.cpu().numpy()
    assert np.array_equal(np.isnan(g), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.max(np.abs(g[ok] - want[ok]) / np.abs(want[ok])) < RHO_RTOL
    # a scalar pressure takes the same kernel
    got_s = core.eos_eval(eos, "density", Td, Sd, 2.0e5)
    want_s = oeos.density(eos, T.astype(np.float64), S.astype(np.float64), 2.0e5)
    ok = ~np.isnan(want_s)
    assert np.max(np.abs(got_s.cpu().numpy()[ok] - want_s[ok]) / np.abs(want_s[ok])) < RHO_RTOL


@pytest.mark.parametrize("eos", ["Wright", "linear"])
@pytest.mark.parametrize("shape", [(7, 4, 516), (75, 16, 160), (3, 1, 8), (12, 60, 1024)])
def test_stream_reference_state(ml, eos, shape):
    """reference.py:71-80 through the ring-staged kernel: rho_ref bit for bit with the plain kernel and within
    1e-10 of the oracle, volo / masso (another summation order) to 1e-13."""
    from momlevel_b200 import core
    from oracle import eos as oeos

    nz, ny, nx = shape
    rng = np.random.default_rng(nz * 31 + nx)
    T = rng.uniform(-2, 32, shape).astype(np.float32)
    S = rng.uniform(30, 40, shape).astype(np.float32)
    V = rng.uniform(1e6, 1e9, shape).astype(np.float32)
    V[:, 0, :3] = np.nan
    T[0, 0, 5] = np.nan  # a hole where the volume is present: skipped by masso, kept by volo
    p = np.linspace(1e5, 6e7, nz)
    args = [torch.from_numpy(x).cuda() for x in (T, S, V)]
    rho, sums = core.reference_state(*args, p, eos=eos)
    prev = core.force_direct(1)
    try:
        rho_p, sums_p = core.reference_state(*args, p, eos=eos)
    finally:
        core.force_direct(prev)
    assert torch.equal(torch.nan_to_num(rho, nan=-7.0), torch.nan_to_num(rho_p, nan=-7.0))
    assert torch.allclose(sums, sums_p, rtol=1e-13, atol=0)
    want = oeos.density(eos, T.astype(np.float64), S.astype(np.float64), p[:, None, None])
    ok = ~np.isnan(want)
    assert np.max(np.abs(rho.cpu().numpy()[ok] - want[ok]) / np.abs(want[ok])) < RHO_RTOL
    V64 = V.astype(np.float64)
    assert float(sums[0]) == pytest.approx(np.nansum(V64), rel=1e-13)
    assert float(sums[1]) == pytest.approx(np.nansum(want * V64), rel=1e-13)


@pytest.mark.parametrize("eos", ["Wright", "linear"])
@pytest.mark.parametrize("shape,bcast", [((7, 5, 4, 1028), None), ((3, 75, 8, 160), "t"), ((3, 75, 8, 160), "s"),
                                         ((1, 2, 1, 4), None), ((9, 3, 2, 2052), None)])
def test_stream_delta_rho(ml, eos, shape, bcast):
    """steric.py:151-153 through the ring-staged kernel: rows that are not whole tiles, more (level, segment) pairs than
    CTAs, a broadcast operand, a volume mask -- bit for bit with the plain kernel, and against the oracle."""
    from momlevel_b200 import core
    from oracle import eos as oeos

    nt, nz, ny, nx = shape
    rng = np.random.default_rng(nt * 100 + nz)
    T = rng.uniform(-2, 32, shape).astype(np.float32)
    S = rng.uniform(30, 40, shape).astype(np.float32)
    V = rng.uniform(1e6, 1e9, shape[1:]).astype(np.float32)
    V[:, 0, :2] = np.nan
    if nt > 1:
        T[nt - 1, 0, 0, 3] = np.nan
    p = np.linspace(1e5, 6e7, nz)
    Td, Sd, Vd = (torch.from_numpy(x).cuda() for x in (T, S, V))
    rho_ref, _ = core.reference_state(Td[0], Sd[0], Vd, p, eos=eos)
    Tin = Td[0].contiguous() if bcast == "t" else Td
    Sin = Sd[0].contiguous() if bcast == "s" else Sd
    kw = dict(eos=eos, t_bcast=bcast == "t", s_bcast=bcast == "s")
    got = core.delta_rho(Tin, Sin, rho_ref, Vd, p, **kw)
    prev = core.force_direct(1)
    try:
        plain = core.delta_rho(Tin, Sin, rho_ref, Vd, p, **kw)
    finally:
        core.force_direct(prev)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(plain, nan=-7.0))
    T64 = (T[0:1] if bcast == "t" else T).astype(np.float64)
    S64 = (S[0:1] if bcast == "s" else S).astype(np.float64)
    rho = np.broadcast_to(oeos.density(eos, T64, S64, p[None, :, None, None]), shape)
    ref = oeos.density(eos, T[0].astype(np.float64), S[0].astype(np.float64), p[:, None, None])
    want = np.where(~np.isnan(V)[None], rho - ref[None], np.nan)
    g = got.cpu().numpy()
    assert np.array_equal(np.isnan(g), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.max(np.abs(g[ok] - want[ok])) < 1e-9


@pytest.mark.parametrize("a_dtype,w_dtype", [(np.float64, np.float32), (np.float64, np.float64), (np.float32, np.float32),
                                             (np.float32, None)])
@pytest.mark.parametrize("nrows,n", [(1, 7), (3, 4099), (12, 1 << 18), (1, 3_000_001)])
def test_weighted_nansum(ml, a_dtype, w_dtype, nrows, n):
    """ml_calc_masso: (rho * volcello).sum() per time step with NaNs skipped (derived.py:435-438) and, without
    weights, volcello.sum() (derived.py:787-789); bitwise reproducible from call to call."""
    from momlevel_b200 import core

    rng = np.random.default_rng(n + nrows)
    a = rng.uniform(990.0, 1060.0, (nrows, n)).astype(a_dtype)
    a[0, n // 2] = np.nan
    w = None
    if w_dtype is not None:
        w = rng.uniform(1e6, 1e9, n).astype(w_dtype)
        w[: max(1, n // 10)] = np.nan
    got = core.weighted_nansum(torch.from_numpy(a).cuda(), None if w is None else torch.from_numpy(w).cuda(), nrows=nrows)
    again = core.weighted_nansum(torch.from_numpy(a).cuda(), None if w is None else torch.from_numpy(w).cuda(), nrows=nrows)
    assert torch.equal(got, again)
    a64 = a.astype(np.float64)
    want = np.nansum(a64 * w.astype(np.float64)[None], axis=1) if w is not None else np.nansum(a64, axis=1)
    assert np.allclose(got.cpu().numpy(), want, rtol=1e-13, atol=0)
