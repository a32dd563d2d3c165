"""Pin the oracle: reference known-answer values + golden vectors made from the reference.

Each constant cites the reference test that holds it.  The reference compares with
``np.allclose`` defaults (rtol 1e-5); here the tolerances are as tight as the digits
the reference prints allow, and bit-exact where the reference prints all 16.
"""

import numpy as np
import pytest

from oracle import eos, spice, steric, stratification, testdata


# --------------------------------------------------------------------------- EOS


def test_wright_scalar_kats():
    # tests/test_wright.py:12,31,51,71,121 -- 16-digit constants, reproduced exactly
    assert eos.wright_density(18.0, 35.0, 200000.0) == pytest.approx(1025.359957453976, rel=1e-15)
    assert eos.wright_drho_dtemp(18.0, 35.0, 200000.0) == pytest.approx(-0.24680005918175105, rel=1e-15)
    assert eos.wright_drho_dsal(18.0, 35.0, 200000.0) == pytest.approx(0.7652676800174607, rel=1e-15)
    assert eos.wright_alpha(18.0, 35.0, 200000.0) == pytest.approx(0.0002406960183958898, rel=1e-15)
    assert eos.wright_beta(18.0, 35.0, 200000.0) == pytest.approx(0.0007463405162784603, rel=1e-15)


def test_wright_density_array_kat():
    # tests/test_wright.py:4-27 -- one generator, three draws in this order
    rng = np.random.default_rng(123)
    T = rng.normal(15.0, 5.0, (5, 5))
    S = rng.normal(35.0, 1.5, (5, 5))
    p = rng.normal(2000.0, 500.0, (5, 5))
    expect_row0 = [1026.77225958, 1027.8498461, 1025.60122596, 1026.20882763, 1024.87391971]
    expect_row4 = [1027.02622475, 1024.91713466, 1023.57331842, 1027.21287132, 1024.2578034]
    rho = eos.wright_density(T, S, p)
    assert np.allclose(rho[0], expect_row0, rtol=0, atol=6e-9)
    assert np.allclose(rho[4], expect_row4, rtol=0, atol=6e-9)


def test_linear_kats():
    # tests/test_linear.py:12
    assert eos.linear_density(18.0, 35.0, 200000.0) == pytest.approx(1024.4, rel=1e-15)
    assert eos.linear_drho_dtemp() == -0.2 and eos.linear_drho_dsal() == 0.8


def test_eos_dispatch():
    # util.py:243-249
    assert eos.density("WRIGHT", 18.0, 35.0, 2e5) == eos.wright_density(18.0, 35.0, 2e5)
    with pytest.raises(ValueError):
        eos.density("teos10", 1.0, 1.0, 1.0)
    with pytest.raises(AssertionError):
        eos.density(3, 1.0, 1.0, 1.0)


def test_wright_golden_bit_exact(golden):
    g = golden("eos_wright.npz")
    T, S, p = g["T"], g["S"], g["p"]
    for name, fn in [
        ("density", eos.wright_density),
        ("drho_dtemp", eos.wright_drho_dtemp),
        ("drho_dsal", eos.wright_drho_dsal),
        ("alpha", eos.wright_alpha),
        ("beta", eos.wright_beta),
    ]:
        np.testing.assert_array_equal(fn(T, S, p), g[name], err_msg=name)
    assert np.isnan(g["density"][3]) and np.isnan(g["density"][4])


def test_linear_golden_bit_exact(golden):
    g = golden("eos_linear.npz")
    T, S, p = g["T"], g["S"], g["p"]
    np.testing.assert_array_equal(eos.linear_density(T, S, p), g["density"])
    np.testing.assert_array_equal(eos.linear_density(T, S, p, rho_ref=1035.0), g["density_rho_ref"])
    np.testing.assert_array_equal(eos.linear_alpha(T, S, p), g["alpha"])
    np.testing.assert_array_equal(eos.linear_beta(T, S, p), g["beta"])


# ------------------------------------------------------------------------- spice


def test_flament_kat():
    # tests/test_flament.py:4-13
    S = np.arange(33.0, 37.1, 0.1)
    T = np.arange(0.0, 31.0, 1.0)
    SS = np.tile(S[None, :], (len(T), 1))
    TT = np.tile(T[:, None], (1, len(S)))
    assert spice.flament_spice(TT, SS).sum() == pytest.approx(3283.680384169385, rel=1e-14)


def test_flament_golden(golden):
    g = golden("spice.npz")
    got = spice.flament_spice(g["T"], g["S"])
    assert np.isnan(got[0, 0]) and np.isnan(g["spice"][0, 0])
    m = ~np.isnan(g["spice"])
    # different summation order from the reference: a few ulp of the largest term
    assert np.max(np.abs(got[m] - g["spice"][m])) < 2e-14 * np.max(np.abs(g["spice"][m]))


def test_flament_scalar_and_shape_assert():
    # flament.py:68-75
    assert spice.flament_spice(10.0, 35.0).shape == (1,)
    with pytest.raises(AssertionError):
        spice.flament_spice(np.zeros(3), np.zeros(4))


def test_calc_spice_kat():
    # tests/test_derived.py:135-137
    d = testdata.generate_test_data()
    assert spice.flament_spice(d["thetao"], d["so"]).sum() == pytest.approx(1412.03593361, abs=5e-9)


# ---------------------------------------------------------------------------- dz


def test_calc_dz_kats():
    # tests/test_derived.py:26-45
    d = testdata.generate_test_data_dz()
    assert np.nansum(steric.calc_dz(d["z_l"], d["z_i"], d["deptho"])) == pytest.approx(1130.67307641, abs=5e-9)
    assert np.nansum(steric.calc_dz(d["z_l"], d["z_i"], d["deptho"], fraction=True)) == pytest.approx(
        85.53726628, abs=5e-9
    )
    assert np.nansum(steric.calc_dz(d["z_l"], d["z_i"], d["deptho"], top=12.0, bottom=33.0)) == pytest.approx(
        363.71725794, abs=5e-9
    )
    bad = d["deptho"].copy()
    bad[4, 4] = -200.0
    with pytest.raises(AssertionError):
        steric.calc_dz(d["z_l"], d["z_i"], bad)


# ------------------------------------------------------------------------ steric


@pytest.fixture(scope="module")
def cfg1():
    d = testdata.generate_test_data()
    ref = steric.reference_state(d["thetao"], d["so"], d["volcello"], d["areacello"], d["z_l"])
    return d, ref


def test_reference_state_kats(cfg1):
    # tests/test_steric.py:32-41
    d, ref = cfg1
    assert ref["thetao"].sum() == pytest.approx(1921.05772939, abs=5e-9)
    assert ref["so"].sum() == pytest.approx(4388.81731882, abs=5e-9)
    assert ref["volcello"].sum() == pytest.approx(125921.15458782, abs=5e-9)
    assert ref["rho"].sum() == pytest.approx(128781.63975736, abs=5e-9)
    assert ref["volo"] == pytest.approx(125921.15458782, abs=5e-9)
    assert ref["rhoga"] == pytest.approx(1030.2309221, abs=5e-8)


def test_steric_broadcast(cfg1):
    # tests/test_steric.py:13-22
    d, ref = cfg1
    rho = eos.wright_density(d["thetao"][0, 1, 2, 3], d["so"][0, 1, 2, 3], d["z_l"][1] * 1.0e4 + 101325.0)
    assert ref["rho"][1, 2, 3] == rho


@pytest.mark.parametrize(
    "variant,eta_sum,drho_sum",
    [
        ("steric", 1.38250197, -11.33133173),  # tests/test_steric.py:64-65
        ("thermosteric", -4.14327109, 33.83631611),  # :76-77
        ("halosteric", 4.39398075, -32.07946717),  # :52-53
    ],
)
def test_local_kats(cfg1, variant, eta_sum, drho_sum):
    d, ref = cfg1
    eta, drho = steric.steric_local(d["thetao"], d["so"], d["z_l"], d["z_i"], d["deptho"], ref, variant=variant)
    assert eta.shape == (5, 5, 5) and drho.shape == (5, 5, 5, 5)
    assert np.nansum(eta) == pytest.approx(eta_sum, abs=5e-9)
    assert np.nansum(drho) == pytest.approx(drho_sum, abs=5e-9)


def test_supplied_reference_kat(cfg1):
    # tests/test_steric.py:128-137
    d, _ = cfg1
    d2 = testdata.generate_test_data(seed=999)
    ref2 = steric.reference_state(d2["thetao"], d2["so"], d2["volcello"], d2["areacello"], d2["z_l"])
    assert ref2["thetao"].sum() == pytest.approx(1917.31113456, abs=5e-9)
    assert ref2["so"].sum() == pytest.approx(4387.69334037, abs=5e-9)
    assert ref2["volcello"].sum() == pytest.approx(125846.22269117, abs=5e-9)
    assert ref2["rho"].sum() == pytest.approx(128780.12974804, abs=5e-9)
    eta, _ = steric.steric_local(d["thetao"], d["so"], d["z_l"], d["z_i"], d["deptho"], ref2)
    assert np.nansum(eta) == pytest.approx(1.25554742, abs=5e-9)


def test_global_follows_reference_code(cfg1):
    # tests/test_steric.py:80-125 hold constants below their own atol:
    # they pin nothing.  What the reference *code* (steric.py:134-147) evaluates to:
    d, ref = cfg1
    expect = {"steric": 1.8975448e-14, "thermosteric": -2.5032375e-13, "halosteric": 2.3429616e-13}
    for variant, val in expect.items():
        eta, href, masso = steric.steric_global(d["thetao"], d["so"], d["z_l"], ref, variant=variant)
        assert eta.shape == (5,) and masso.shape == (5,)
        assert eta[0] == 0.0 or variant != "steric"  # t=0 is the reference itself
        assert eta.sum() == pytest.approx(val, rel=1e-6)
        assert href == pytest.approx(3.4870492e-10, rel=1e-7)
        # and the reference's stale constants still pass its own vacuous check
        assert np.allclose(eta.sum(), {"steric": 6.29048941e-14, "thermosteric": -1.38053154e-13,
                                        "halosteric": 1.98293992e-13}[variant])


def test_masso_kat(cfg1):
    # tests/test_derived.py:84-87: calc_masso(rho4d, volcello4d).sum() with pres = z_l*1e4 (no patm)
    d, _ = cfg1
    rho = eos.wright_density(d["thetao"], d["so"], (d["z_l"] * 1.0e4)[None, :, None, None])
    masso = np.nansum(rho * d["volcello"], axis=(1, 2, 3))
    assert masso.sum() == pytest.approx(6.45215577e08, rel=2e-9)


def test_nan_semantics():
    """Land / missing cells: unpinned upstream, follows steric.py:151-166 literally."""
    d = testdata.generate_test_data()
    T, S, V = d["thetao"].copy(), d["so"].copy(), d["volcello"].copy()
    V[:, :, 0, 0] = np.nan  # land column (volume missing everywhere)
    T[:, :, 0, 0] = np.nan
    S[:, :, 0, 0] = np.nan
    V[:, 3:, 1, 1] = np.nan  # shallow column: bottom two cells missing
    T[:, 3:, 1, 1] = np.nan
    S[:, 3:, 1, 1] = np.nan
    T[2, 1, 2, 2] = np.nan  # a transient hole in T only
    ref = steric.reference_state(T, S, V, d["areacello"], d["z_l"])
    eta, drho = steric.steric_local(T, S, d["z_l"], d["z_i"], d["deptho"], ref)
    assert np.all(np.isnan(eta[:, 0, 0]))  # masked by surface volcello
    assert np.all(np.isfinite(eta[:, 1, 1]))  # partial column still integrates
    assert np.isfinite(eta[2, 2, 2]) and np.isnan(drho[2, 1, 2, 2])  # hole is skipped
    assert np.all(eta[0][np.isfinite(eta[0])] == 0.0)
    g, href, masso = steric.steric_global(T, S, d["z_l"], ref)
    assert np.all(np.isfinite(g)) and g[0] == 0.0


# ---------------------------------------------------------------- stratification (next row)


def test_calc_n2_kats():
    # tests/test_derived.py:54-61 and :14-18
    d = testdata.generate_test_data()
    n2 = stratification.calc_n2(d["thetao"], d["so"], d["z_l"])
    assert n2.shape == (5, 5, 5, 5)
    assert n2.sum() == pytest.approx(0.00338354, abs=5e-9)
    adj = stratification.calc_n2(d["thetao"], d["so"], d["z_l"], adjust_negative=True)
    assert np.nansum(adj) == pytest.approx(0.12093286, abs=5e-9)
    assert np.array_equal(adj, stratification.adjust_negative_n2(n2), equal_nan=True)
    # time 0 is filled everywhere (the reference's `adjusted[0]` is the first axis), later steps only forward-filled
    assert np.all(adj[0] > 0) and np.all(adj[np.isfinite(adj)] > 0)


def test_stability_angle_and_wave_speed_kats():
    # tests/test_derived.py:140-151
    d = testdata.generate_test_data()
    tu = stratification.calc_stability_angle(d["thetao"], d["so"], d["z_l"] * 1.0e4, d["z_l"])
    assert tu.shape == (5, 5, 5, 5)
    assert tu.sum() == pytest.approx(5838.68533435, abs=5e-8)
    n2 = stratification.calc_n2(d["thetao"], d["so"], d["z_l"])
    dz = steric.calc_dz(d["z_l"], d["z_i"], d["deptho"])
    c1, broadcast = stratification.calc_wave_speed(n2, dz)
    assert c1.shape == (5, 5, 5) and broadcast.shape == (5, 5, 5, 5)  # (t,y,x) and the reference's (z,y,x,t)
    assert broadcast.sum() == pytest.approx(524.30956095, abs=5e-8)
    # no missing values here: the broadcast repeats the column sums over the five levels
    assert np.array_equal(broadcast[2], np.moveaxis(c1, 0, -1))


# ------------------------------------------------------------------ properties of the oracle itself


def test_dz_partitions_the_water_column():
    # derived.py:295-318: the clipped thicknesses of a column add up to its depth (capped by the grid's bottom)
    rng = np.random.default_rng(5)
    z_i = np.concatenate([[0.0], np.cumsum(rng.uniform(1.0, 300.0, 20))])
    z_l = 0.5 * (z_i[1:] + z_i[:-1])
    depth = rng.uniform(0.0, 1.2 * z_i[-1], (7, 9))
    depth[0, 0] = np.nan  # land: filled with 0 (derived.py:295)
    dz = steric.calc_dz(z_l, z_i, depth)
    assert dz.shape == (20, 7, 9) and np.all(dz >= 0)
    assert np.allclose(dz.sum(0), np.minimum(np.nan_to_num(depth), z_i[-1]), rtol=0, atol=1e-9)
    frac = steric.calc_dz(z_l, z_i, depth, fraction=True)
    wet = np.isfinite(frac)  # derived.py:320-323: empty cells become NaN, not 0
    assert np.array_equal(wet, dz > 0) and np.all((frac[wet] > 0) & (frac[wet] <= 1))


def test_linear_eos_heights_are_additive():
    # with a linear equation of state rho(T,S) - rho0 = [rho(T,S0) - rho0] + [rho(T0,S) - rho0]
    d = testdata.generate_test_data()
    ref = steric.reference_state(d["thetao"], d["so"], d["volcello"], d["areacello"], d["z_l"], eos="linear")
    eta = {v: steric.steric_local(d["thetao"], d["so"], d["z_l"], d["z_i"], d["deptho"], ref, eos="linear", variant=v)[0]
           for v in steric.VARIANTS}
    assert np.allclose(eta["steric"], eta["thermosteric"] + eta["halosteric"], rtol=0, atol=1e-12)
    assert np.all(eta["steric"][0] == 0.0)  # step 0 is the reference state


def test_adjusted_n2_is_positive_and_idempotent():
    d = testdata.generate_test_data()
    n2 = stratification.calc_n2(d["thetao"], d["so"], d["z_l"])
    adj = stratification.adjust_negative_n2(n2)
    assert np.all(adj[np.isfinite(adj)] > 0)
    assert np.array_equal(stratification.adjust_negative_n2(adj), adj, equal_nan=True)
    # values that were positive are untouched
    keep = n2 > 0
    assert np.array_equal(adj[keep], n2[keep])
