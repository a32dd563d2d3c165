"""GPU parity of the stratification diagnostics against the oracle.

``calc_n2`` / ``adjust_negative_n2`` / ``calc_stability_angle`` / ``calc_wave_speed`` follow
``src/momlevel/derived.py:30-71, 328-411, 714-828``; the known-answer sums are the reference's own
(``tests/test_derived.py:14-18, 54-61, 140-151``).  Tolerances: the kernels contract the three-point
stencil into FMAs and take alpha, beta from one shared division, so values differ from numpy's by
rounding: 1e-10 relative to the magnitude of the field (BASELINE's density tolerance carried over).
"""

import numpy as np
import pytest
import torch

from oracle import steric as osteric
from oracle import stratification as ostrat
from oracle import testdata

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ml():
    import momlevel_b200

    assert torch.cuda.is_available(), "GPU tests need a CUDA device: there is no CPU path"
    return momlevel_b200


@pytest.fixture(scope="module")
def dset(ml):
    return ml.test_data.generate_test_data()


def _close(got, want, rtol=1e-10):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    m = np.isfinite(want)
    assert np.array_equal(np.isinf(got), np.isinf(want))
    if m.any():
        scale = np.abs(want[m]).max()
        assert np.abs(got[m] - want[m]).max() <= rtol * scale, f"max err {np.abs(got[m] - want[m]).max():.3e} vs scale {scale:.3e}"


def test_reference_kats(ml, dset):
    n2 = ml.derived.calc_n2(dset["thetao"], dset["so"])
    assert n2.dims == ("time", "z_l", "yh", "xh")
    assert n2.attrs["standard_name"] == "square_of_brunt_vaisala_frequency_in_sea_water"
    assert float(n2.sum()) == pytest.approx(0.00338354, abs=5e-9)  # tests/test_derived.py:54-56
    adj = ml.derived.calc_n2(dset["thetao"], dset["so"], adjust_negative=True)
    assert float(adj.sum()) == pytest.approx(0.12093286, abs=5e-9)  # :59-61
    adj2 = ml.derived.adjust_negative_n2(n2)
    assert float(adj2.sum()) == pytest.approx(0.12093286, abs=5e-9)  # :15-18
    assert adj2.attrs["comment"] == "adjustment applied for negative values"
    assert np.array_equal(adj.values, adj2.values, equal_nan=True)
    tu = ml.derived.calc_stability_angle(dset["thetao"], dset["so"], dset["z_l"] * 1.0e4, eos="Wright")
    assert tu.name == "tu_angle" and tu.attrs["units"] == "degrees"
    assert float(tu.sum()) == pytest.approx(5838.68533435, abs=5e-7)  # :140-144
    dz = ml.derived.calc_dz(dset["z_l"], dset["z_i"], dset["deptho"])
    ws = ml.derived.calc_wave_speed(n2, dz)
    assert ws.dims == ("z_l", "yh", "xh", "time")  # the reference's name-based broadcast of n2[0] against the sums
    assert float(ws.sum()) == pytest.approx(524.30956095, abs=5e-7)  # :147-151


@pytest.mark.parametrize("eos", ["Wright", "linear"])
def test_config1_against_oracle(ml, dset, eos):
    o = testdata.generate_test_data()
    for adjust in (False, True):
        got = ml.derived.calc_n2(dset["thetao"], dset["so"], eos=eos, adjust_negative=adjust).values
        _close(got, ostrat.calc_n2(o["thetao"], o["so"], o["z_l"], eos=eos, adjust_negative=adjust))
    got = ml.derived.calc_stability_angle(dset["thetao"], dset["so"], dset["z_l"] * 1.0e4, eos=eos).values
    _close(got, ostrat.calc_stability_angle(o["thetao"], o["so"], o["z_l"] * 1.0e4, o["z_l"], eos=eos), rtol=1e-9)


@pytest.mark.parametrize("shape", [(3, 10, 37, 53), (2, 75, 24, 128), (1, 3, 4, 7)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_synthetic_ocean_against_oracle(ml, shape, dtype):
    """Land, sea floor and transient holes: any missing value in the three-level stencil voids the cell."""
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(*shape, seed=17, device="cuda", dtype=dtype)
    T, S = ds["thetao"].data.clone(), ds["so"].data.clone()
    if shape[0] > 1:
        T[1, shape[1] // 2, 1, 2] = float("nan")
        S[0, 0, 2, 3] = float("nan")
    T64, S64 = T.double().cpu().numpy(), S.double().cpu().numpy()
    z_l = ds["z_l"].values
    want = ostrat.calc_n2(T64, S64, z_l)
    got = core.calc_n2(T, S, z_l)
    assert got.dtype == torch.float64 and got.is_cuda
    _close(got.cpu().numpy(), want)
    wadj = ostrat.calc_n2(T64, S64, z_l, adjust_negative=True)
    _close(core.calc_n2(T, S, z_l, adjust_negative=True).cpu().numpy(), wadj)
    _close(core.adjust_negative_n2(got).cpu().numpy(), ostrat.adjust_negative_n2(want))
    # 3-D field: the level axis leads and adjusted[0] is the surface level
    _close(core.adjust_negative_n2(got[-1], z_axis=0).cpu().numpy(), ostrat.adjust_negative_n2(want[-1], z_axis=0))
    pres = z_l * 1.0e4
    tu = core.stability_angle(T, S, pres, z_l).cpu().numpy()
    wtu = ostrat.calc_stability_angle(T64, S64, pres, z_l)
    # the angle is ill-conditioned where the density ratio is close to one; compare where it is not
    ok = np.isfinite(wtu)
    assert np.array_equal(np.isnan(tu), np.isnan(wtu))
    assert np.abs(tu[ok] - wtu[ok]).max() < 1e-6
    dz = osteric.calc_dz(z_l, ds["z_i"].values, ds["deptho"].values)
    c1, _ = ostrat.calc_wave_speed(want, dz)
    _close(core.wave_speed(got, torch.from_numpy(dz).cuda()).cpu().numpy(), c1)


def test_argument_errors(ml):
    from momlevel_b200 import core

    T = torch.zeros((2, 2, 8), device="cuda")
    with pytest.raises(ml._lib.MLError):  # numpy.gradient(edge_order=2) needs three levels
        core.calc_n2(T, T, np.array([1.0, 2.0]), z_axis=1)
    with pytest.raises(NotImplementedError):
        ml.derived.calc_n2(ml.DataArray(T, ("time", "z_l", "xh"), coords={"z_l": np.array([1.0, 2.0])}),
                           ml.DataArray(T, ("time", "z_l", "xh")), interfaces=object())
    with pytest.raises(ValueError):
        ml.derived.calc_n2(ml.DataArray(T, ("time", "z_l", "xh"), coords={"z_l": np.array([1.0, 2.0])}),
                           ml.DataArray(T, ("time", "z_l", "xh")), eos="teos10")
