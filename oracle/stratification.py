"""oracle.stratification -- numpy restatement of the buoyancy-frequency diagnostics (TEST INFRASTRUCTURE).

Follows ``src/momlevel/derived.py:328-411`` (``calc_n2`` at cell centres; the ``interfaces``
branch needs xgcm and is out of scope) and ``src/momlevel/derived.py:30-71``
(``adjust_negative_n2``, Chelton et al. 1998).  ``DataArray.differentiate(z, edge_order=2)`` is
``numpy.gradient(values, z, axis, edge_order=2)``; ``ffill`` is a forward fill along z.

Pinned by ``tests/test_derived.py:14-18, 54-61`` of the reference (sums 0.00338354 and 0.12093286).
"""

import numpy as np

from . import eos as _eos

__all__ = ["calc_n2", "adjust_negative_n2", "calc_stability_angle", "calc_wave_speed"]


def adjust_negative_n2(n2, z_axis=1):
    """derived.py:30-71.  Note ``adjusted[0]`` indexes the array's FIRST axis (time for 4-D input)."""
    n2 = np.asarray(n2, dtype=np.float64)
    mask = np.where(np.isnan(n2), np.nan, 1.0)
    with np.errstate(invalid="ignore"):
        adjusted = np.where(n2 <= 0.0, np.nan, n2)
    adjusted[0] = np.where(np.isnan(adjusted[0]), 1.0e-8, adjusted[0])
    adjusted = np.moveaxis(adjusted, z_axis, 0).copy()
    for k in range(1, adjusted.shape[0]):
        adjusted[k] = np.where(np.isnan(adjusted[k]), adjusted[k - 1], adjusted[k])
    adjusted = np.moveaxis(adjusted, 0, z_axis)
    return adjusted * mask


def calc_n2(thetao, so, z_l, eos="Wright", gravity=-9.8, patm=101325.0, z_axis=1, adjust_negative=False):
    """derived.py:391-411: ``g * (alpha * dT/dz - beta * dS/dz)`` with locally referenced pressure."""
    assert eos.lower() in ("wright", "linear")
    thetao = np.asarray(thetao, dtype=np.float64)
    so = np.asarray(so, dtype=np.float64)
    z_l = np.asarray(z_l, dtype=np.float64)
    shape = [1] * thetao.ndim
    shape[z_axis] = z_l.size
    pres = ((z_l * 1.0e4) + patm).reshape(shape)
    if eos.lower() == "wright":
        alpha, beta = _eos.wright_alpha(thetao, so, pres), _eos.wright_beta(thetao, so, pres)
    else:
        alpha, beta = _eos.linear_alpha(thetao, so, pres), _eos.linear_beta(thetao, so, pres)
    dtdz = np.gradient(thetao, z_l, axis=z_axis, edge_order=2)
    dsdz = np.gradient(so, z_l, axis=z_axis, edge_order=2)
    n2 = gravity * ((alpha * dtdz) - (beta * dsdz))
    return adjust_negative_n2(n2, z_axis=z_axis) if adjust_negative else n2


def calc_stability_angle(thetao, so, pres, z_l, eos="Wright", z_axis=1):
    """derived.py:714-766: Turner angle ``degrees(arctan((1 + R) / (1 - R)))``, ``R = beta dS/dz / (alpha dT/dz)``.

    ``pres`` is the caller's 1-D pressure over z (the reference's test passes ``z_l * 1e4``, without patm).
    """
    assert eos.lower() in ("wright", "linear")
    thetao = np.asarray(thetao, dtype=np.float64)
    so = np.asarray(so, dtype=np.float64)
    z_l = np.asarray(z_l, dtype=np.float64)
    shape = [1] * thetao.ndim
    shape[z_axis] = z_l.size
    pres = np.asarray(pres, dtype=np.float64).reshape(shape)
    if eos.lower() == "wright":
        alpha, beta = _eos.wright_alpha(thetao, so, pres), _eos.wright_beta(thetao, so, pres)
    else:
        alpha, beta = _eos.linear_alpha(thetao, so, pres), _eos.linear_beta(thetao, so, pres)
    dtdz = np.gradient(thetao, z_l, axis=z_axis, edge_order=2)
    dsdz = np.gradient(so, z_l, axis=z_axis, edge_order=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        r_rho = (beta * dsdz) / (alpha * dtdz)
        return np.degrees(np.arctan((1 + r_rho) / (1 - r_rho)))


def calc_wave_speed(n2, dz, z_axis=1):
    """derived.py:798-828: first-baroclinic-mode gravity wave speed ``sum_z(sqrt(N2adj) dz) / pi``.

    ``n2`` is ``[t][z][y][x]``, ``dz`` is ``[z][y][x]``.  Returns ``(c1 [t][y][x], broadcast)`` where
    ``broadcast`` is what the reference actually returns: ``xr.where(n2[0].isnull(), nan, c1)`` pairs the
    FIRST TIME STEP of n2 (dims z,y,x) with c1 (dims t,y,x), so the result has dims ``(z, y, x, t)``.
    """
    n2 = np.asarray(n2, dtype=np.float64)
    assert n2.ndim == 4 and z_axis == 1
    adj = adjust_negative_n2(n2, z_axis=z_axis)
    with np.errstate(invalid="ignore"):
        c1 = np.nansum(np.sqrt(adj) * np.asarray(dz, dtype=np.float64)[None], axis=1) / np.pi
    # dims (z,y,x) against (t,y,x) -> (z,y,x,t)
    broadcast = np.where(np.isnan(n2[0])[..., None], np.nan, np.moveaxis(c1, 0, -1)[None])
    return c1, broadcast
