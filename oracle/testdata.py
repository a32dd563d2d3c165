"""oracle.testdata -- numpy-only clone of ``momlevel.test_data`` (TEST INFRASTRUCTURE).

The reference builds its 5x5x5 synthetic MOM6 dataset with xarray
(``src/momlevel/test_data/__init__.py:16-105``); the random draws themselves are plain
``numpy.random.default_rng(seed)`` calls, restated here so that every reference
known-answer test can be re-run without xarray:

* ``thetao``   ~ N(15, 5),   ``so`` ~ N(35, 1.5), ``volcello`` ~ N(1000, 100), each from a
  *fresh* generator with the same seed, shape ``(nt, 5, 5, 5)`` (``__init__.py:66-103``)
* ``areacello`` ~ N(100, 10) normalised to 3.6111092e14 m2 (``tripolar/horizontal.py:110-115``)
* ``z_i``, ``z_l`` fixed, ``deptho`` rows ~ U(0, z_i[k+1]) (``tripolar/vertical.py:37-68``)
* the dz fixture (``__init__.py:108-140``)
"""

import numpy as np

__all__ = ["generate_test_data", "generate_test_data_dz"]


def generate_test_data(ntimes=5, seed=123):
    """Dict of arrays equivalent to ``generate_test_data(seed=seed)`` (``nyears=0``).

    ``ntimes`` other than 5 corresponds to the ``nyears >= 1`` monthly datasets
    (``ntimes = 12 * nyears``); the time axis itself is not reproduced.
    """
    z_i = np.array([0.0, 5.0, 15.0, 185.0, 1815.0, 6185.0])
    z_l = np.array([2.5, 10.0, 100.0, 1000.0, 4000.0])
    deptho = np.array([np.random.default_rng(seed).uniform(0.0, hi, 5) for hi in z_i[1:]])
    area = np.random.default_rng(seed).normal(100.0, 10.0, (5, 5))
    area = area / area.sum()
    shape = (ntimes, 5, 5, 5)
    return {
        "time": np.arange(1.0, ntimes + 1.0),
        "xh": np.arange(1.0, 6.0),
        "yh": np.arange(1.0, 6.0),
        "z_l": z_l,
        "z_i": z_i,
        "deptho": deptho,
        "areacello": area * 3.6111092e14,
        "thetao": np.random.default_rng(seed).normal(15.0, 5.0, shape),
        "so": np.random.default_rng(seed).normal(35.0, 1.5, shape),
        "volcello": np.random.default_rng(seed).normal(1000.0, 100.0, shape),
    }


def generate_test_data_dz(seed=123):
    """``generate_test_data_dz`` (``__init__.py:108-140``)."""
    deptho = np.random.default_rng(seed).uniform(0.0, 100.0, (5, 5))
    deptho[2, 2] = np.nan
    deptho[2, 3] = np.nan
    z_i = np.array([0.0, 5.0, 10.0, 20.0, 50.0, 100.0])
    z_l = (z_i[1:] + z_i[:-1]) / 2.0
    return {"deptho": deptho, "z_l": z_l, "z_i": z_i, "xh": np.arange(1, 6), "yh": np.arange(10, 60, 10)}
