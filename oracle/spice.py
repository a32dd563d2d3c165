"""oracle.spice -- numpy restatement of Flament (2002) spiciness (TEST INFRASTRUCTURE).

Follows ``src/momlevel/spice/flament.py:7-95``:

    pi(T, S) = sum_{i=0..5} sum_{j=0..4} b[i][j] * T**i * (S - 35)**j

The reference materialises N x 6 and N x 5 power tables and an N x 6 x 5 product
(flament.py:82-90).  Here the 29 non-zero terms are accumulated one at a time, which
needs no temporaries larger than the input; the two orders agree to a few ulp
(``tests/test_oracle.py`` bounds the gap against ``tests/golden/spice.npz``).
"""

import numpy as np

__all__ = ["FLAMENT_B", "flament_spice"]

# flament.py:7-40 -- b[i][j], i = power of T, j = power of (S - 35)
FLAMENT_B = np.array(
    [
        [0.0, 7.7442e-1, -5.85e-3, -9.84e-4, -2.06e-4],
        [5.1655e-2, 2.034e-3, -2.742e-4, -8.5e-6, 1.36e-5],
        [6.64783e-3, -2.4681e-4, -1.428e-5, 3.337e-5, 7.894e-6],
        [-5.4023e-5, 7.326e-6, 7.0036e-6, -3.0412e-6, -1.0853e-6],
        [3.949e-7, -3.029e-8, -3.8209e-7, 1.0012e-7, 4.7133e-8],
        [-6.36e-10, -1.309e-9, 6.048e-9, -1.1409e-9, -6.676e-10],
    ]
)


def flament_spice(thetao, so):
    """Spiciness, same shape as the inputs (flament.py:43-95)."""
    # flament.py:68-70: python scalars become 1-element arrays
    if isinstance(thetao, (float, int)):
        thetao = np.array([float(thetao)])
    if isinstance(so, (float, int)):
        so = np.array([float(so)])
    thetao = np.asarray(thetao, dtype=np.float64)
    so = np.asarray(so, dtype=np.float64)
    # flament.py:75
    assert thetao.shape == so.shape, "thetao and so must have the same shape"

    ds = so - 35.0
    tpow = [thetao**i for i in range(6)]
    spow = [ds**j for j in range(5)]
    out = np.zeros_like(thetao)
    for i in range(6):
        for j in range(5):
            out = out + (spow[j] * tpow[i]) * FLAMENT_B[i, j]
    return out
