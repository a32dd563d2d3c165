"""dynamic.py -- inverse barometer height (mirrors ``src/momlevel/dynamic.py:8-41``).

A 2-D consumer of the elementwise EOS kernel (``ml_eos_eval``); listed as a "next" row of the
hot path.
"""

import numpy as np

from .derived import calc_rho
from .labeled import DataArray

__all__ = ["inverse_barometer"]


def inverse_barometer(tos, sos, pso, gravity=9.8, equation_of_state="Wright"):
    """Inverse barometer height in m: ``pso * (-1 / (rho(tos, sos, pso) * gravity))`` (dynamic.py:34-36)."""
    rho_conv = calc_rho(tos, sos, pso, eos=equation_of_state)
    if isinstance(pso, DataArray):
        ibh = pso * (-1.0 / (rho_conv * gravity))
    else:
        ibh = (-1.0 / (rho_conv * gravity)) * float(np.asarray(pso))
    ibh.name = "ibh"
    ibh.attrs = {"long_name": "Inverse Barometer Height", "units": "m"}
    return ibh
