""" wright.py -- Wright (1997) equation of state on the GPU.

Same call signatures as ``momlevel.eos.wright`` (src/momlevel/eos/wright.py:23-165):
elementwise ``f(T, S, p)`` with numpy broadcasting.  The arithmetic runs in
libmomlevel_b200 (``ml_eos_eval``, csrc/ml_common.cuh).
"""

from ._dispatch import evaluate as _evaluate

__all__ = ["density", "drho_dtemp", "drho_dsal", "alpha", "beta"]


def density(T, S, p):
    """In-situ density in kg m-3 (wright.py:23-50)."""
    return _evaluate("wright", "density", T, S, p)


def drho_dtemp(T, S, p):
    """Density derivative with respect to potential temperature (wright.py:53-85)."""
    return _evaluate("wright", "drho_dtemp", T, S, p)


def drho_dsal(T, S, p):
    """Density derivative with respect to salinity (wright.py:88-119)."""
    return _evaluate("wright", "drho_dsal", T, S, p)


def alpha(T, S, p):
    """Thermal expansion coefficient in degC-1 (wright.py:122-142)."""
    return _evaluate("wright", "alpha", T, S, p)


def beta(T, S, p):
    """Haline contraction coefficient in PSU-1 (wright.py:145-165)."""
    return _evaluate("wright", "beta", T, S, p)
