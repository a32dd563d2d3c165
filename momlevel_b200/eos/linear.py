""" linear.py -- linear equation of state on the GPU.

Same call signatures as ``momlevel.eos.linear`` (src/momlevel/eos/linear.py:26-162).
Pressure is accepted and ignored, as in the reference; the derivatives are constants.
"""

from ._dispatch import evaluate as _evaluate

__all__ = ["density", "drho_dtemp", "drho_dsal", "alpha", "beta"]

# linear.py:17-23
RHO_REF = 1035.0
RHO_T0_S0 = 1000.0
DRHO_DT = -0.2
DRHO_DS = 0.8


def density(T, S, p=None, rho_ref=None):
    """In-situ density in kg m-3 (linear.py:26-58)."""
    rho = _evaluate("linear", "density", T, S, None)
    # linear.py:55: an optional reference density shifts the constant term
    return rho if rho_ref is None else rho - rho_ref


def drho_dtemp(T=None, S=None, p=None):
    """linear.py:61-85."""
    return DRHO_DT


def drho_dsal(T=None, S=None, p=None):
    """linear.py:88-110."""
    return DRHO_DS


def alpha(T, S, p):
    """Thermal expansion coefficient (linear.py:113-136)."""
    return _evaluate("linear", "alpha", T, S, None)


def beta(T, S, p):
    """Haline contraction coefficient (linear.py:139-162)."""
    return _evaluate("linear", "beta", T, S, None)
