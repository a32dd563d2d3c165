"""oracle.eos -- numpy restatement of the two equations of state (TEST INFRASTRUCTURE).

Follows ``src/momlevel/eos/wright.py:6-165`` and ``src/momlevel/eos/linear.py:17-162``.
The floating-point operation order of each expression matches the reference, so on
fp64 inputs the results agree with the reference bit for bit (checked in
``tests/test_oracle.py`` against ``tests/golden/eos_*.npz``).

Wright (1997), J. Atmos. Ocean. Tech. 14, 735-740, reduced-range fit:

    rho = (p + p0) / (lam + al0 * (p + p0))
    al0 = a0 + a1 T + a2 S
    p0  = b0 + b4 S + T (b1 + T (b2 + b3 T) + b5 S)
    lam = c0 + c4 S + T (c1 + T (c2 + c3 T) + c5 S)
"""

import numpy as np

__all__ = [
    "WRIGHT_A",
    "WRIGHT_B",
    "WRIGHT_C",
    "LINEAR",
    "wright_density",
    "wright_drho_dtemp",
    "wright_drho_dsal",
    "wright_alpha",
    "wright_beta",
    "linear_density",
    "linear_drho_dtemp",
    "linear_drho_dsal",
    "linear_alpha",
    "linear_beta",
    "density",
]

# wright.py:6-20 (published fit coefficients)
WRIGHT_A = (7.057924e-4, 3.480336e-7, -1.112733e-7)
WRIGHT_B = (5.790749e8, 3.516535e6, -4.002714e4, 2.084372e2, 5.944068e5, -9.643486e3)
WRIGHT_C = (1.704853e5, 7.904722e2, -7.984422, 5.140652e-2, -2.302158e2, -3.079464)

# linear.py:17-23
LINEAR = {"rho_ref": 1035.0, "rho_t0_s0": 1000.0, "drho_dt": -0.2, "drho_ds": 0.8}


def _wright_terms(T, S):
    """al0, p0, lam exactly as wright.py:44-46 evaluates them."""
    a0, a1, a2 = WRIGHT_A
    b0, b1, b2, b3, b4, b5 = WRIGHT_B
    c0, c1, c2, c3, c4, c5 = WRIGHT_C
    al0 = a0 + a1 * T + a2 * S
    p0 = b0 + b4 * S + T * (b1 + T * (b2 + b3 * T) + b5 * S)
    lam = c0 + c4 * S + T * (c1 + T * (c2 + c3 * T) + c5 * S)
    return al0, p0, lam


def wright_density(T, S, p):
    """In-situ density, kg m-3 (wright.py:23-50)."""
    al0, p0, lam = _wright_terms(T, S)
    inv_denom = 1.0 / (lam + al0 * (p + p0))
    return (p + p0) * inv_denom


def wright_drho_dtemp(T, S, p):
    """d rho / d T (wright.py:53-85)."""
    a0, a1, a2 = WRIGHT_A
    b0, b1, b2, b3, b4, b5 = WRIGHT_B
    c0, c1, c2, c3, c4, c5 = WRIGHT_C
    al0, p0, lam = _wright_terms(T, S)
    inv2 = 1.0 / (lam + al0 * (p + p0))
    inv2 = inv2 * inv2
    return inv2 * (
        lam * (b1 + T * (2.0 * b2 + 3.0 * b3 * T) + b5 * S)
        - (p + p0) * ((p + p0) * a1 + (c1 + T * (c2 * 2.0 + c3 * 3.0 * T) + c5 * S))
    )


def wright_drho_dsal(T, S, p):
    """d rho / d S (wright.py:88-119)."""
    a0, a1, a2 = WRIGHT_A
    b0, b1, b2, b3, b4, b5 = WRIGHT_B
    c0, c1, c2, c3, c4, c5 = WRIGHT_C
    al0, p0, lam = _wright_terms(T, S)
    inv2 = 1.0 / (lam + al0 * (p + p0))
    inv2 = inv2 * inv2
    return inv2 * (lam * (b4 + b5 * T) - (p + p0) * ((p + p0) * a2 + (c4 + c5 * T)))


def wright_alpha(T, S, p):
    """Thermal expansion coefficient (wright.py:122-142)."""
    return -1.0 * (wright_drho_dtemp(T, S, p) / wright_density(T, S, p))


def wright_beta(T, S, p):
    """Haline contraction coefficient (wright.py:145-165)."""
    return wright_drho_dsal(T, S, p) / wright_density(T, S, p)


def linear_density(T, S, p=None, rho_ref=None):
    """Linear EOS density (linear.py:26-58); ``p`` is ignored by design."""
    base = LINEAR["rho_t0_s0"] if rho_ref is None else (LINEAR["rho_t0_s0"] - rho_ref)
    return base + ((LINEAR["drho_dt"] * T) + (LINEAR["drho_ds"] * S))


def linear_drho_dtemp(T=None, S=None, p=None):
    """Constant (linear.py:61-85)."""
    return LINEAR["drho_dt"]


def linear_drho_dsal(T=None, S=None, p=None):
    """Constant (linear.py:88-110)."""
    return LINEAR["drho_ds"]


def linear_alpha(T, S, p):
    """linear.py:113-136."""
    return -1.0 * (np.full_like(T, fill_value=LINEAR["drho_dt"]) / linear_density(T, S, p))


def linear_beta(T, S, p):
    """linear.py:139-162."""
    return np.full_like(T, fill_value=LINEAR["drho_ds"]) / linear_density(T, S, p)


_DENSITY = {"wright": wright_density, "linear": linear_density}


def density(eos, T, S, p):
    """Name dispatch the way ``util.eos_func_from_str`` does it (util.py:227-249)."""
    assert isinstance(eos, str), "Expecting string for equation of state"
    key = eos.lower()
    if key not in _DENSITY:
        raise ValueError(f"Unknown equation of state: {key}")
    return _DENSITY[key](T, S, p)


def use_reference_modules(root):
    """Swap the reference's OWN numpy kernels in under the oracle's driver, when an install of it is at hand.

    ``root`` is a directory that holds the reference package (``<root>/momlevel/eos/wright.py`` after
    ``pip install --target baseline/_ref``, or ``<root>/src/momlevel/...`` for a source tree).  The
    package itself cannot be imported without xarray, but ``eos/wright.py`` and ``eos/linear.py`` depend on
    numpy alone and load by file path.  Returns the list of files now in use (empty: nothing found, the
    restatement above stays).  Used by ``bench.py --impl reference`` and its ``cpu_baseline`` leg.
    """
    import importlib.util
    import pathlib

    used = []
    for base in (pathlib.Path(root) / "momlevel", pathlib.Path(root) / "src" / "momlevel"):
        for name in ("wright", "linear"):
            path = base / "eos" / f"{name}.py"
            if not path.exists() or name in [pathlib.Path(u).stem for u in used]:
                continue
            try:
                spec = importlib.util.spec_from_file_location(f"_momlevel_ref_{name}", str(path))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _DENSITY[name] = mod.density
                used.append(str(path))
            except Exception:  # noqa: BLE001 -- a broken install leaves the restatement in place
                continue
    return used
