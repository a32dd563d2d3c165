"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

The reference's numpy-level modules (``eos/wright.py``, ``eos/linear.py``,
``spice/flament.py``) depend on numpy alone, so they are imported by file path from
``/root/reference`` -- xarray is not installed here, which rules out importing the
package itself -- evaluated on seeded inputs, and the inputs and outputs are stored.
The GPU box has no ``/root/reference``; tests read only the committed ``.npz`` files.

    python tests/golden/make_golden.py [/root/reference]
"""

import importlib.util
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent


def _load(root, rel, name):
    spec = importlib.util.spec_from_file_location(name, str(pathlib.Path(root) / rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(root="/root/reference"):
    wright = _load(root, "src/momlevel/eos/wright.py", "_ref_wright")
    linear = _load(root, "src/momlevel/eos/linear.py", "_ref_linear")
    flament = _load(root, "src/momlevel/spice/flament.py", "_ref_flament")

    rng = np.random.default_rng(20261018)
    n = 2048
    # fp32-representable T/S (what MOM6 writes), upcast to fp64 -- the parity definition
    T = rng.uniform(-2.0, 40.0, n).astype(np.float32).astype(np.float64)
    S = rng.uniform(0.0, 42.0, n).astype(np.float32).astype(np.float64)
    p = rng.uniform(1.0e5, 7.0e7, n)
    # a few awkward points: exact zeros, freezing point, NaN (land)
    T[:4] = [0.0, -1.9, 0.0, np.nan]
    S[:4] = [0.0, 35.0, 35.0, 35.0]
    T[4], S[4] = 10.0, np.nan

    np.savez_compressed(
        HERE / "eos_wright.npz",
        T=T,
        S=S,
        p=p,
        density=wright.density(T, S, p),
        drho_dtemp=wright.drho_dtemp(T, S, p),
        drho_dsal=wright.drho_dsal(T, S, p),
        alpha=wright.alpha(T, S, p),
        beta=wright.beta(T, S, p),
    )
    np.savez_compressed(
        HERE / "eos_linear.npz",
        T=T,
        S=S,
        p=p,
        density=linear.density(T, S, p),
        density_rho_ref=linear.density(T, S, p, rho_ref=1035.0),
        alpha=linear.alpha(T, S, p),
        beta=linear.beta(T, S, p),
        drho_dtemp=np.float64(linear.drho_dtemp()),
        drho_dsal=np.float64(linear.drho_dsal()),
    )
    Ts = rng.uniform(-2.0, 32.0, (32, 64)).astype(np.float32).astype(np.float64)
    Ss = rng.uniform(30.0, 40.0, (32, 64)).astype(np.float32).astype(np.float64)
    Ts[0, 0] = np.nan
    np.savez_compressed(HERE / "spice.npz", T=Ts, S=Ss, spice=flament.spice(Ts, Ss))
    print("wrote", sorted(x.name for x in HERE.glob("*.npz")))


if __name__ == "__main__":
    main(*sys.argv[1:])
