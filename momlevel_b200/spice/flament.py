""" flament.py -- P. Flament (2002) spiciness on the GPU.

Same call signature as ``momlevel.spice.flament.spice`` (src/momlevel/spice/flament.py:43-95).
"""

import numpy as np
import torch

from .. import core

__all__ = ["spice"]


def spice(thetao, so):
    """Sea water spiciness, same shape as the inputs (flament.py:43-95)."""
    if isinstance(thetao, torch.Tensor) and thetao.is_cuda:
        return core.flament_spice(thetao, core.to_device(so))
    # flament.py:68-70: python scalars become 1-element arrays
    thetao = np.array([float(thetao)]) if isinstance(thetao, (float, int)) else np.asarray(thetao)
    so = np.array([float(so)]) if isinstance(so, (float, int)) else np.asarray(so)
    assert thetao.shape == so.shape, "thetao and so must have the same shape"  # flament.py:75
    if thetao.size == 0:
        return np.empty(thetao.shape, dtype=np.float64)
    return core.flament_spice(np.ascontiguousarray(thetao), np.ascontiguousarray(so)).cpu().numpy()
