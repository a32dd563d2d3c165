"""Operator-level entry: numpy-broadcasting ``f(T, S, p)`` evaluated by the CUDA library.

``momlevel.eos.<name>.<func>`` are plain numpy functions (wright.py:23, linear.py:26) that
``xr.apply_ufunc`` applies to already-broadcast arrays (derived.py:624-630).  Here numpy
(or python scalar) arguments come back as numpy fp64; CUDA tensors come back as CUDA tensors
without leaving the device.
"""

import numpy as np
import torch

from .. import core


def evaluate(eos, func, T, S, p):
    on_device = any(isinstance(x, torch.Tensor) and x.is_cuda for x in (T, S, p))
    if on_device:
        T, S = torch.broadcast_tensors(core.to_device(T), core.to_device(S))
        if p is not None:
            p = core.to_device(p, torch.float64)
            if p.numel() != 1:
                T, S, p = torch.broadcast_tensors(T, S, p)
        return core.eos_eval(eos, func, T.contiguous(), S.contiguous(), None if p is None else p.contiguous())
    arrs = [np.asarray(x) for x in (T, S)]
    pa = None if p is None else np.asarray(p, dtype=np.float64)
    if pa is not None and pa.size != 1:
        Tb, Sb, pa = np.broadcast_arrays(arrs[0], arrs[1], pa)
    else:
        Tb, Sb = np.broadcast_arrays(*arrs)
    shape = Tb.shape
    keep32 = Tb.dtype == np.float32 and Sb.dtype == np.float32
    dt = np.float32 if keep32 else np.float64
    Tb = np.ascontiguousarray(Tb, dtype=dt).reshape(-1)
    Sb = np.ascontiguousarray(Sb, dtype=dt).reshape(-1)
    if pa is not None:
        pa = np.ascontiguousarray(pa).reshape(-1)
    if Tb.size == 0:
        return np.empty(shape, dtype=np.float64)
    out = core.eos_eval(eos, func, Tb, Sb, pa).cpu().numpy().reshape(shape)
    return out if out.shape else out[()]
