"""Shared plumbing for the converter modules: host numpy arrays -> device float64 -> exact CUDA chain -> host."""

import numpy as np
import torch

from .. import engine
from .._native import lib, check


def _to_dev(a, dev):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return torch.from_numpy(arr.reshape(-1)).to(dev), arr.shape


def _back(t, shape, scalar):
    out = t.cpu().numpy().reshape(shape)
    return float(out.reshape(-1)[0]) if scalar else out


def run3(fn_name, a, b, c, *params):
    """a,b,c -> 3 outputs through lib.<fn_name>(ctx, a, b, c, n, *params, o1, o2, o3, stream)."""
    ctx, dev = engine.default_context()
    scalar = np.isscalar(a) and np.isscalar(b) and np.isscalar(c)
    a, b, c = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64),
                                  np.asarray(c, dtype=np.float64))
    da, shape = _to_dev(a, dev)
    db, _ = _to_dev(b, dev)
    dc, _ = _to_dev(c, dev)
    o = [torch.empty_like(da) for _ in range(3)]
    check(getattr(lib, fn_name)(ctx.handle, engine._ptr(da), engine._ptr(db), engine._ptr(dc), da.numel(), *params,
                                engine._ptr(o[0]), engine._ptr(o[1]), engine._ptr(o[2]), engine._stream(dev)), fn_name)
    return tuple(_back(t, shape, scalar) for t in o)


def run2(fn_name, a, b, *params):
    ctx, dev = engine.default_context()
    scalar = np.isscalar(a) and np.isscalar(b)
    a, b = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
    da, shape = _to_dev(a, dev)
    db, _ = _to_dev(b, dev)
    o = [torch.empty_like(da) for _ in range(2)]
    check(getattr(lib, fn_name)(ctx.handle, engine._ptr(da), engine._ptr(db), da.numel(), *params,
                                engine._ptr(o[0]), engine._ptr(o[1]), engine._stream(dev)), fn_name)
    return tuple(_back(t, shape, scalar) for t in o)
