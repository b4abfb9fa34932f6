"""Mirror of the SWIG module DOTA_devkit/polyiou (polyiou.i:1-19, polyiou.h:9) over liborp_b200.so.

`VectorDouble(iterable)` and `iou_poly(p, q) -> float` keep their names; the arithmetic is the
reference's fp64 algorithm (polyiou.cpp:108-128) executed on the GPU, so results are bit-identical to
the compiled reference.  `iou_poly_pairs` is the batched form callers should prefer (one launch).
"""
import numpy as np
import torch

from .. import _lib


class VectorDouble(list):
    """std::vector<double> stand-in: a list of floats."""

    def __init__(self, it=()):
        super().__init__(float(v) for v in it)


def iou_poly_pairs(p, q, device=None):
    """p, q: array-like [N,8] -> numpy float64 [N] (aligned pairs)."""
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    pt = torch.as_tensor(np.asarray(p, dtype=np.float64).reshape(-1, 8)).to(dev).contiguous()
    qt = torch.as_tensor(np.asarray(q, dtype=np.float64).reshape(-1, 8)).to(dev).contiguous()
    n = pt.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=dev)
    if n:
        with torch.cuda.device(dev):
            rc = _lib.lib().orp_iou_poly_f64_pairs(_lib.ptr(pt), _lib.ptr(qt), n, _lib.ptr(out),
                                                   _lib.current_stream_ptr())
        _lib.check(rc, "orp_iou_poly_f64_pairs")
    return out.cpu().numpy()


def iou_poly(p, q):
    """double iou_poly(std::vector<double> p, std::vector<double> q) - reads p[0..7], q[0..7]."""
    return float(iou_poly_pairs(list(p)[:8], list(q)[:8])[0])
