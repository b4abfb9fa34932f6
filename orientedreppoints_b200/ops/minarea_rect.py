"""Host-side mirror of mmdet/ops/minarearect/minarea_rect.py:4-7 over liborp_b200.so."""
import torch

from .. import _lib


def minaerarect(pred, *, scale=1.0, center=None, return_hull_map=False):
    """pred: cuda float32 [N,18] (x0,y0,...,x8,y8) -> cuda float32 [N,8].

    Reference: `minarearect.minareabbox(pred).view(-1, 8)`; empty input gives an empty CPU float
    tensor there (minarearect_cuda.cpp:7-8) - kept.  Keyword extensions: the fused affine
    `rect*scale + center` (orientedreppoints_head.py:748-749) and the hull index map
    (points_to_convex_ind, minarearect_kernel.cu:330-340).
    """
    if not pred.is_cuda:
        raise RuntimeError("ex_boxes must be a CUDA tensor")   # AT_CHECK in minarearect_cuda.cpp:6
    if pred.numel() == 0:
        out = torch.empty((0, 8), dtype=torch.float32)
        return (out, torch.empty((0, 9), dtype=torch.int32)) if return_hull_map else out
    p = pred
    if p.dtype != torch.float32 or not p.is_contiguous():
        p = p.float().contiguous()
    p = p.view(-1, 18)
    n = p.shape[0]
    out = torch.empty((n, 8), dtype=torch.float32, device=p.device)
    hmap = torch.empty((n, 9), dtype=torch.int32, device=p.device) if return_hull_map else None
    c = None
    if center is not None:
        c = center.to(device=p.device, dtype=torch.float32).contiguous().view(n, 2)
    with torch.cuda.device(p.device):
        rc = _lib.lib().orp_minarearect(_lib.ptr(p), n, _lib.ptr(out), _lib.ptr(hmap), float(scale), _lib.ptr(c),
                                        _lib.current_stream_ptr())
    _lib.check(rc, "orp_minarearect")
    return (out, hmap) if return_hull_map else out
