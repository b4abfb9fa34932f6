"""Host-side mirror of mmdet/ops/nms/nms_wrapper.py:177-199 (`rnms`) over liborp_b200.so.

Same name, arguments and error behaviour as the reference:
  * dets: Tensor [N,9] (x1,y1,...,x4,y4,score) or ndarray; returns (dets[inds], inds)
  * inds: int64 on dets.device, ASCENDING original index (rnms_kernel.cu:261-264)
  * CPU tensor -> TypeError('dets must be cuda tensor') (nms_wrapper.py:197)
  * empty input -> empty int64 tensor (nms_wrapper.py:191-192)
Extension (keyword-only, default = reference behaviour): `segments` restricts suppression to boxes
of the same segment (what multiclass_rnms emulates with coordinate offsets), `mode` selects the IoU
arithmetic ('exact64' default, 'compat32' = bit-faithful reference fp32).
"""
import numpy as np
import torch

from .. import _lib

_MODES = {"exact64": _lib.ORP_NMS_EXACT64, "compat32": _lib.ORP_NMS_COMPAT32}


def rnms_indices(dets_th, iou_thr, segments=None, mode="exact64", union_mode=_lib.ORP_UNION_NAN_KEEPS,
                 order=_lib.ORP_ORDER_INDEX_ASC, return_count_tensor=False):
    """Device-side NMS: returns kept indices (int64, on device). One host sync (the count)."""
    if not dets_th.is_cuda:
        raise TypeError('dets must be cuda tensor')
    n = dets_th.shape[0]
    if n == 0:
        return dets_th.new_zeros(0, dtype=torch.long)
    d = dets_th
    if d.dtype != torch.float32 or not d.is_contiguous():
        d = d.float().contiguous()
    seg = None
    if segments is not None:
        seg = segments.to(device=d.device, dtype=torch.int32).contiguous()
    keep = torch.empty(n, dtype=torch.int64, device=d.device)
    cnt = torch.empty(1, dtype=torch.int32, device=d.device)
    with torch.cuda.device(d.device):
        rc = _lib.lib().orp_rnms(_lib.ptr(d), _lib.ptr(seg), n, float(iou_thr), _MODES[mode], union_mode, order,
                                 _lib.ptr(keep), _lib.ptr(cnt), _lib.current_stream_ptr())
    _lib.check(rc, "orp_rnms")
    if return_count_tensor:
        return keep, cnt
    k = int(cnt.item())
    return keep[:k]


def rnms(dets, iou_thr, device_id=None, *, segments=None, mode="exact64"):
    # convert dets (tensor or numpy array) to tensor  (nms_wrapper.py:178-189)
    if isinstance(dets, torch.Tensor):
        dets_th = dets
    elif isinstance(dets, np.ndarray):
        device = 'cpu' if device_id is None else 'cuda:{}'.format(device_id)
        dets_th = torch.from_numpy(dets).to(device)
    else:
        raise TypeError(
            'dets must be either a Tensor or numpy array, but got {}'.format(type(dets)))
    if dets_th.shape[0] == 0:
        inds = dets_th.new_zeros(0, dtype=torch.long)
    else:
        if dets_th.is_cuda:
            inds = rnms_indices(dets_th, iou_thr, segments=segments, mode=mode)
        else:
            raise TypeError('dets must be cuda tensor')
    if isinstance(dets, np.ndarray):
        # the reference indexes the ndarray with a CUDA tensor here (latent bug);
        # we return the rows the caller evidently wanted
        return dets[inds.cpu().numpy(), :], inds
    return dets[inds, :], inds
