"""Host-side mirror of mmdet/ops/nms/nms_wrapper.py:177-199 (`rnms`) over liborp_b200.so.

Same name, arguments and error behaviour as the reference:
  * dets: Tensor [N,9] (x1,y1,...,x4,y4,score) or ndarray; returns (dets[inds], inds)
  * inds: int64 on dets.device, ASCENDING original index (rnms_kernel.cu:261-264)
  * CPU tensor -> TypeError('dets must be cuda tensor') (nms_wrapper.py:197)
  * empty input -> empty int64 tensor (nms_wrapper.py:191-192)
Extensions (keyword-only): `segments` restricts suppression to boxes of the same segment (what
multiclass_rnms emulates with coordinate offsets); `mode` selects the IoU arithmetic: 'exact64' (default) decides
every pair as the reference's fp64 polyiou would - equal to the reference's fp32 rnms wherever that is numerically
sound, and correct where it is not (large coordinates); 'compat32' reproduces the reference's fp32
arithmetic bit for bit, unsound cases included.
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


def soft_rnms(dets, iou_thr, method='linear', sigma=0.5, min_score=1e-3):
    """Mirror of mmdet/ops/nms/nms_wrapper.py:120-175 `soft_rnms` -> (new_dets [K,9], inds [K]), same container type and
    dtype as the input.  The reference runs this on the CPU only (rnms_cpu.cpp:165-320: selection by repeated arg-max with
    in-place swaps, score decay `1 - iou` above the threshold / `exp(-iou^2 / sigma)` / hard 0, boxes whose score falls
    below min_score are dropped by swapping with the last box).  Here the N x N fp32 IoUs - the O(N^2) polygon clipping,
    in the reference's own fp32 arithmetic (ORP_NMS_COMPAT32) - come from ONE device launch; the order-dependent
    selection then runs on the host over that matrix, reproducing the reference's array permutations step for step."""
    if isinstance(dets, torch.Tensor):
        is_tensor, d_np = True, dets.detach().cpu().numpy()
    elif isinstance(dets, np.ndarray):
        is_tensor, d_np = False, dets
    else:
        raise TypeError('dets must be either a Tensor or numpy array, but got {}'.format(type(dets)))
    method_codes = {'linear': 1, 'gaussian': 2, 'original': 0}
    if method not in method_codes:
        raise ValueError('Invalid method for SoftNMS: {}'.format(method))
    code = method_codes[method]
    n = d_np.shape[0]
    work_dtype = np.float64 if d_np.dtype == np.float64 else np.float32
    if n == 0:
        new_dets, inds = np.zeros((0, 9), d_np.dtype), np.zeros((0,), np.int64)
    else:
        if n > 20000:
            raise ValueError("soft_rnms: %d boxes need a %d MB IoU matrix; use rnms for sets this large" % (n, n * n * 4 >> 20))
        from .box_iou_rotated import quad_iou_matrix
        dev = dets.device if (is_tensor and dets.is_cuda) else torch.device('cuda', torch.cuda.current_device())
        q = torch.from_numpy(np.ascontiguousarray(d_np[:, :8], dtype=np.float32)).to(dev)
        iou = quad_iou_matrix(q, q, mode="compat32").cpu().numpy()            # float32 [n, n], rotate_iou of rnms_cpu.cpp
        scores = d_np[:, 8].astype(work_dtype).copy()
        idx = np.arange(n)                                                     # which original box sits at each position
        thr, sg, ms = np.float32(iou_thr), np.float32(sigma), np.float32(min_score)
        nd = n
        i = 0
        while i < nd:
            # the best remaining box moves to position i (first maximum, as the strict `<` of the reference picks)
            m = i + int(np.argmax(scores[i:nd]))
            scores[[i, m]] = scores[[m, i]]
            idx[[i, m]] = idx[[m, i]]
            if i + 1 < nd:
                ovr = iou[idx[i], idx[i + 1:nd]]                               # float32
                if code == 1:
                    w = np.where(ovr > thr, (work_dtype(1) - ovr.astype(work_dtype)), work_dtype(1))
                elif code == 2:
                    w = np.exp(-(ovr * ovr) / sg).astype(work_dtype)
                else:
                    w = np.where(ovr > thr, work_dtype(0), work_dtype(1))
                scores[i + 1:nd] = w * scores[i + 1:nd]
                # drop boxes below min_score exactly as the reference's scan does: a dead position takes the last box
                # (and is examined again), i.e. dead front slots are refilled with live boxes from the tail, tail first
                dead = scores[i + 1:nd] < ms
                if dead.any():
                    live = ~dead
                    keep_n = int(live.sum())
                    pos = np.arange(i + 1, nd)
                    front_dead = pos[:keep_n][dead[:keep_n]]
                    tail_live = pos[keep_n:][live[keep_n:]][::-1]
                    scores[front_dead] = scores[tail_live]
                    idx[front_dead] = idx[tail_live]
                    nd = i + 1 + keep_n
            i += 1
        new_dets = np.concatenate([d_np[idx[:nd], :8], scores[:nd, None].astype(d_np.dtype)], 1).astype(d_np.dtype)
        inds = idx[:nd].astype(np.int64)
    if is_tensor:
        return (torch.from_numpy(new_dets).to(device=dets.device, dtype=dets.dtype),
                torch.from_numpy(inds).to(device=dets.device, dtype=torch.long))
    return new_dets.astype(d_np.dtype), inds
