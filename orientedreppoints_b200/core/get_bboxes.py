"""Mirror of OrientedRepPointsHead.get_bboxes / get_bboxes_single
(mmdet/models/anchor_heads/orientedreppoints_head.py:673-779) on NHWC head outputs.

Per level: sigmoid, max-over-class top-k(nms_pre), (dy,dx)->(x,y), minaerarect with the
`*stride + centre` affine fused into the kernel, reppoints likewise; levels concatenated; multiclass_rnms.
"""
import torch

from ..ops import minaerarect
from .bbox_nms import multiclass_rnms


def grid_points(h, w, stride, device):
    """PointGenerator.grid_points (mmdet/core/anchor/point_generator.py:14-22): (x*s, y*s), x fastest."""
    xs = torch.arange(0, w, device=device, dtype=torch.float32) * stride
    ys = torch.arange(0, h, device=device, dtype=torch.float32) * stride
    return torch.stack([xs.repeat(h), ys.view(-1, 1).repeat(1, w).view(-1)], dim=1)


def get_bboxes_single(cls_scores, points_preds, strides, scale_factor, cfg, rescale=False, nms=True):
    """cls_scores[l]: [H,W,15] logits, points_preds[l]: [H,W,18] (dy,dx interleaved, stride units) of ONE image."""
    mlvl_bboxes, mlvl_scores, mlvl_reppoints = [], [], []
    nms_pre = cfg.get('nms_pre', -1)
    for cls_score, points_pred, stride in zip(cls_scores, points_preds, strides):
        h, w, c = cls_score.shape
        scores = cls_score.reshape(-1, c).sigmoid()
        points_pred = points_pred.reshape(-1, 18)
        points = grid_points(h, w, stride, cls_score.device)
        if nms_pre > 0 and scores.shape[0] > nms_pre:
            max_scores, _ = scores.max(dim=1)
            # torch.topk in the reference; ties (implementation-defined there) -> lower index first
            _, order = max_scores.sort(descending=True, stable=True)
            topk_inds = order[:nms_pre]
            points = points[topk_inds, :]
            points_pred = points_pred[topk_inds, :]
            scores = scores[topk_inds, :]
        pts = points_pred.reshape(-1, 9, 2)
        pts_xy = torch.cat([pts[:, :, 1:2], pts[:, :, 0:1]], dim=2).reshape(-1, 18).contiguous()
        bboxes = minaerarect(pts_xy, scale=float(stride), center=points)         # rect*stride + centre (:748-749)
        reppoints = pts_xy * stride + points.repeat(1, 9)                         # :754-760
        mlvl_bboxes.append(bboxes)
        mlvl_scores.append(scores)
        mlvl_reppoints.append(reppoints)
    mlvl_bboxes = torch.cat(mlvl_bboxes)
    mlvl_reppoints = torch.cat(mlvl_reppoints)
    if rescale:
        mlvl_bboxes = mlvl_bboxes / mlvl_bboxes.new_tensor(scale_factor)
        mlvl_reppoints = mlvl_reppoints / mlvl_reppoints.new_tensor(scale_factor)
    mlvl_scores = torch.cat(mlvl_scores)
    padding = mlvl_scores.new_zeros(mlvl_scores.shape[0], 1)
    mlvl_scores = torch.cat([padding, mlvl_scores], dim=1)
    if nms:
        return multiclass_rnms(mlvl_bboxes, mlvl_scores, cfg['score_thr'], cfg['nms'], cfg['max_per_img'],
                               multi_reppoints=mlvl_reppoints)
    return mlvl_bboxes, mlvl_scores


def scalar_scale_factor(sf):
    """img_meta['scale_factor'] as one number.  mmdet's Resize writes a float with keep_ratio=True (every config of this path)
    and a 4-vector (w, h, w, h) otherwise; the head divides its 8 box / 18 point coordinates by it (:766-768), which only
    broadcasts for a scalar - a vector is accepted here when all its entries agree"""
    import numpy as np
    a = np.asarray(sf, dtype=np.float64).reshape(-1)
    if a.size == 0 or not np.all(a == a[0]):
        raise ValueError("scale_factor %r: the rotated-box head needs one scale for x and y (keep_ratio=True)" % (sf,))
    return float(a[0])


def get_bboxes_fused(cls_scores, pts_preds_refine, strides, img_metas, cfg, rescale=False):
    """The same computation as get_bboxes() as ONE device-resident pipeline (orp_head_postprocess): returns
    padded (dets [B,max_per_img,27], labels [B,max_per_img], counts [B]) device tensors, no host sync."""
    import ctypes

    from .. import _lib
    n = len(cls_scores)
    b = cls_scores[0].shape[0]
    dev = cls_scores[0].device
    cls_c = [c.contiguous() for c in cls_scores]
    ref_c = [p.contiguous() for p in pts_preds_refine]
    pa = (ctypes.c_void_p * n)(*[c.data_ptr() for c in cls_c])
    pr = (ctypes.c_void_p * n)(*[p.data_ptr() for p in ref_c])
    hs = (ctypes.c_int * n)(*[c.shape[1] for c in cls_c])
    ws = (ctypes.c_int * n)(*[c.shape[2] for c in cls_c])
    ss = (ctypes.c_int * n)(*[int(s) for s in strides])
    cap = int(cfg['max_per_img'])
    dets = torch.empty((b, cap, 27), dtype=torch.float32, device=dev)
    labels = torch.empty((b, cap), dtype=torch.int64, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)
    sf = None
    if rescale:
        sf = torch.tensor([scalar_scale_factor(m['scale_factor']) for m in img_metas], dtype=torch.float32).to(dev, non_blocking=True)
    nms_cfg = cfg['nms']
    if nms_cfg.get('type', 'rnms') != 'rnms' or nms_cfg.get('mode', 'exact64') != 'exact64':
        raise ValueError("get_bboxes_fused serves nms type 'rnms' in the default arithmetic; use get_bboxes() for %r" % (nms_cfg,))
    with torch.cuda.device(dev):
        rc = _lib.lib().orp_head_postprocess(n, pa, pr, hs, ws, ss, b, cls_c[0].shape[3], int(cfg.get('nms_pre', -1)),
                                             float(cfg['score_thr']), float(nms_cfg['iou_thr']), cap, _lib.ptr(sf),
                                             _lib.ptr(dets), _lib.ptr(labels), _lib.ptr(counts), _lib.current_stream_ptr())
    _lib.check(rc, "orp_head_postprocess")
    return dets, labels, counts


def get_bboxes(cls_scores, pts_preds_refine, strides, img_metas, cfg, rescale=False, nms=True):
    """cls_scores[l]: [N,H,W,15]; pts_preds_refine[l]: [N,H,W,18] -> list of (dets [k,27], labels [k])"""
    out = []
    for img_id in range(cls_scores[0].shape[0]):
        out.append(get_bboxes_single([c[img_id] for c in cls_scores], [p[img_id] for p in pts_preds_refine], strides,
                                     img_metas[img_id]['scale_factor'], cfg, rescale, nms))
    return out
