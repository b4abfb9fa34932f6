"""Mirror of mmdet/ops/iou/iou_wrapper.py (convex_iou :21-25, convex_overlaps :27-30): IoU between the convex hull of
each 9-point set and each quadrilateral.  convex_giou (:14-19) is the training loss with its gradient - out of scope
(inference path only)."""
import torch

from .. import _lib


def convex_iou(pred, target):
    """pred: [ex_num, 18] cuda float (x0,y0,...,x8,y8); target: [gt_num, 8] -> [ex_num, gt_num] on pred.device"""
    if not (torch.is_tensor(pred) and pred.is_cuda and torch.is_tensor(target) and target.is_cuda):
        raise TypeError('ex_boxes must be a CUDA tensor')          # convex_iou_kernel.cu:317-318 AT_ASSERTM
    ex_num, gt_num = pred.size(0), target.size(0)
    if ex_num == 0 or gt_num == 0:
        return pred.new_zeros((ex_num, gt_num), dtype=torch.float32)
    p = pred.detach().float().contiguous().reshape(ex_num, 18)
    t = target.detach().float().contiguous().reshape(gt_num, 8)
    out = torch.empty((ex_num, gt_num), dtype=torch.float32, device=pred.device)
    with torch.cuda.device(pred.device):
        _lib.check(_lib.lib().orp_convex_iou(_lib.ptr(p), ex_num, _lib.ptr(t), gt_num, _lib.ptr(out),
                                             _lib.current_stream_ptr()), "orp_convex_iou")
    return out


def convex_overlaps(gt_rbboxes, points):
    return convex_iou(points, gt_rbboxes).transpose(1, 0)


def convex_giou(pred, target):
    raise NotImplementedError("convex_giou is the training loss (forward + gradient); only the inference path is built")
