"""Mirror of mmdet/core/post_processing/bbox_nms.py:93-182 (`multiclass_rnms`).

Same arguments, same output layout ([k, 18+8+1] rows = reppoints | box | score, 0-based labels, survivors
in ascending candidate order unless more than max_num survive, then top-max_num by score).  One deliberate
difference, documented in DESIGN.md: the reference makes NMS class-aware by adding label*(max_coord+1) to
every coordinate (:156-158) and running its fp32 rnms there, which is numerically unstable (511
vs 660 survivors on the 1k fixture at +16000).  Here the label is passed to the NMS kernel as the segment
id - the exact meaning of the offset trick - and coordinates are left untouched.

`nms_cfg = dict(type='rnms', iou_thr=..., mode='compat32')` switches to the reference's literal behaviour for
bit-for-bit reproduction runs: coordinates offset by label * (max + 1) exactly as :156-158 and the fp32 IoU arithmetic of
rnms_kernel.cu (ORP_NMS_COMPAT32), no segments.
"""
import torch

from ..ops import nms_wrapper


def multiclass_rnms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None,
                    multi_reppoints=None):
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 8:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 8)[:, 1:]
    else:
        bboxes = multi_bboxes[:, None].expand(-1, num_classes, 8)
    if multi_reppoints is not None:
        reppoints = multi_reppoints[:, None].expand(-1, num_classes, multi_reppoints.size(-1))
    scores = multi_scores[:, 1:]
    valid_mask = scores > score_thr
    bboxes = bboxes[valid_mask]
    if multi_reppoints is not None:
        reppoints = reppoints[valid_mask]
    if score_factors is not None:
        scores = scores * score_factors[:, None]
    scores = scores[valid_mask]
    labels = valid_mask.nonzero()[:, 1]
    if bboxes.numel() == 0:
        if multi_reppoints is None:
            bboxes = multi_bboxes.new_zeros((0, 9))
        else:
            bboxes = multi_bboxes.new_zeros((0, reppoints.size(-1) + 9))
        labels = multi_bboxes.new_zeros((0, ), dtype=torch.long)
        return bboxes, labels
    nms_cfg_ = nms_cfg.copy()
    nms_type = nms_cfg_.pop('type', 'rnms')
    nms_op = getattr(nms_wrapper, nms_type)
    if nms_cfg_.get('mode') == 'compat32':
        # the reference, literally: class-aware through coordinate offsets, fp32 IoU (bbox_nms.py:156-164)
        max_coordinate = bboxes.max()
        offsets = labels.to(bboxes) * (max_coordinate + 1)
        dets, keep = nms_op(torch.cat([bboxes + offsets[:, None], scores[:, None]], 1), **nms_cfg_)
    else:
        dets, keep = nms_op(torch.cat([bboxes, scores[:, None]], 1), segments=labels, **nms_cfg_)
    bboxes = bboxes[keep]
    if multi_reppoints is not None:
        reppoints = reppoints[keep]
        bboxes = torch.cat([reppoints, bboxes], dim=1)
    scores = dets[:, -1]
    labels = labels[keep]
    if keep.size(0) > max_num:
        _, inds = scores.sort(descending=True, stable=True)
        inds = inds[:max_num]
        bboxes = bboxes[inds]
        scores = scores[inds]
        labels = labels[inds]
    return torch.cat([bboxes, scores[:, None]], 1), labels
