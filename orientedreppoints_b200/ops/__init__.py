"""Operator surface of the path, named as in the reference's mmdet/ops/__init__.py:1-32."""
from .box_iou_rotated import box_iou_rotated, quad_iou_matrix
from .convex_iou import convex_giou, convex_iou, convex_overlaps
from .minarea_rect import minaerarect
from .nms_wrapper import rnms, rnms_indices

__all__ = ['rnms', 'rnms_indices', 'minaerarect', 'box_iou_rotated', 'quad_iou_matrix', 'convex_iou', 'convex_overlaps',
           'convex_giou']
