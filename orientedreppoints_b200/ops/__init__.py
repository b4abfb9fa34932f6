"""Operator surface of the path, named as in the reference's mmdet/ops/__init__.py:1-32 (the names on the dense-inference
path; training-only ops - RoI pooling, focal loss, chamfer distance, ... - are out of scope)."""
from .box_iou_rotated import box_iou_rotated, quad_iou_matrix
from .conv import build_conv_layer
from .conv_module import ConvModule
from .convex_iou import convex_giou, convex_iou, convex_overlaps
from .dcn import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack, deform_conv,
                  modulated_deform_conv)
from .minarea_rect import minaerarect
from .nms_wrapper import rnms, rnms_indices, soft_rnms
from .norm import build_norm_layer

__all__ = ['rnms', 'rnms_indices', 'soft_rnms', 'minaerarect', 'box_iou_rotated', 'quad_iou_matrix', 'convex_iou', 'convex_overlaps',
           'convex_giou', 'DeformConv', 'DeformConvPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack', 'deform_conv',
           'modulated_deform_conv', 'ConvModule', 'build_conv_layer', 'build_norm_layer']
