"""Golden vectors for the head post-processing (SURVEY §8 a7 + a9): the reference's OWN python
  OrientedRepPointsHead.get_bboxes_single   (mmdet/models/anchor_heads/orientedreppoints_head.py:707-779)
  multiclass_rnms                           (mmdet/core/post_processing/bbox_nms.py:93-182)
are extracted with `ast` (the mmdet package cannot be imported here: mmcv 0.6.2 is absent) and executed on the CPU with
the two CUDA-only ops they call replaced by their CPU oracles - `minaerarect` by oracle_minarearect.c and
`nms_wrapper.rnms` by the fp64 polygon NMS (ascending keep indices, like rnms_kernel.cu:262-265 returns them).

    python tests/golden/gen_golden_postprocess.py     # needs /root/reference; writes tests/golden/postprocess.npz
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference/mmdet"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po                                                    # noqa: E402


def extract(path, name):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return textwrap.dedent(ast.get_source_segment(src, node))
    raise KeyError(name)


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    def minaerarect(pts):
        rect, _, _ = po.minarearect(pts.numpy())
        return torch.from_numpy(rect)

    def rnms(dets, iou_thr, device_id=None):
        keep = np.sort(po.nms_poly_f64(dets.numpy().astype(np.float64), iou_thr))
        inds = torch.from_numpy(keep).long()
        return dets[inds, :], inds

    nms_wrapper = types.SimpleNamespace(rnms=rnms)
    ns = {"torch": torch, "np": np, "nms_wrapper": nms_wrapper}
    exec(extract(os.path.join(REF, "core/post_processing/bbox_nms.py"), "multiclass_rnms"), ns)
    ns["minaerarect"] = minaerarect
    exec(extract(os.path.join(REF, "models/anchor_heads/orientedreppoints_head.py"), "get_bboxes_single"), ns)
    head = types.SimpleNamespace(cls_out_channels=15, use_sigmoid_cls=True, num_points=9, point_strides=[8, 16, 32, 64, 128])

    out = {}
    shapes = [(24, 32), (12, 16), (6, 8), (3, 4), (2, 2)]
    for case, (seed, score_thr, nms_pre, max_per_img) in enumerate([(0, 0.3, 300, 200), (1, 0.5, 2000, 20000), (2, 0.45, 100, 50), (3, 0.8, 150, 2000)]):
        g = torch.Generator().manual_seed(seed)
        cls = [torch.randn(15, h, w, generator=g) for h, w in shapes]
        pts = [torch.randn(18, h, w, generator=g) * 1.5 for h, w in shapes]
        mlvl_points = []
        for (h, w), s in zip(shapes, head.point_strides):                       # PointGenerator.grid_points, point_generator.py:14-22
            xs, ys = torch.arange(0, w) * s, torch.arange(0, h) * s
            xx, yy = xs.repeat(len(ys)), ys.view(-1, 1).repeat(1, len(xs)).view(-1)
            mlvl_points.append(torch.stack([xx, yy, xx.new_full((xx.shape[0],), s)], dim=-1).float())
        cfg = Cfg(nms_pre=nms_pre, min_bbox_size=0, score_thr=score_thr, nms=dict(type='rnms', iou_thr=0.4), max_per_img=max_per_img)
        dets, labels = ns["get_bboxes_single"](head, cls, pts, mlvl_points, (192, 256, 3), 1.0, cfg, False, True)
        out["c%d_cfg" % case] = np.array([seed, score_thr, nms_pre, max_per_img], dtype=np.float64)
        for l in range(5):
            out["c%d_cls%d" % (case, l)] = cls[l].numpy()
            out["c%d_pts%d" % (case, l)] = pts[l].numpy()
        out["c%d_dets" % case] = dets.numpy()
        out["c%d_labels" % case] = labels.numpy()
        print("case", case, "->", tuple(dets.shape))
    np.savez_compressed(os.path.join(HERE, "postprocess.npz"), **out)


if __name__ == "__main__":
    main()
