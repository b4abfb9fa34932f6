"""Mirror of the Cython modules DOTA_devkit/poly_nms_gpu/{poly_nms,poly_overlaps,nms_wrapper}
(poly_nms.pyx:9-24, poly_overlaps.pyx:4-12, nms_wrapper.py:11-16) over liborp_b200.so.

Host numpy in / host data out, blocking - these are the reference's host-facing entry points and go
through the *_host C-ABI functions that replace `_poly_nms` / `_overlaps`.
"""
import ctypes

import numpy as np

from .. import _lib


def poly_gpu_nms(dets, thresh, device_id=0):
    """np.float32 [N,9] -> list of kept ORIGINAL indices in score-descending selection order."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % dets.ndim)
    boxes_num, boxes_dim = dets.shape
    if boxes_num == 0:
        return []
    # poly_nms.pyx:19-21 sorts on the host (scores.argsort()[::-1]) and hands `_poly_nms` the sorted rows.  The device path
    # orders by score itself (ties: lower row first, what a stable host sort gives), so the rows go down unsorted and the
    # kept ORIGINAL indices come back in score order - the same list, without a 100k-element host sort per call.
    keep = np.empty(boxes_num, dtype=np.int32)
    num_out = ctypes.c_int(0)
    rc = _lib.lib().orp_poly_nms_host(keep.ctypes.data_as(ctypes.c_void_p), ctypes.addressof(num_out),
                                      dets.ctypes.data_as(ctypes.c_void_p), boxes_num, boxes_dim,
                                      float(thresh), int(device_id))
    _lib.check(rc, "orp_poly_nms_host")
    return keep[:num_out.value].tolist()


def poly_overlaps(boxes, query_boxes, device_id=0):
    """np.float32 [N,5], [K,5] (cx,cy,w,h,theta) -> np.float32 [N,K]."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    query_boxes = np.ascontiguousarray(query_boxes, dtype=np.float32)
    n, k = boxes.shape[0], query_boxes.shape[0]
    overlaps = np.zeros((n, k), dtype=np.float32)
    if n and k:
        rc = _lib.lib().orp_poly_overlaps_host(overlaps.ctypes.data_as(ctypes.c_void_p),
                                               boxes.ctypes.data_as(ctypes.c_void_p),
                                               query_boxes.ctypes.data_as(ctypes.c_void_p), n, k, int(device_id))
        _lib.check(rc, "orp_poly_overlaps_host")
    return overlaps


def poly_nms_gpu(dets, thresh, force_cpu=False):
    """nms_wrapper.py:11-16: [] for empty input; `force_cpu` is accepted and ignored as there."""
    if dets.shape[0] == 0:
        return []
    return poly_gpu_nms(dets, thresh, device_id=0)
