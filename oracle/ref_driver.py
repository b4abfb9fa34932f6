"""CPU baseline drivers (TEST/BENCH INFRASTRUCTURE ONLY - never imported by the product path).

`py_cpu_nms_poly_driver` restates the CALL PATTERN of the reference's python NMS loop
(DOTA_devkit/ResultMerge.py:18-41: argsort descending, one SWIG iou_poly call per remaining box,
keep `iou <= thresh`) over the reference's OWN compiled SWIG module oracle/_ref/_polyiou*.so
(built from DOTA_devkit/polyiou_wrap.cxx + polyiou.cpp by oracle/build_ref.py).  /root/reference
does not exist on the GPU box, so the ten-line python loop is restated here while every IoU is
computed by the reference binary.  If oracle/_ref is absent the C oracle port is used instead and
the baseline is labelled "port".
"""
import os
import sys
import time

import numpy as np

from . import pyoracle as po


def _swig():
    if not os.path.exists(os.path.join(po.REF_DIR, "polyiou.py")):
        return None
    if po.REF_DIR not in sys.path:
        sys.path.insert(0, po.REF_DIR)
    try:
        import polyiou  # the reference's SWIG module
        return polyiou
    except Exception:
        return None


def py_cpu_nms_poly_driver(dets, thresh):
    """returns (keep list, number of iou_poly calls, kind)"""
    polyiou = _swig()
    if polyiou is None:
        keep = po.nms_poly_f64(dets, thresh)
        n = len(dets)
        return list(keep), None, "port"
    dets = np.asarray(dets, dtype=np.float64)
    scores = dets[:, 8]
    polys = [polyiou.VectorDouble([float(v) for v in dets[i, :8]]) for i in range(len(dets))]
    order = scores.argsort()[::-1]
    keep, calls = [], 0
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        ovr = np.array([polyiou.iou_poly(polys[i], polys[order[j + 1]]) for j in range(order.size - 1)])
        calls += order.size - 1
        inds = np.where(ovr <= thresh)[0]
        order = order[inds + 1]
    return keep, calls, "reference"


def _shard_worker(args):
    dets, thresh = args
    t = time.perf_counter()
    keep, calls, kind = py_cpu_nms_poly_driver(dets, thresh)
    return len(keep), calls, kind, time.perf_counter() - t


def timed_baseline(n_boxes, thresh, shards, seed=0, extent=1024.0):
    """`shards` independent box sets of n_boxes each, one process per shard (the reference's
    mergebase_parallel uses Pool(16) over class files, ResultMerge_multi_process.py:225-231).
    Returns dict(value=Mpairs/s over N(N-1)/2 pairs per shard, seconds, cores, kind, sample)."""
    import multiprocessing as mp
    sets = [po.gen_rotated_boxes(n_boxes, seed=seed + s, extent=extent) for s in range(shards)]
    t0 = time.perf_counter()
    if shards == 1:
        res = [_shard_worker((sets[0], thresh))]
    else:
        with mp.get_context("fork").Pool(shards) as pool:
            res = pool.map(_shard_worker, [(s, thresh) for s in sets])
    dt = time.perf_counter() - t0
    pairs = shards * n_boxes * (n_boxes - 1) / 2.0
    return {"value": pairs / dt / 1e6, "unit": "Mpairs/s", "seconds": dt, "cores": shards,
            "kind": res[0][2], "kept": [r[0] for r in res],
            "sample": "%d shard(s) x %d boxes (SURVEY 8d generator, extent %.0f), thr %.2f, python loop of "
                      "ResultMerge.py:18-41 over the reference's compiled SWIG polyiou" % (shards, n_boxes, extent, thresh)}
