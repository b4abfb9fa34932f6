"""Mint tests/golden/box_iou_rotated.npz from the reference's own CPU implementation
(mmdet/ops/box_iou_rotated/src/box_iou_rotated_cpu.cpp compiled unmodified by oracle/build_ref.py).
Authoring container only (needs /root/reference)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, pyoracle as po  # noqa: E402


def main():
    build_ref.build(verbose=False)
    import torch  # noqa: F401
    lib = ctypes.CDLL(os.path.join(po.REF_DIR, "ref_box_iou_rotated.so"))
    rng = np.random.RandomState(7)
    n, m = 96, 80
    base = np.stack([rng.uniform(0, 300, n), rng.uniform(0, 300, n), rng.uniform(8, 80, n), rng.uniform(4, 40, n),
                     rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
    q = base[rng.randint(0, n, m)].copy()
    q[:, :2] += rng.normal(0, 6, (m, 2)).astype(np.float32)          # near copies: plenty of overlap
    q[:, 4] += rng.normal(0, 0.3, m).astype(np.float32)
    q[:8] = base[:8]                                                  # identical boxes -> IoU 1
    out = np.zeros((n, m), np.float32)
    lib.ref_box_iou_rotated(base.ctypes.data_as(ctypes.c_void_p), n, q.ctypes.data_as(ctypes.c_void_p), m,
                            out.ctypes.data_as(ctypes.c_void_p))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "box_iou_rotated.npz"), b1=base, b2=q, iou=out)
    print("box_iou_rotated golden:", out.shape, "pairs > 0.05:", int((out > 0.05).sum()), "max", float(out.max()))


if __name__ == "__main__":
    main()
