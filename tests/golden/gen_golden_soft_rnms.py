"""Golden vectors for soft_rnms minted by the REFERENCE ITSELF: mmdet/ops/nms/src/rnms_cpu.cpp compiled unmodified
(oracle/build_ref.py -> oracle/_ref/ref_rnms_cpu.so), its `soft_rnms(dets, thr, method, sigma, min_score)` entry
(rnms_cpu.cpp:165-320) run here for the three methods of mmdet/ops/nms/nms_wrapper.py:120-175.

    python tests/golden/gen_golden_soft_rnms.py      # needs /root/reference (authoring container only)
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    import torch
    from oracle import build_ref, pyoracle as po
    build_ref.build(verbose=False)
    spec = importlib.util.spec_from_file_location("ref_rnms_cpu", os.path.join(po.REF_DIR, "ref_rnms_cpu.so"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    sets = {"rand600": po.gen_rotated_boxes(600, seed=5, extent=256.0), "clustered": po.gen_clustered_boxes(40, 10, seed=9)}
    for name, d in sets.items():
        out[name + "_dets"] = d
        for method, code in (("original", 0), ("linear", 1), ("gaussian", 2)):
            for thr in (0.3, 0.5):
                r = ref.soft_rnms(torch.from_numpy(d.copy()), float(thr), code, 0.5, 1e-3).numpy()
                out["%s_%s_thr%02d" % (name, method, int(thr * 10))] = r            # [K,10]: 8 coords, new score, original index
    np.savez_compressed(os.path.join(OUT, "soft_rnms.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
