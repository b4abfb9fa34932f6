"""Golden vectors for ResultMerge: runs the REFERENCE's own DOTA_devkit/ResultMerge_multi_process.py
(mergesingle :182-223 with py_cpu_nms_poly_fast :60-121 over the SWIG polyiou compiled by oracle/build_ref.py from the
reference's polyiou.cpp) on synthetic tile-level Task1 lines, and stores input lines + merged output lines.

    python oracle/build_ref.py && python tests/golden/gen_golden_result_merge.py     # needs /root/reference
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/DOTA_devkit"


def make_lines(seed=0):
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as po
    rng = np.random.RandomState(seed)
    lines = []
    for img in ("P0007", "P0003", "P0011"):
        for (tx, ty) in ((0, 0), (824, 0), (0, 824), (824, 824)):
            d = po.gen_rotated_boxes(120, seed=int(rng.randint(1 << 30)), extent=1024.0)
            for r in d:
                lines.append("%s__1__%d___%d %s %s\n" % (img, tx, ty, repr(float(r[8])), " ".join("%.1f" % v for v in r[:8])))
    rng.shuffle(lines)
    return lines


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import polyiou                                            # SWIG module built from the reference's polyiou.cpp
    pkg = types.ModuleType("DOTA_devkit")
    pkg.__path__ = []
    utils = types.ModuleType("DOTA_devkit.dota_utils")      # the real one pulls shapely; mergesingle needs one helper
    utils.custombasename = lambda fullname: os.path.basename(os.path.splitext(fullname)[0])
    sys.modules["DOTA_devkit"] = pkg
    sys.modules["DOTA_devkit.dota_utils"] = utils
    sys.modules["DOTA_devkit.polyiou"] = polyiou
    pkg.dota_utils, pkg.polyiou = utils, polyiou
    sys.path.insert(0, REF)
    import ResultMerge_multi_process as R
    lines = make_lines(0)
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "Task1_plane.txt")
        open(src, "w").write("".join(lines))
        dst = os.path.join(td, "out")
        os.mkdir(dst)
        R.mergesingle(dst, R.py_cpu_nms_poly_fast, src)
        merged = open(os.path.join(dst, "Task1_plane.txt")).read().splitlines()
        os.mkdir(os.path.join(td, "out2"))
        R.mergesingle(os.path.join(td, "out2"), R.py_cpu_nms_poly, src)             # the plain variant gives the same keep set
        merged_plain = open(os.path.join(td, "out2", "Task1_plane.txt")).read().splitlines()
    assert merged == merged_plain
    json.dump({"nms_thresh": R.nms_thresh, "lines": lines, "merged": merged}, open(os.path.join(HERE, "result_merge.json"), "w"))
    print("wrote result_merge.json:", len(lines), "tile-level lines ->", len(merged), "merged lines")


if __name__ == "__main__":
    main()
