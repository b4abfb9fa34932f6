"""Golden vectors for the tile producer: runs the REFERENCE's own splitbase.SplitSingle
(DOTA_devkit/SplitOnlyImage_multi_process.py:51-87) in this container with cv2.imread / cv2.imwrite replaced by
recorders, and stores the tile names + a checksum of every tile for a few image sizes.

    python tests/golden/gen_golden_split.py        # needs /root/reference; writes tests/golden/split_tiles.json
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/DOTA_devkit"


def main():
    sys.path.insert(0, REF)
    # dota_utils pulls shapely etc.; SplitSingle itself never touches it
    sys.modules.setdefault("dota_utils", types.ModuleType("dota_utils"))
    import cv2
    import SplitOnlyImage_multi_process as S

    cases = [(1024, 1024), (1500, 900), (2048, 2048), (4000, 3000), (700, 500), (1024, 1025), (1849, 1848), (2672, 1024)]
    out = {"gap": 200, "subsize": 1024, "cases": []}
    for (w, h) in cases:
        rng = np.random.RandomState(w * 7 + h)
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        rec = []
        sb = S.splitbase.__new__(S.splitbase)
        sb.srcpath = sb.dstpath = sb.outpath = "/nonexistent"
        sb.gap, sb.subsize, sb.slide, sb.ext, sb.padding = 200, 1024, 824, ".png", True
        orig_read, orig_write = cv2.imread, cv2.imwrite
        cv2.imread = lambda path: img
        cv2.imwrite = lambda path, arr: rec.append((os.path.basename(path)[:-4],
                                                    hashlib.sha1(np.ascontiguousarray(arr).astype(np.uint8).tobytes()).hexdigest()))
        try:
            sb.SplitSingle("P%04dx%04d" % (w, h), 1, ".png")
        finally:
            cv2.imread, cv2.imwrite = orig_read, orig_write
        out["cases"].append({"w": w, "h": h, "seed": w * 7 + h, "tiles": rec})
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "split_tiles.json")
    json.dump(out, open(dst, "w"))
    print("wrote", dst, sum(len(c["tiles"]) for c in out["cases"]), "tiles")


if __name__ == "__main__":
    main()
