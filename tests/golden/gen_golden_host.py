"""Golden vectors for the host-side helpers either side of the kernels: the reference's OWN python functions are
extracted with `ast` (the mmdet package itself cannot be imported here: mmcv 0.6.2 is absent) and executed:

  mmdet/core/bbox/transforms.py                      rbbox2result
  mmdet/models/detectors/orientedreppoints_detector.py   rbbox_flip, rbox_mapping_back, merge_aug_results

    python tests/golden/gen_golden_host.py     # needs /root/reference; writes tests/golden/host_helpers.npz
"""
import ast
import os
import textwrap

import numpy as np
import torch

REF = "/root/reference/mmdet"
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out[node.name] = textwrap.dedent(ast.get_source_segment(src, node))
    ns = {"np": np, "torch": torch}
    for n in names:
        exec(out[n], ns)
    return ns


def main():
    t = extract(os.path.join(REF, "core/bbox/transforms.py"), ["rbbox2result"])
    d = extract(os.path.join(REF, "models/detectors/orientedreppoints_detector.py"),
                ["rbbox_flip", "rbox_mapping_back", "merge_aug_results"])

    class Det:                                    # the three methods only use each other through self
        rbbox_flip = d["rbbox_flip"]
        rbox_mapping_back = d["rbox_mapping_back"]
        merge_aug_results = d["merge_aug_results"]

    g = torch.Generator().manual_seed(0)
    out = {}
    boxes = torch.rand(9, 27, generator=g) * 100
    labels = torch.tensor([0, 3, 3, 14, 0, 7, 7, 7, 2])
    res = t["rbbox2result"](boxes, labels, 16)
    out["r2r_boxes"], out["r2r_labels"] = boxes.numpy(), labels.numpy()
    out["r2r_counts"] = np.array([a.shape[0] for a in res])
    out["r2r_concat"] = np.concatenate(res, 0)
    out["r2r_empty_shape"] = np.array(t["rbbox2result"](torch.zeros(0, 27), torch.zeros(0, dtype=torch.long), 16)[0].shape)

    b = torch.rand(7, 16, generator=g) * 200
    shape = (120, 200, 3)
    out["flip_in"] = b.numpy()
    out["flip_h"] = Det().rbbox_flip(b, shape).numpy()
    out["flip_v"] = Det().rbbox_flip(b, shape, "vertical").numpy()

    b1, b2, b3 = torch.rand(5, 8, generator=g) * 50, torch.rand(3, 8, generator=g) * 50, torch.rand(4, 8, generator=g) * 50
    s1, s2, s3 = torch.rand(5, 16, generator=g), torch.rand(3, 16, generator=g), torch.rand(4, 16, generator=g)
    metas = [[dict(img_shape=(64, 96, 3), scale_factor=0.5, flip=False)], [dict(img_shape=(64, 96, 3), scale_factor=2.0, flip=True)],
             [dict(img_shape=(128, 192, 3), scale_factor=1.5, flip=True)]]
    mb, ms = Det().merge_aug_results([b1, b2, b3], [s1, s2, s3], metas)
    for k, v in dict(m_b1=b1, m_b2=b2, m_b3=b3, m_s1=s1, m_s2=s2, m_s3=s3, m_out_b=mb, m_out_s=ms).items():
        out[k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "host_helpers.npz"), **out)
    print("wrote host_helpers.npz", sorted(out))


if __name__ == "__main__":
    main()
