"""GPU: tile producer bit-exact against the reference's own splitter; aug_test (n3,
mmdet/models/detectors/orientedreppoints_detector.py:48-144)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "split_tiles.json")))


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda", 0)


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: "%dx%d" % (c["w"], c["h"]))
def test_split_tiles_bit_exact(cuda, case):
    from orientedreppoints_b200.dota.split_tiles import split_image
    img = np.random.RandomState(case["seed"]).randint(0, 256, size=(case["h"], case["w"], 3)).astype(np.uint8)
    tiles, names, origins = split_image(img, "P%04dx%04d" % (case["w"], case["h"]), 1, GOLD["subsize"], GOLD["gap"], device=cuda)
    assert tiles.shape == (len(case["tiles"]), 1024, 1024, 3) and tiles.dtype == torch.uint8
    host = tiles.cpu().numpy()
    for i, (name, sha) in enumerate(case["tiles"]):
        assert names[i] == name
        assert hashlib.sha1(np.ascontiguousarray(host[i]).tobytes()).hexdigest() == sha, name


def test_split_tiles_feed_the_detector(cuda):
    """tiles come out in the layout simple_test() takes (uint8 HWC batch)"""
    from orientedreppoints_b200.detector import OrientedRepPointsDetector
    from orientedreppoints_b200.dota.split_tiles import split_image
    from orientedreppoints_b200.weights import random_state_dict
    img = np.random.RandomState(3).randint(0, 256, size=(300, 420, 3)).astype(np.uint8)
    tiles, names, _ = split_image(img, "P1", 1, subsize=256, gap=64, device=cuda)
    assert tiles.shape[0] == len(names) == 4
    det = OrientedRepPointsDetector(random_state_dict(50, seed=0, reference_init=True), 50, cuda, "bf16", test_cfg=dict(score_thr=0.0))
    res = det.simple_test(tiles, return_tensors=True)
    assert len(res) == 4 and all(d.shape[1] == 27 for d, _ in res)


def _meta(shape, flip, sf=1.0):
    return [dict(img_shape=shape, scale_factor=sf, flip=flip)]


def _freeze_dense(det, views):
    """GroupNorm sums use atomics (not bit-reproducible run to run): evaluate the dense graph once per view and let
    both sides of a comparison consume the same outputs"""
    cache = {id(v): det.forward_dense(v) for v in views}
    det.forward_dense = lambda v: cache[id(v)]


def test_aug_test_single_view_equals_simple_test(cuda):
    from orientedreppoints_b200.detector import OrientedRepPointsDetector
    from orientedreppoints_b200.weights import random_state_dict
    det = OrientedRepPointsDetector(random_state_dict(50, seed=0, reference_init=True), 50, cuda, "fp32",
                                    test_cfg=dict(score_thr=0.0, max_per_img=300))
    det.fused_post = False
    img = torch.randn(1, 3, 128, 160, generator=torch.Generator().manual_seed(2)).to(cuda)
    _freeze_dense(det, [img])
    aug = det.aug_test([img], [_meta((128, 160, 3), False)], rescale=True)
    ref = det.simple_test(img, [dict(scale_factor=1.0)], rescale=True)[0]
    assert len(aug) == len(ref) == 15
    for a, r in zip(aug, ref):
        assert a.shape[0] == r.shape[0]
        if a.shape[0]:
            assert np.array_equal(a[:, :9], r[:, 18:27])       # same boxes + scores (simple_test rows carry 18 reppoints first)


def test_aug_test_flip_merge(cuda):
    """two views (identity + horizontal flip): the merged candidate set is the concatenation, flipped boxes are mapped
    back by x -> w - x - 1 (orientedreppoints_detector.py:48-79), then ONE multiclass_rnms"""
    from orientedreppoints_b200.core.bbox_nms import multiclass_rnms
    from orientedreppoints_b200.core.get_bboxes import get_bboxes
    from orientedreppoints_b200.detector import STRIDES, OrientedRepPointsDetector
    from orientedreppoints_b200.weights import random_state_dict
    det = OrientedRepPointsDetector(random_state_dict(50, seed=0, reference_init=True), 50, cuda, "fp32",
                                    test_cfg=dict(score_thr=0.0, max_per_img=500))
    h, w = 128, 192
    img = torch.randn(1, 3, h, w, generator=torch.Generator().manual_seed(4)).to(cuda)
    views = [img, img.flip(-1)]
    metas = [_meta((h, w, 3), False, 0.5), _meta((h, w, 3), True, 0.5)]
    _freeze_dense(det, views)
    out = det.aug_test(views, metas, rescale=True)
    # independent restatement of the merge
    boxes, scores = [], []
    for v, m in zip(views, metas):
        outs, _ = det.forward_dense(v)
        b, s = get_bboxes([o[0] for o in outs], [o[2] for o in outs], STRIDES, m, det.test_cfg, False, nms=False)[0]
        if m[0]["flip"]:
            b = b.clone()
            b[:, 0::2] = w - b[:, 0::2] - 1
        boxes.append(b / m[0]["scale_factor"])
        scores.append(s)
    d, l = multiclass_rnms(torch.cat(boxes), torch.cat(scores), 0.0, det.test_cfg["nms"], 500)
    assert sum(o.shape[0] for o in out) == d.shape[0] <= 500
    for c in range(15):
        sel = d[l == c].cpu().numpy()
        assert out[c].shape == sel.shape
        assert np.array_equal(out[c], sel)
    # rescale=False multiplies the first view's scale factor back in (:139-141)
    out2 = det.aug_test(views, metas, rescale=False)
    tot = sum(o.shape[0] for o in out2)
    assert tot == d.shape[0]
    for c in range(15):
        assert np.allclose(out2[c][:, :8], out[c][:, :8] * 0.5, rtol=1e-6) and np.array_equal(out2[c][:, 8], out[c][:, 8])
    assert torch.allclose(det.rbbox_flip(det.rbbox_flip(d[:, :8], (h, w, 3)), (h, w, 3)), d[:, :8], atol=1e-4)   # w-(w-x-1)-1 in fp32
    with pytest.raises(ValueError):
        det.rbbox_flip(d[:, :8], (h, w, 3), "diagonal")


def test_detect_image_equals_file_based_merge(cuda, tmp_path):
    """tile producer -> detector -> ResultMerge in memory == the same through Task1 files and mergebypoly
    (the file-based mirror of ResultMerge_multi_process.py, itself checked against the restated reference)"""
    from orientedreppoints_b200.detector import OrientedRepPointsDetector
    from orientedreppoints_b200.dota import result_merge as rm
    from orientedreppoints_b200.dota.pipeline import DOTA_CLASSES, detect_image
    from orientedreppoints_b200.dota.split_tiles import split_image
    from orientedreppoints_b200.weights import random_state_dict
    det = OrientedRepPointsDetector(random_state_dict(50, seed=0, reference_init=True), 50, cuda, "bf16",
                                    test_cfg=dict(score_thr=0.0, max_per_img=60))
    img = np.random.RandomState(11).randint(0, 256, size=(420, 610, 3)).astype(np.uint8)
    # freeze the per-tile results (GroupNorm sums use atomics): both routes consume the same detections
    tiles, names, origins = split_image(img, "P0042", 1, 256, 64, device=cuda)
    assert len(names) == 6
    res = det.simple_test(tiles)
    det.simple_test = lambda t: res[:t.shape[0]] if t.shape[0] == len(res) else None
    merged = detect_image(det, img, "P0042", 1, subsize=256, gap=64, batch=16)
    assert set(merged) == set(DOTA_CLASSES)
    raw, out = tmp_path / "raw", tmp_path / "merged"
    rm.write_task1_raw(res, names, DOTA_CLASSES, str(raw))
    rm.mergebypoly(str(raw), str(out))
    total = 0
    for c in DOTA_CLASSES:
        lines = [l.rstrip("\n") for l in open(out / ("Task1_%s.txt" % c))]
        assert lines == merged[c], c
        total += len(lines)
        for l in lines:
            sp = l.split(" ")
            assert sp[0] == "P0042" and len(sp) == 10
    assert 0 < total <= 6 * 60
    # coordinates are in image space: a detection of the last tile is shifted by that tile's origin
    last = names[-1]
    (l, u) = origins[-1]
    assert (l, u) == (610 - 256, 420 - 256) and last.endswith("__%d___%d" % (l, u))
