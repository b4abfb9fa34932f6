"""Host-side logic that needs no GPU: aug_test's box mapping (mmdet/models/detectors/orientedreppoints_detector.py:48-110),
rbbox2result (mmdet/core/bbox/transforms.py:356-375), detection packing for the all-gather."""
import numpy as np
import pytest
import torch


def test_rbbox_flip_matches_reference_formula():
    from orientedreppoints_b200.detector import OrientedRepPointsDetector as D
    g = torch.Generator().manual_seed(0)
    b = torch.rand(7, 16, generator=g) * 100                       # two boxes per row (8*k columns, :51)
    shape = (120, 200, 3)
    f = D.rbbox_flip(b, shape)
    ref = b.clone()
    for k in (0, 2, 4, 6):                                         # the reference's four strided assignments (:58-61)
        ref[..., k::8] = shape[1] - b[..., k::8] - 1
    assert torch.equal(f, ref)
    v = D.rbbox_flip(b, shape, 'vertical')
    ref = b.clone()
    for k in (1, 3, 5, 7):
        ref[..., k::8] = shape[0] - b[..., k::8] - 1
    assert torch.equal(v, ref)
    with pytest.raises(ValueError):
        D.rbbox_flip(b, shape, 'diagonal')
    with pytest.raises(AssertionError):
        D.rbbox_flip(torch.zeros(3, 9), shape)


def test_merge_aug_results():
    from orientedreppoints_b200.detector import OrientedRepPointsDetector as D

    class _Self:
        rbbox_flip = staticmethod(D.rbbox_flip)
    b1, b2 = torch.rand(5, 8) * 50, torch.rand(3, 8) * 50
    s1, s2 = torch.rand(5, 16), torch.rand(3, 16)
    metas = [[dict(img_shape=(64, 96, 3), scale_factor=0.5, flip=False)], [dict(img_shape=(64, 96, 3), scale_factor=2.0, flip=True)]]
    mb, ms = D.merge_aug_results(_Self(), [b1, b2], [s1, s2], metas)
    assert mb.shape == (8, 8) and ms.shape == (8, 16)
    assert torch.equal(mb[:5], b1 / 0.5) and torch.equal(ms, torch.cat([s1, s2]))
    assert torch.equal(mb[5:], D.rbbox_flip(b2, (64, 96, 3)) / 2.0)
    assert torch.equal(D.merge_aug_results(_Self(), [b1, b2], None, metas), mb)


def test_rbbox2result_layout():
    from orientedreppoints_b200.core.transforms import rbbox2result
    d = torch.arange(6 * 27, dtype=torch.float32).reshape(6, 27)
    l = torch.tensor([0, 3, 3, 14, 0, 7])
    r = rbbox2result(d, l, 16)
    assert len(r) == 15 and [a.shape[0] for a in r] == [2, 0, 0, 2, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1]
    assert np.array_equal(r[3], d[[1, 2]].numpy())
    e = rbbox2result(torch.zeros(0, 27), torch.zeros(0, dtype=torch.int64), 16)
    assert len(e) == 15 and all(a.shape == (0, 9) for a in e)     # the reference's empty case (transforms.py:367-370)


def test_pack_roundtrip_single_rank():
    from orientedreppoints_b200 import gather as G
    cap, tiles = 8, 2
    dets = torch.rand(tiles, cap, 27)
    labels = torch.randint(0, 15, (tiles, cap))
    counts = torch.tensor([3, 8], dtype=torch.int32)
    buf, cnt = G.pack(dets, labels, counts)
    assert buf.shape == (tiles, cap + 1, 28) and torch.equal(cnt, counts)
    assert torch.equal(buf[0, :3, :27], dets[0, :3]) and torch.equal(buf[1, :cap, 27].long(), labels[1])
    assert buf[:, cap, 0].tolist() == [3.0, 8.0]
    ab, ac = G.all_gather_detections(buf)                         # no process group: world of one
    assert ab.shape == (1, tiles, cap, 28) and torch.equal(ac, counts.unsqueeze(0))
    h = G.all_gather_detections(buf, async_op=True)
    ab2, ac2 = h.wait()
    assert torch.equal(ab2, ab) and torch.equal(ac2, ac)
    out = G.interleave(ab, ac, dataset_len=2)
    assert torch.equal(out[0][0], dets[0, :3]) and torch.equal(out[1][1], labels[1])


def test_host_helpers_against_reference_golden():
    """tests/golden/host_helpers.npz holds outputs of the reference's OWN rbbox2result / rbbox_flip / merge_aug_results
    (extracted with ast and executed by tests/golden/gen_golden_host.py)"""
    import os
    from orientedreppoints_b200.core.transforms import rbbox2result
    from orientedreppoints_b200.detector import OrientedRepPointsDetector as D
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "host_helpers.npz"))
    res = rbbox2result(torch.from_numpy(g["r2r_boxes"]), torch.from_numpy(g["r2r_labels"]), 16)
    assert [a.shape[0] for a in res] == g["r2r_counts"].tolist()
    assert np.array_equal(np.concatenate(res, 0), g["r2r_concat"])
    assert tuple(rbbox2result(torch.zeros(0, 27), torch.zeros(0, dtype=torch.long), 16)[0].shape) == tuple(g["r2r_empty_shape"])
    b = torch.from_numpy(g["flip_in"])
    assert np.array_equal(D.rbbox_flip(b, (120, 200, 3)).numpy(), g["flip_h"])
    assert np.array_equal(D.rbbox_flip(b, (120, 200, 3), "vertical").numpy(), g["flip_v"])

    class _Self:
        rbbox_flip = staticmethod(D.rbbox_flip)
    metas = [[dict(img_shape=(64, 96, 3), scale_factor=0.5, flip=False)], [dict(img_shape=(64, 96, 3), scale_factor=2.0, flip=True)],
             [dict(img_shape=(128, 192, 3), scale_factor=1.5, flip=True)]]
    mb, ms = D.merge_aug_results(_Self(), [torch.from_numpy(g[k]) for k in ("m_b1", "m_b2", "m_b3")],
                                 [torch.from_numpy(g[k]) for k in ("m_s1", "m_s2", "m_s3")], metas)
    assert np.array_equal(mb.numpy(), g["m_out_b"]) and np.array_equal(ms.numpy(), g["m_out_s"])


def test_f16x3_weight_row_padding_rule():
    """output columns the tensor-core kernel computes per layer (engine_tc.EngineTCSplit._pad_cout): power-of-two widths of the
    ResNets unchanged, small fp32 heads in 32s, Swin's 192-channel layers as ONE 256-wide tile, everything else in 64s"""
    from orientedreppoints_b200.engine_tc import EngineTCSplit as E
    want = {15: 32, 18: 32, 32: 32, 64: 64, 96: 128, 128: 128, 192: 256, 256: 256, 288: 320, 384: 384, 512: 512, 576: 576,
            768: 768, 1024: 1024, 1152: 1152, 2048: 2048}
    for cout, p in want.items():
        assert E._pad_cout(cout) == p, cout
        assert p >= cout and (p % 32 == 0)
