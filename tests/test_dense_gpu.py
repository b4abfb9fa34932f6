"""GPU: the dense path (backbone / FPN / head / DCN) against the PyTorch fp32 re-declaration of the
reference graph (oracle/torch_reference.py).  fp32 engine: tight tolerances (same arithmetic up to
summation order); bf16 tensor-core engine: tolerances stated per test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.fixture(scope="module")
def sd():
    from orientedreppoints_b200.weights import random_state_dict
    return random_state_dict(50, seed=0, reference_init=False)


def test_conv_and_dcn_f32_vs_torch(cuda):
    from oracle import torch_reference as tr
    from orientedreppoints_b200.detector import ConvLayer, EngineF32
    import torch.nn.functional as F
    e = EngineF32(cuda)
    g = torch.Generator().manual_seed(0)
    for (cin, cout, k, s, p, h, w) in [(64, 64, 3, 1, 1, 37, 53), (128, 256, 1, 2, 0, 40, 40), (256, 18, 1, 1, 0, 19, 23),
                                       (4, 64, 7, 2, 3, 96, 80), (256, 256, 3, 2, 1, 33, 31)]:
        x = torch.randn(2, cin, h, w, generator=g)
        wt = torch.randn(cout, cin, k, k, generator=g) * 0.05
        b = torch.randn(cout, generator=g)
        ref = F.conv2d(x.double(), wt.double(), b.double(), s, p).float()
        L = ConvLayer(wt, b, s, p, cuda)
        y = e.conv(x.permute(0, 2, 3, 1).contiguous().to(cuda), L)
        assert _rel(_nchw(y).cpu(), ref) < 1e-5, (cin, cout, k, s)
    # deformable conv incl. samples that leave the image
    x = torch.randn(2, 64, 21, 27, generator=g)
    off = torch.randn(2, 18, 21, 27, generator=g) * 3.0
    wt = torch.randn(32, 64, 3, 3, generator=g) * 0.05
    ref = tr.deform_conv_ref(x.double(), off.double(), wt.double()).float()
    L = ConvLayer(wt, None, 1, 1, cuda)
    y = e.deform_conv(x.permute(0, 2, 3, 1).contiguous().to(cuda), off.permute(0, 2, 3, 1).contiguous().to(cuda), L)
    assert _rel(_nchw(y).cpu(), ref) < 1e-5
    # DCNv2 (mask) surface
    m = torch.rand(2, 9, 21, 27, generator=g)
    ref = tr.deform_conv_ref(x.double(), off.double(), wt.double(), mask=m.double()).float()
    y = e.deform_conv(x.permute(0, 2, 3, 1).contiguous().to(cuda), off.permute(0, 2, 3, 1).contiguous().to(cuda), L,
                      mask=m.permute(0, 2, 3, 1).contiguous().to(cuda))
    assert _rel(_nchw(y).cpu(), ref) < 1e-5


def test_dense_graph_f32_vs_torch(cuda, sd):
    from oracle import torch_reference as tr
    from orientedreppoints_b200.detector import OrientedRepPointsDetector
    det = OrientedRepPointsDetector(sd, 50, cuda, "fp32")
    img = torch.randn(2, 3, 256, 320, generator=torch.Generator().manual_seed(1))
    outs, feats = det.forward_dense(img.to(cuda))
    # fp64 torch graph as the gold standard (torch's fp32 GPU convs silently use TF32)
    sdg = {k: v.to(cuda).double() for k, v in sd.items()}
    with torch.no_grad():
        ref_outs, ref_feats = tr.forward_dense(sdg, img.to(cuda).double())
    ref_feats = [f.float() for f in ref_feats]
    ref_outs = [[t.float() for t in o] for o in ref_outs]
    for lvl in range(5):
        assert _rel(_nchw(feats[lvl]), ref_feats[lvl]) < 2e-4, lvl
        for k, name in enumerate(("cls", "init", "refine")):
            a, b = _nchw(outs[lvl][k]), ref_outs[lvl][k]
            assert a.shape == b.shape
            assert float((a - b).abs().max()) < 2e-4 * max(1.0, float(b.abs().max())), (lvl, name)


def test_simple_test_matches_reference_pipeline(cuda, sd):
    """whole tile: detections (boxes, reppoints, scores, labels, ORDER) against the restated reference
    post-processing over the torch graph's outputs"""
    from oracle import torch_reference as tr
    from orientedreppoints_b200.detector import OrientedRepPointsDetector
    cfg = dict(score_thr=0.02)
    det = OrientedRepPointsDetector(sd, 50, cuda, "fp32", test_cfg=cfg)
    img = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(2))
    from orientedreppoints_b200.core.get_bboxes import get_bboxes
    from orientedreppoints_b200.detector import STRIDES
    outs, _ = det.forward_dense(img.to(cuda))
    (dets, labels), = get_bboxes([o[0] for o in outs], [o[2] for o in outs], STRIDES, [dict(scale_factor=1.0)],
                                 det.test_cfg, rescale=True)
    # post-processing oracle on the SAME dense outputs (GN statistics use atomics: two forward passes differ
    # in the last bit); sigmoid evaluated by the same device routine, everything after it on the CPU oracle
    cls = [_nchw(o[0])[0] for o in outs]
    ref = [_nchw(o[2])[0].cpu() for o in outs]
    rd, rl = tr.get_bboxes_single(cls, ref, score_thr=0.02)
    assert dets.shape == rd.shape and dets.shape[0] > 10
    assert torch.equal(labels.cpu(), rl)                                    # index work: bit-exact
    assert float((dets.cpu() - rd).abs().max()) < 1e-3                      # coordinates in pixels (1e-4 * stride scale)
    assert torch.equal(dets[:, -1].cpu(), rd[:, -1])                        # scores
