"""GPU: the `mmdet.ops.dcn` operator surface (ops/dcn.py) through the reference's own signature - NCHW fp32 tensors,
offset [B, 18, H, W], weight [out, in, 3, 3] - against deform_conv_ref in fp64 (pinned to the reference's im2col kernels),
DCNv1 and DCNv2, offsets that leave the image; and the config-built module-level detector against the engine."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def test_deform_conv_nchw_signature_v1_v2(cuda):
    from oracle import torch_reference as tr
    from orientedreppoints_b200.ops import DeformConv, ModulatedDeformConv, deform_conv, modulated_deform_conv
    from orientedreppoints_b200.ops import dcn
    g = torch.Generator().manual_seed(0)
    for (b, cin, cout, h, w) in [(2, 256, 256, 40, 56), (1, 64, 96, 17, 23), (3, 128, 256, 9, 5)]:
        x = torch.randn(b, cin, h, w, generator=g)
        off = torch.randn(b, 18, h, w, generator=g) * 3.0           # many samples leave the image
        m = torch.rand(b, 9, h, w, generator=g)
        mod = DeformConv(cin, cout, 3, padding=1).to(cuda)
        ref = tr.deform_conv_ref(x.double(), off.double(), mod.weight.detach().cpu().double())
        with torch.no_grad():
            y = mod(x.to(cuda), off.to(cuda))
        assert y.shape == ref.shape and y.dtype == torch.float32 and y.is_cuda and y.is_contiguous()
        assert _rel(y.cpu(), ref) < 2.5e-5                           # f16x3 tensor-core path (default)
        with torch.no_grad():
            y2 = deform_conv(x.to(cuda), off.to(cuda), mod.weight, 1, 1, 1, 1, 1, 64)
        assert torch.equal(y, y2)
        mm = ModulatedDeformConv(cin, cout, 3, padding=1, bias=True).to(cuda)
        with torch.no_grad():
            mm.bias.normal_(0, 0.1, generator=None)
            y3 = mm(x.to(cuda), off.to(cuda), m.to(cuda))
        ref3 = tr.deform_conv_ref(x.double(), off.double(), mm.weight.detach().cpu().double(), mask=m.double()) \
            + mm.bias.detach().cpu().double().view(1, -1, 1, 1)
        assert _rel(y3.cpu(), ref3) < 2.5e-5
        with torch.no_grad():
            y4 = modulated_deform_conv(x.to(cuda), off.to(cuda), m.to(cuda), mm.weight, None, 1, 1, 1, 1, 1)
        assert _rel(y4.cpu(), ref3 - mm.bias.detach().cpu().double().view(1, -1, 1, 1)) < 2.5e-5
    # the other arithmetic modes of the same surface
    x = torch.randn(1, 64, 12, 12, generator=g)
    off = torch.randn(1, 18, 12, 12, generator=g)
    wt = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    ref = tr.deform_conv_ref(x.double(), off.double(), wt.double())
    try:
        dcn.set_precision("fp32")
        assert _rel(deform_conv(x.to(cuda), off.to(cuda), wt.to(cuda), 1, 1).cpu(), ref) < 1e-5
        dcn.set_precision("bf16")
        assert _rel(deform_conv(x.to(cuda), off.to(cuda), wt.to(cuda), 1, 1).cpu(), ref) < 2e-2
    finally:
        dcn.set_precision("f16x3")
    # channel counts the tensor-core kernel does not tile (Cin % 64 != 0) run on the fp32 kernel, same surface
    x = torch.randn(2, 24, 11, 13, generator=g)
    off = torch.randn(2, 18, 11, 13, generator=g)
    wt = torch.randn(10, 24, 3, 3, generator=g) * 0.1
    assert _rel(deform_conv(x.to(cuda), off.to(cuda), wt.to(cuda), 1, 1).cpu(), tr.deform_conv_ref(x.double(), off.double(), wt.double())) < 1e-5


def test_deform_conv_error_behaviour(cuda):
    from orientedreppoints_b200.ops import DeformConvPack, ModulatedDeformConvPack, deform_conv
    x = torch.randn(1, 64, 8, 8, device=cuda)
    wt = torch.randn(64, 64, 3, 3, device=cuda)
    with pytest.raises(RuntimeError):
        deform_conv(x, torch.zeros(1, 16, 8, 8, device=cuda), wt, 1, 1)        # offset channels != 2*kh*kw (deform_conv_cuda.cpp:130-136)
    with pytest.raises(RuntimeError):
        deform_conv(x, torch.zeros(1, 18, 7, 8, device=cuda), wt, 1, 1)        # offset height != output height
    with pytest.raises(ValueError):
        deform_conv(x[0], torch.zeros(1, 18, 8, 8, device=cuda), wt, 1, 1)
    with pytest.raises(NotImplementedError):
        deform_conv(x, torch.zeros(1, 18, 8, 8, device=cuda), wt[:, :32], 1, 1, 1, 2, 1)   # groups = 2: not built
    # zero offsets (freshly initialised Pack layers) = the plain convolution
    import torch.nn.functional as F
    p = DeformConvPack(64, 64, 3, padding=1, bias=False).to(cuda)
    with torch.no_grad():
        y = p(x)
    assert _rel(y.cpu(), F.conv2d(x.double().cpu(), p.weight.detach().double().cpu(), None, 1, 1)) < 2.5e-5
    p2 = ModulatedDeformConvPack(64, 64, 3, padding=1).to(cuda)
    with torch.no_grad():
        y2 = p2(x)                                               # mask = sigmoid(0) = 0.5
    assert _rel(y2.cpu(), 0.5 * F.conv2d(x.double().cpu(), p2.weight.detach().double().cpu(), None, 1, 1)) < 2.5e-5
    # input smaller than the kernel: padded and cropped back (deform_conv.py:239-255)
    from orientedreppoints_b200.ops import DeformConv
    m = DeformConv(64, 64, 3, padding=1).to(cuda)
    with torch.no_grad():
        out = m(torch.randn(1, 64, 2, 2, device=cuda), torch.zeros(1, 18, 2, 2, device=cuda))
    assert out.shape == (1, 64, 2, 2)


def test_config_built_detector_runs_on_engine(cuda):
    """build_detector(reference config dict) -> simple_test == the engine fed with the same state_dict"""
    import importlib.util
    import os
    from orientedreppoints_b200.detector import OrientedRepPointsDetector as Engine
    from orientedreppoints_b200.models import build_detector
    from orientedreppoints_b200.weights import random_state_dict
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("c", os.path.join(root, "configs", "dota", "orientedrepoints_r50_demo.py"))
    cfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfg)
    tc = dict(cfg.test_cfg, score_thr=0.02)
    det = build_detector(cfg.model, test_cfg=tc)
    sd = random_state_dict(50, seed=0, reference_init=False)
    det.load_state_dict(sd, strict=True)
    det = det.to(cuda)
    img = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(3)).to(cuda)
    res = det(img, [dict(scale_factor=1.0)], return_loss=False, rescale=True)
    eng = Engine(sd, 50, cuda, "f16x3", test_cfg=tc)
    ref = eng.simple_test(img, [dict(scale_factor=1.0)], rescale=True)
    assert len(res) == 1 and len(res[0]) == 15
    # GroupNorm statistics are accumulated with atomics: two passes agree to the last bits, not bit for bit, so a candidate
    # sitting exactly on a threshold may fall either way - match detections by content (class + coordinates), not by position
    import numpy as np

    def flat(r):
        rows = [np.concatenate([a, np.full((a.shape[0], 1), c, np.float32)], 1) for c, a in enumerate(r) if a.shape[0]]
        return torch.from_numpy(np.concatenate(rows, 0)) if rows else torch.zeros((0, 28))

    fa, fb = flat(res[0]), flat(ref[0])
    assert fa.shape[0] > 10 and abs(fa.shape[0] - fb.shape[0]) <= max(2, fa.shape[0] // 50)
    d = torch.cdist(fa[:, :26].double(), fb[:, :26].double(), p=float("inf")) + (fa[:, 27:28] != fb[:, 27:28].T).double() * 1e6
    assert float((d.min(dim=1).values < 1e-2).float().mean()) > 0.97
    assert float((d.min(dim=0).values < 1e-2).float().mean()) > 0.97
