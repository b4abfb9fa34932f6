"""CPU: the module-structure side of the drop-in boundary: registries, conv / norm builders, ConvModule,
the DCN classes' constructor / error behaviour, and that the reference's config dicts build into parameter trees with the
reference's state_dict keys.  No kernel runs here."""
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "configs", "dota", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_registry_both_decorator_forms_and_errors():
    from orientedreppoints_b200.utils import Registry, build_from_cfg
    R = Registry('thing')

    @R.register_module
    class A(object):
        def __init__(self, x=1, y=2):
            self.x, self.y = x, y

    @R.register_module()
    class B(object):
        pass

    assert R.get('A') is A and R.get('B') is B and 'thing' in repr(R)
    with pytest.raises(KeyError):
        R.register_module(A)                                   # registry.py:39-41
    R.register_module(A, force=True)
    with pytest.raises(TypeError):
        R.register_module(3)
    a = build_from_cfg(dict(type='A', x=5), R, default_args=dict(y=7, x=0))
    assert (a.x, a.y) == (5, 7)
    assert isinstance(build_from_cfg(dict(type=B), R), B)
    with pytest.raises(KeyError):
        build_from_cfg(dict(type='C'), R)
    with pytest.raises(TypeError):
        build_from_cfg(dict(type=3), R)


def test_conv_and_norm_builders():
    import torch.nn as nn
    from orientedreppoints_b200.ops import (ConvModule, DeformConvPack, ModulatedDeformConvPack, build_conv_layer,
                                            build_norm_layer)
    assert isinstance(build_conv_layer(None, 8, 16, 3, padding=1), nn.Conv2d)
    d = build_conv_layer(dict(type='DCN'), 64, 64, 3, padding=1, bias=False)
    assert isinstance(d, DeformConvPack) and d.conv_offset.out_channels == 18 and tuple(d.weight.shape) == (64, 64, 3, 3)
    assert float(d.conv_offset.weight.abs().sum()) == 0.0       # init_offset
    d2 = build_conv_layer(dict(type='DCNv2'), 64, 32, 3, padding=1)
    assert isinstance(d2, ModulatedDeformConvPack) and d2.conv_offset.out_channels == 27 and d2.bias is not None
    with pytest.raises(KeyError):
        build_conv_layer(dict(type='Nope'), 1, 1, 1)
    name, gn = build_norm_layer(dict(type='GN', num_groups=32, requires_grad=True), 256)
    assert name == 'gn' and isinstance(gn, nn.GroupNorm) and gn.eps == 1e-5
    name, bn = build_norm_layer(dict(type='BN', requires_grad=False), 64, postfix=1)
    assert name == 'bn1' and not any(p.requires_grad for p in bn.parameters())
    with pytest.raises(KeyError):
        build_norm_layer(dict(type='LN'), 8)
    m = ConvModule(256, 256, 3, padding=1, norm_cfg=dict(type='GN', num_groups=32))
    assert sorted(m.state_dict()) == ['conv.weight', 'gn.bias', 'gn.weight'] and m.norm is m.gn and not m.with_bias
    y = m(torch.randn(1, 256, 5, 5))                            # plain layers, in order
    assert y.shape == (1, 256, 5, 5) and float(y.min()) >= 0.0
    assert ConvModule(8, 8, 1).with_bias and ConvModule(8, 8, 1, act_cfg=None).with_activation is False


def test_dcn_surface_errors_on_cpu():
    from orientedreppoints_b200.ops import DeformConv, ModulatedDeformConv, deform_conv, modulated_deform_conv
    with pytest.raises(AssertionError):
        DeformConv(8, 8, 3, bias=True)                          # deform_conv.py:206
    m = DeformConv(64, 32, 3, padding=1)
    assert tuple(m.weight.shape) == (32, 64, 3, 3) and m.stride == (1, 1) and m.transposed is False
    bound = 1.0 / (64 * 9) ** 0.5
    assert float(m.weight.abs().max()) <= bound
    x, off = torch.randn(1, 64, 6, 6), torch.zeros(1, 18, 6, 6)
    with pytest.raises(NotImplementedError):
        m(x, off)                                               # CPU tensors (:46-47)
    with pytest.raises(ValueError):
        deform_conv(torch.randn(64, 6, 6), off, m.weight)       # not 4-D (:27-30)
    mm = ModulatedDeformConv(64, 32, 3, padding=1)
    assert mm.bias is not None and float(mm.bias.abs().sum()) == 0.0
    with pytest.raises(NotImplementedError):
        modulated_deform_conv(x, off, torch.ones(1, 9, 6, 6), mm.weight, mm.bias, 1, 1, 1, 1, 1)


@pytest.mark.parametrize("name,depth", [("orientedrepoints_r50_demo", 50), ("orientedrepoints_r101_demo", 101)])
def test_reference_config_builds_resnet(name, depth):
    from orientedreppoints_b200.models import build_detector
    from orientedreppoints_b200.ops import ConvModule, DeformConv
    from orientedreppoints_b200.weights import random_state_dict
    cfg = _cfg(name)
    det = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    ref = random_state_dict(depth, seed=0, reference_init=True)
    sd = det.state_dict()
    assert set(k for k in sd if not k.endswith("num_batches_tracked")) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
        assert torch.equal(sd[k], v), k                          # init_weights() = the reference initialisation
    assert isinstance(det.neck.lateral_convs[0], ConvModule) and isinstance(det.bbox_head.reppoints_cls_conv, DeformConv)
    assert len(det.neck.fpn_convs) == 5 and det.neck.fpn_convs[3].conv.in_channels == 2048 and det.neck.fpn_convs[3].stride == (2, 2)
    assert det.test_cfg["nms"]["type"] == "rnms" and det.bbox_head.cls_out_channels == 15
    with pytest.raises(NotImplementedError):
        det(torch.zeros(1, 3, 32, 32), [dict()], return_loss=True)
    with pytest.raises(NotImplementedError):
        det.simple_test(torch.zeros(1, 3, 32, 32))               # CPU: no fallback
    with pytest.raises(KeyError):
        build_detector(dict(cfg.model, backbone=dict(type='VGG')))


def test_reference_config_builds_swin():
    from orientedreppoints_b200.models import build_detector
    from orientedreppoints_b200.swin import random_swin_state_dict
    cfg = _cfg("orientedrepoints_swin_tiny_demo")
    det = build_detector(cfg.model, test_cfg=cfg.test_cfg)
    ref = random_swin_state_dict(0)
    sd = det.state_dict()
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
