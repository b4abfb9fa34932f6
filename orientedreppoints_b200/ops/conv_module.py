"""`mmdet.ops.conv_module.ConvModule` (mmdet/ops/conv_module.py:11-132): conv (+ norm) (+ activation) bundle with the
reference's attribute / parameter names (`.conv`, `.gn` / `.bn`, `.activate`), so its state_dict keys are the reference's.
Inside OrientedRepPointsDetector the bundle is executed by the fused kernels of liborp_b200 (conv epilogue + GroupNorm
statistics + apply); called on its own it runs its layers in order."""
import warnings

import torch.nn as nn

from .conv import build_conv_layer
from .norm import build_norm_layer

_ACT = {'ReLU': nn.ReLU, 'LeakyReLU': nn.LeakyReLU}


class ConvModule(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias='auto',
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, order=('conv', 'norm', 'act')):
        super(ConvModule, self).__init__()
        assert conv_cfg is None or isinstance(conv_cfg, dict)
        assert norm_cfg is None or isinstance(norm_cfg, dict)
        assert act_cfg is None or isinstance(act_cfg, dict)
        assert isinstance(order, tuple) and len(order) == 3 and set(order) == {'conv', 'norm', 'act'}
        self.conv_cfg, self.norm_cfg, self.act_cfg, self.inplace, self.order = conv_cfg, norm_cfg, act_cfg, inplace, order
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':                                   # a conv followed by a norm needs no bias
            bias = not self.with_norm
        self.with_bias = bias
        if self.with_norm and self.with_bias:
            warnings.warn('ConvModule has norm and bias at the same time')
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                     dilation=dilation, groups=groups, bias=bias)
        for a in ('in_channels', 'out_channels', 'kernel_size', 'stride', 'padding', 'dilation', 'transposed', 'output_padding', 'groups'):
            setattr(self, a, getattr(self.conv, a))
        if self.with_norm:
            nch = out_channels if order.index('norm') > order.index('conv') else in_channels
            self.norm_name, layer = build_norm_layer(norm_cfg, nch)
            self.add_module(self.norm_name, layer)
        if self.with_activation:
            spec = dict(act_cfg)
            kind = spec.pop('type')
            spec.setdefault('inplace', inplace)
            self.activate = _ACT[kind](**spec)
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name)

    def init_weights(self):
        # mmcv kaiming_init (fan_out, normal) + constant_init of the norm (conv_module.py:110-117)
        leaky = self.with_activation and self.act_cfg['type'] == 'LeakyReLU'
        if hasattr(self.conv, 'weight') and self.conv.weight is not None:
            nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='leaky_relu' if leaky else 'relu')
        if getattr(self.conv, 'bias', None) is not None:
            nn.init.constant_(self.conv.bias, 0)
        if self.with_norm:
            nn.init.constant_(self.norm.weight, 1)
            nn.init.constant_(self.norm.bias, 0)

    def forward(self, x, activate=True, norm=True):
        for step in self.order:
            if step == 'conv':
                x = self.conv(x)
            elif step == 'norm' and norm and self.with_norm:
                x = self.norm(x)
            elif step == 'act' and activate and self.with_activation:
                x = self.activate(x)
        return x
