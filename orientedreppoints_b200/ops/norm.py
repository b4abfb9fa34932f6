"""`mmdet.ops.norm` surface (mmdet/ops/norm.py:3-55): (name, layer) for 'BN' / 'SyncBN' / 'GN' configs; eps defaults to 1e-5,
`requires_grad` is applied to the layer's parameters, the returned name is the abbreviation + postfix ('bn1', 'gn')."""
import torch.nn as nn

norm_cfg = {
    'BN': ('bn', nn.BatchNorm2d),
    'SyncBN': ('bn', nn.SyncBatchNorm),
    'GN': ('gn', nn.GroupNorm),
}


def build_norm_layer(cfg, num_features, postfix=''):
    assert isinstance(cfg, dict) and 'type' in cfg
    spec = dict(cfg)
    kind = spec.pop('type')
    if kind not in norm_cfg:
        raise KeyError('Unrecognized norm type {}'.format(kind))
    abbr, cls = norm_cfg[kind]
    assert isinstance(postfix, (int, str))
    trainable = spec.pop('requires_grad', True)
    spec.setdefault('eps', 1e-5)
    if kind == 'GN':
        assert 'num_groups' in spec
        layer = cls(num_channels=num_features, **spec)
    else:
        layer = cls(num_features, **spec)
    for p in layer.parameters():
        p.requires_grad = trainable
    return abbr + str(postfix), layer
