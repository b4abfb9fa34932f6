"""`mmdet.utils.registry` surface (mmdet/utils/registry.py:7-79): a name -> class table per component kind and the
config-dict constructor.  Both decorator spellings of the reference work: `@R.register_module` (ResNet, FPN, the head)
and `@R.register_module()` (swin_transformer.py:449)."""
import inspect


class Registry(object):

    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __repr__(self):
        return '{}(name={}, items={})'.format(type(self).__name__, self._name, list(self._module_dict))

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def get(self, key):
        return self._module_dict.get(key)

    def _register_module(self, module_class, force=False):
        if not inspect.isclass(module_class):
            raise TypeError('module must be a class, but got {}'.format(type(module_class)))
        key = module_class.__name__
        if key in self._module_dict and not force:
            raise KeyError('{} is already registered in {}'.format(key, self.name))
        self._module_dict[key] = module_class

    def register_module(self, cls=None, force=False):
        if cls is None:                                   # called form: @R.register_module() / @R.register_module(force=True)
            return lambda c: self.register_module(c, force=force)
        self._register_module(cls, force=force)
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    """cfg['type'] (a registered name or a class) constructed with the remaining keys; default_args fill what cfg omits.
    Unknown name -> KeyError, wrong type of 'type' -> TypeError (registry.py:51-79)."""
    assert isinstance(cfg, dict) and 'type' in cfg
    assert default_args is None or isinstance(default_args, dict)
    kwargs = dict(cfg)
    kind = kwargs.pop('type')
    if isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError('{} is not in the {} registry'.format(kind, registry.name))
    elif inspect.isclass(kind):
        cls = kind
    else:
        raise TypeError('type must be a str or valid type, but got {}'.format(type(kind)))
    for k, v in (default_args or {}).items():
        kwargs.setdefault(k, v)
    return cls(**kwargs)
