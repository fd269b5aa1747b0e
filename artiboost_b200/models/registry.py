"""Registry + builders with the reference's semantics (anakin/utils/registry.py:4-69, anakin/utils/builder.py:5-30,
85-100): `build_from_cfg` pops TYPE, fills defaults, calls cls(**args); unknown TYPE raises KeyError."""
import functools
import inspect


class Registry(object):

    def __init__(self, name):
        self._name = name
        self._module_dict = dict()

    def __repr__(self):
        return "{}(name={}, items={})".format(self.__class__.__name__, self._name, list(self._module_dict.keys()))

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key, None)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError("module must be a class, but got {}".format(type(cls)))
        if cls.__name__ in self._module_dict:
            raise KeyError("{} is already registered in {}".format(cls.__name__, self.name))
        self._module_dict[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    assert isinstance(cfg, dict) and "TYPE" in cfg
    assert isinstance(default_args, dict) or default_args is None
    args = cfg.copy()
    obj_type = args.pop("TYPE")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError("{} is not in the {} registry".format(obj_type, registry.name))
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError("type must be a str or valid type, but got {}".format(type(obj_type)))
    if default_args is not None:
        for name, value in default_args.items():
            args.setdefault(name, value)
    return obj_cls(**args)


MODEL = Registry("model")
BACKBONE = Registry("backbone")
HEAD = Registry("head")


def build_backbone(cfg, default_args=None):
    return build_from_cfg(cfg, BACKBONE, default_args)


def build_head(cfg, default_args=None):
    return build_from_cfg(cfg, HEAD, default_args)


def build_model(cfg, default_args=None):
    return build_from_cfg(cfg, MODEL, default_args)


def build_arch_model_list(cfg, preset_cfg, **kwargs):
    default_args = {"DATA_PRESET": preset_cfg}
    default_args.update(kwargs)
    if isinstance(cfg, list):
        return [build_model(c, default_args) for c in cfg]
    return [build_model(cfg, default_args)]


def enable_lower_param(func):
    """anakin/utils/misc.py:30-38: keyword names are upper-cased before the call."""

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        return func(*args, **{k.upper(): v for k, v in kwargs.items()})

    return wrapper
