"""Test-only stand-in for fvcore.common.config.CfgNode (a yacs-style attribute dict).

Only the surface the reference touches while building a model is provided: attribute access,
clone/freeze/defrost, yaml loading with _BASE_ inheritance, merge_from_other_cfg, merge_from_list.
Written from the documented behaviour of yacs; not a copy of fvcore.
"""
import copy
import os
from ast import literal_eval

import yaml

BASE_KEY = "_BASE_"


class CfgNode(dict):
    IMMUTABLE = "__immutable__"
    NEW_ALLOWED = "__new_allowed__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        init_dict = {} if init_dict is None else init_dict
        super().__init__()
        for k, v in init_dict.items():
            if isinstance(v, dict) and not isinstance(v, CfgNode):
                v = type(self)(v)
            dict.__setitem__(self, k, v)
        self.__dict__[CfgNode.IMMUTABLE] = False
        self.__dict__[CfgNode.NEW_ALLOWED] = new_allowed

    # attribute access -------------------------------------------------------------------------
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = type(self)(value)
        self[name] = value

    # freezing ---------------------------------------------------------------------------------
    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def _immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def is_new_allowed(self):
        return self.__dict__[CfgNode.NEW_ALLOWED]

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        new = type(self)()
        for k, v in self.items():
            dict.__setitem__(new, k, copy.deepcopy(v, memo))
        new.__dict__[CfgNode.IMMUTABLE] = self.__dict__[CfgNode.IMMUTABLE]
        new.__dict__[CfgNode.NEW_ALLOWED] = self.__dict__[CfgNode.NEW_ALLOWED]
        return new

    # yaml -------------------------------------------------------------------------------------
    @classmethod
    def _open_cfg(cls, filename):
        return open(filename, "r")

    @classmethod
    def load_yaml_with_base(cls, filename, allow_unsafe=False):
        with cls._open_cfg(filename) as f:
            cfg = yaml.unsafe_load(f) if allow_unsafe else yaml.safe_load(f)

        def merge_a_into_b(a, b):
            for k, v in a.items():
                if isinstance(v, dict) and k in b:
                    assert isinstance(b[k], dict), f"Cannot inherit key '{k}' from base!"
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        if BASE_KEY in cfg:
            base = cfg.pop(BASE_KEY)
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not any(map(base.startswith, ["/", "https://", "http://"])):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = cls.load_yaml_with_base(base, allow_unsafe=allow_unsafe)
            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    # merging ----------------------------------------------------------------------------------
    @staticmethod
    def _coerce(new, old, key):
        if isinstance(new, str):
            try:
                new = literal_eval(new)
            except (ValueError, SyntaxError):
                pass
        if old is None or new is None or type(new) is type(old):
            return new
        if isinstance(old, tuple) and isinstance(new, list):
            return tuple(new)
        if isinstance(old, list) and isinstance(new, tuple):
            return list(new)
        if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
            return float(new)
        if isinstance(old, CfgNode) and isinstance(new, dict):
            return new
        raise ValueError(f"Type mismatch for {key}: {type(old)} vs {type(new)} ({old!r} vs {new!r})")

    def merge_from_other_cfg(self, other):
        def merge(a, b, path):
            for k, v in a.items():
                full = ".".join(path + [k])
                if k in b:
                    if isinstance(b[k], CfgNode) and isinstance(v, dict):
                        merge(v, b[k], path + [k])
                    else:
                        if isinstance(v, dict) and not isinstance(v, CfgNode):
                            v = type(self)(v)
                        dict.__setitem__(b, k, CfgNode._coerce(copy.deepcopy(v), b[k], full))
                elif b.is_new_allowed():
                    if isinstance(v, dict) and not isinstance(v, CfgNode):
                        v = type(self)(v, new_allowed=True)
                    dict.__setitem__(b, k, copy.deepcopy(v))
                else:
                    raise KeyError(f"Non-existent config key: {full}")

        if self.is_frozen():
            raise AttributeError("cfg is frozen")
        merge(other, self, [])

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0
        if self.is_frozen():
            raise AttributeError("cfg is frozen")
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            keys = full_key.split(".")
            d = self
            for sub in keys[:-1]:
                assert sub in d, f"Non-existent key: {full_key}"
                d = d[sub]
            assert keys[-1] in d, f"Non-existent key: {full_key}"
            dict.__setitem__(d, keys[-1], CfgNode._coerce(v, d[keys[-1]], full_key))

    def dump(self, **kwargs):
        def to_dict(n):
            return {k: to_dict(v) if isinstance(v, CfgNode) else v for k, v in n.items()}

        return yaml.safe_dump(to_dict(self), **kwargs)
