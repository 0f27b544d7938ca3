"""Minimal yacs-compatible CfgNode used ONLY to import the reference in the fixture generator."""
import copy
import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v
        self.__dict__["_frozen"] = False

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = CfgNode(value)
        self[name] = value

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and k in self and isinstance(self[k], CfgNode):
                self[k]._merge(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def freeze(self):
        pass

    def defrost(self):
        pass

    def clone(self):
        return copy.deepcopy(self)

    def dump(self, **kw):
        def plain(n):
            return {k: plain(v) if isinstance(v, dict) else v for k, v in n.items()}
        return yaml.safe_dump(plain(self), **kw)
