"""Test-harness stand-in for omegaconf (not installed in this image, no network) - TEST INFRASTRUCTURE.

Closer to the real thing than tests/golden/_refstub where it matters for the drop-in boundary: DictConfig is a
MutableMapping with attribute access and ListConfig a MutableSequence - NEITHER is a dict / list subclass, exactly
like omegaconf's - so code that only works on plain containers (isinstance checks, json, `cfg.pop`) fails here the
way it would under genuine Hydra."""
from collections.abc import MutableMapping, MutableSequence


def _wrap(v):
    if isinstance(v, (DictConfig, ListConfig)):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(v)
    return v


class DictConfig(MutableMapping):
    def __init__(self, content=None):
        object.__setattr__(self, "_content", {k: _wrap(v) for k, v in dict(content or {}).items()})

    def __getitem__(self, k):
        return self._content[k]

    def __setitem__(self, k, v):
        self._content[k] = _wrap(v)

    def __delitem__(self, k):
        del self._content[k]

    def __iter__(self):
        return iter(self._content)

    def __len__(self):
        return len(self._content)

    def __getattr__(self, k):
        try:
            return self._content[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self._content[k] = _wrap(v)

    def __repr__(self):
        return f"DictConfig({OmegaConf.to_container(self)!r})"


class ListConfig(MutableSequence):
    def __init__(self, content=()):
        self._content = [_wrap(v) for v in content]

    def __getitem__(self, i):
        return self._content[i]

    def __setitem__(self, i, v):
        self._content[i] = _wrap(v)

    def __delitem__(self, i):
        del self._content[i]

    def __len__(self):
        return len(self._content)

    def insert(self, i, v):
        self._content.insert(i, _wrap(v))

    def __repr__(self):
        return f"ListConfig({OmegaConf.to_container(self)!r})"


class OmegaConf:
    @staticmethod
    def create(obj=None):
        return _wrap({} if obj is None else obj)

    @staticmethod
    def to_container(cfg, resolve=True):
        if isinstance(cfg, DictConfig):
            return {k: OmegaConf.to_container(v) for k, v in cfg.items()}
        if isinstance(cfg, ListConfig):
            return [OmegaConf.to_container(v) for v in cfg]
        return cfg

    @staticmethod
    def update(cfg, dotted, value):
        keys = dotted.split(".")
        node = cfg
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = value
