import importlib

from omegaconf import DictConfig, ListConfig


_original_cwd = None


def get_original_cwd():
    assert _original_cwd is not None, "only valid inside a @hydra.main task"
    return _original_cwd


def _locate(target: str):
    mod, _, name = target.rpartition(".")
    return getattr(importlib.import_module(mod), name)  # attribute lookup at call time, like hydra's _locate


def _build(node, recursive):
    if recursive and isinstance(node, DictConfig) and "_target_" in node:
        return instantiate(node)
    if recursive and isinstance(node, ListConfig):
        return [_build(v, recursive) for v in node]
    return node


def instantiate(config, *args, **kwargs):
    """hydra.utils.instantiate: `_target_(*args, **config, **kwargs)`; `_recursive_` (kwarg or config key, default
    True) decides whether nested `_target_` nodes are instantiated or passed through as config nodes."""
    recursive = kwargs.pop("_recursive_", config.get("_recursive_", True))
    kwargs.pop("_convert_", None)
    params = {k: v for k, v in config.items() if k not in ("_target_", "_recursive_", "_convert_", "_partial_")}
    params.update(kwargs)
    params = {k: _build(v, recursive) for k, v in params.items()}
    return _locate(config["_target_"])(*args, **params)
