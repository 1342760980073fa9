import importlib


def instantiate(config, *args, **kwargs):
    """Resolve ``_target_`` and call it with the remaining keys (non-recursive)."""
    cfg = dict(config)
    kwargs.pop("_recursive_", None)
    target = cfg.pop("_target_")
    mod, _, name = target.rpartition(".")
    cls = getattr(importlib.import_module(mod), name)
    cfg.update(kwargs)
    return cls(*args, **cfg)
