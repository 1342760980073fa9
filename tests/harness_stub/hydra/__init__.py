"""Test-harness stand-in for hydra-core (not installed in this image, no network) - TEST INFRASTRUCTURE.

Implements what the reference's entry points use (neddf/scripts/run_eval.py:23-34, nerf_trainer.py:32-36,
base_trainer.py:97-113): initialize_config_dir + compose of an already composed `.hydra/config.yaml` with dotted
overrides, GlobalHydra.instance().clear(), and utils.instantiate with Hydra's semantics for `_recursive_` (nested
`_target_` nodes are built only when it is true; with false they reach the callee as DictConfig / ListConfig).
NEDDF_HARNESS_OVERRIDES (comma separated `a.b=value`) appends overrides - the CPU harness test uses it to point
`trainer.device` at the CPU, which run_eval.py itself offers no switch for."""
import os

import yaml
from omegaconf import DictConfig, ListConfig, OmegaConf

from . import utils  # noqa: F401

_state = {"config_dir": None}


def initialize_config_dir(config_dir, version_base=None, job_name=None):
    assert os.path.isabs(config_dir) and os.path.isdir(config_dir), config_dir
    _state["config_dir"] = config_dir


def compose(config_name="config", overrides=()):
    assert _state["config_dir"] is not None, "hydra.initialize_config_dir was not called"
    with open(os.path.join(_state["config_dir"], config_name + ".yaml")) as f:
        cfg = OmegaConf.create(yaml.safe_load(f))
    extra = [o for o in os.environ.get("NEDDF_HARNESS_OVERRIDES", "").split(",") if o]
    for ov in list(overrides) + extra:
        key, _, val = ov.partition("=")
        OmegaConf.update(cfg, key, yaml.safe_load(val))
    return cfg


def _compose_tree(config_dir, config_name, overrides):
    """config.yaml with a defaults list of `group: option` entries (config/config.yaml of the reference), group
    overrides `group=option` and value overrides `a.b=value`."""
    with open(os.path.join(config_dir, config_name + ".yaml")) as f:
        root = yaml.safe_load(f) or {}
    defaults = root.pop("defaults", [])
    choices = {}
    for d in defaults:
        if isinstance(d, dict):
            choices.update(d)
    value_ov = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        if key in choices and "." not in key:
            choices[key] = val
        else:
            value_ov.append((key, val))
    cfg = OmegaConf.create(root)
    for group, option in choices.items():
        with open(os.path.join(config_dir, group, str(option) + ".yaml")) as f:
            cfg[group] = yaml.safe_load(f)
    for key, val in value_ov:
        OmegaConf.update(cfg, key, yaml.safe_load(val))
    return cfg


def main(config_path=None, config_name="config", version_base=None):
    """@hydra.main: compose the config tree next to the decorated function's file, apply the command-line overrides,
    run the task from a fresh working directory (hydra.run.dir; here NEDDF_HARNESS_RUN_DIR or a temp dir)."""
    import functools
    import inspect
    import sys
    import tempfile

    def deco(fn):
        @functools.wraps(fn)
        def wrapper():
            src_dir = os.path.dirname(os.path.abspath(inspect.getsourcefile(fn)))
            config_dir = os.path.normpath(os.path.join(src_dir, config_path))
            extra = [o for o in os.environ.get("NEDDF_HARNESS_OVERRIDES", "").split(",") if o]
            cfg = _compose_tree(config_dir, config_name, [a for a in sys.argv[1:] if "=" in a] + extra)
            utils._original_cwd = os.getcwd()
            run_dir = os.environ.get("NEDDF_HARNESS_RUN_DIR") or tempfile.mkdtemp(prefix="hydra_run_")
            os.makedirs(run_dir, exist_ok=True)
            os.chdir(run_dir)
            try:
                return fn(cfg)
            finally:
                os.chdir(utils._original_cwd)
        return wrapper
    return deco


__all__ = ["initialize_config_dir", "compose", "main", "utils", "DictConfig", "ListConfig"]
