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


__all__ = ["initialize_config_dir", "compose", "utils", "DictConfig", "ListConfig"]
