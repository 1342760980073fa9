"""The reference's own evaluation entry point, unmodified, against the B200 classes - as far as a box without a GPU
can take it.

`python -m neddf_b200.launch neddf/scripts/run_eval.py <run dir>` is executed in a subprocess from /root/reference
with Hydra / omegaconf / scikit-image stand-ins (tests/harness_stub: DictConfig and ListConfig are NOT dict / list
subclasses, `instantiate` honours `_recursive_=False`, like the real packages that are not installable here).  The
script composes the saved `.hydra/config.yaml`, Hydra instantiates the reference's NeRFTrainer, which instantiates
`neddf.render.NeRFRender` with the un-instantiated network node (nerf_trainer.py:32-36) - rebound to neddf_b200 by
the launcher -, builds Adam over get_parameters_list(), loads the real checkpoint (base_trainer.py:121, strict) and
calls render_all -> render_image.  That last call is the kernel boundary: on this GPU-less box it must fail loudly
from neddf_b200 (no CPU fallback).  /root/reference does not exist on the GPU box, so the GPU half of this harness
cannot run there; the image path itself is parity-tested on the GPU against the reference's image golden."""
import os
import shutil
import subprocess
import sys

import pytest
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pretrained", "bunny_smoke")), reason="needs /root/reference")
def test_run_eval_reaches_the_kernel_boundary(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("CPU harness: on a GPU box the run would render 50 full frames")
    run_dir = tmp_path / "bunny_smoke"
    shutil.copytree(os.path.join(REF, "pretrained", "bunny_smoke"), run_dir)
    for root, dirs, files in os.walk(run_dir):  # the reference tree is read-only
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "harness_stub"), REF, REPO])
    env["NEDDF_HARNESS_OVERRIDES"] = "trainer.device=cpu"  # the saved config says cuda:0; run_eval.py has no switch
    r = subprocess.run([sys.executable, os.path.join(HERE, "harness_stub", "run_cpu.py"), "neddf/scripts/run_eval.py", str(run_dir)],
                       cwd=REF, env=env, capture_output=True, text=True, timeout=900)
    log = r.stdout + r.stderr
    assert r.returncode != 0, log[-2000:]
    assert "rendering from camera 0" in r.stdout, log[-2000:]           # base_trainer.py:187: trainer built, checkpoint loaded
    assert "neddf/trainer/base_trainer.py" in r.stderr and "neddf_b200/render.py" in r.stderr, log[-2000:]  # their caller, our class
    assert "neddf_b200.NeRFRender renders on CUDA devices only" in r.stderr, log[-2000:]  # loud, no CPU fallback
    assert (run_dir / "eval").is_dir()  # run_eval.py:41-42 created its output directory before rendering


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "data", "bunny_smoke")), reason="needs /root/reference")
@pytest.mark.parametrize("argv,patch", [
    ([], "0"),                                                                                    # config/config.yaml defaults: NeDDF
    (["network=nerf", "render=nerf_render", "loss=nerf_loss", "trainer=nerf_trainer"], "1"),  # NeRF groups + trainer glue
], ids=["neddf-defaults", "nerf-groups-patched-trainer"])
def test_run_train_reaches_the_kernel_boundary(tmp_path, argv, patch):
    """The reference's TRAINING entry point, unmodified (`neddf/scripts/run.py`, @hydra.main over config/config.yaml
    with its defaults list): dataset, cameras, loss modules, NeRFTrainer, Adam, `run_train` -> `run_train_step` ->
    `neural_render.render_rays(uv, camera)` on the B200 class, which refuses the CPU tensors loudly.  With
    NEDDF_B200_PATCH_TRAINER=1 the launcher also rebinds the loss classes / ground-truth gather / render_all that
    Hydra then instantiates from the `loss` group."""
    if torch.cuda.is_available():
        pytest.skip("CPU harness: on a GPU box this would start a 2000-epoch training run")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "harness_stub"), REF, REPO])
    env["NEDDF_HARNESS_OVERRIDES"] = "trainer.device=cpu"
    env["NEDDF_HARNESS_RUN_DIR"] = str(tmp_path / "run")
    env["NEDDF_B200_PATCH_TRAINER"] = patch
    r = subprocess.run([sys.executable, os.path.join(HERE, "harness_stub", "run_cpu.py"), "neddf/scripts/run.py"] + argv,
                       cwd=REF, env=env, capture_output=True, text=True, timeout=900)
    log = r.stdout + r.stderr
    assert r.returncode != 0, log[-2000:]
    assert "epoch:  0" in r.stdout, log[-2000:]                                   # nerf_trainer.py:58
    assert "nerf_trainer.py" in r.stderr and "run_train_step" in r.stderr, log[-2000:]
    assert "neddf_b200/render.py" in r.stderr and "must live on the CUDA device" in r.stderr, log[-2000:]
    assert (tmp_path / "run" / "models").is_dir()                                 # nerf_trainer.py:52, in hydra's run dir


def test_constructors_accept_hydra_config_nodes():
    """`_recursive_=False` hands NeRFRender the network node as a DictConfig whose `skips` / `penalty_weight` are
    ListConfig / DictConfig (not list / dict); the render node's scalars arrive as keyword arguments."""
    sys.path.insert(0, os.path.join(HERE, "harness_stub"))
    try:
        for m in [k for k in sys.modules if k.split(".")[0] in ("omegaconf", "hydra")]:
            del sys.modules[m]
        import hydra
        from omegaconf import DictConfig, ListConfig, OmegaConf
        import neddf_b200
        cfg = OmegaConf.create({
            "render": {"_target_": "neddf_b200.NeRFRender", "sample_coarse": 32, "sample_fine": 48, "use_coarse_network": True,
                       "sampling_type": "cone"},
            "network": {"_target_": "neddf.network.NeDDF", "col_layer_count": 3, "skips": [2, 5], "penalty_weight": {"range_color": 0.25}},
        })
        assert isinstance(cfg.network, DictConfig) and not isinstance(cfg.network, dict)
        assert isinstance(cfg.network.skips, ListConfig) and not isinstance(cfg.network.skips, list)
        r = hydra.utils.instantiate(cfg.render, network_config=cfg.network, _recursive_=False)
        assert isinstance(r, neddf_b200.NeRFRender) and r.sample_coarse == 32 and r.network_coarse is not r.network_fine
        assert r.network_fine.skips == [2, 5] and r.network_fine.penalty_weight == {"range_color": 0.25}
        assert len(r.network_fine.layers_col) == 2
        for target, cls in (("neddf.network.NeRF", neddf_b200.NeRF), ("neddf.network.NeuS", neddf_b200.NeuS)):
            node = OmegaConf.create({"_target_": target, "skips": [1]})
            rv = hydra.utils.instantiate(cfg.render, network_config=node, _recursive_=False)
            assert isinstance(rv.network_fine, cls) and rv.network_fine.skips == [1]
    finally:
        sys.path.remove(os.path.join(HERE, "harness_stub"))
        for m in [k for k in sys.modules if k.split(".")[0] in ("omegaconf", "hydra", "skimage")]:
            del sys.modules[m]
