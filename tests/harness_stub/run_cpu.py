"""CPU harness runner (TEST INFRASTRUCTURE): `python run_cpu.py <reference script> [args]` = `python -m
neddf_b200.launch ...` on a box without a GPU.  The only thing it adds: torch.load maps CUDA storages of the
reference's checkpoint to the CPU (base_trainer.py:121 calls torch.load without map_location, which cannot work
without a CUDA device whatever renderer is bound)."""
import sys

import torch

_load = torch.load


def _load_cpu(*a, **k):
    if not torch.cuda.is_available():
        k.setdefault("map_location", "cpu")
    return _load(*a, **k)


torch.load = _load_cpu

from neddf_b200.launch import main  # noqa: E402

if __name__ == "__main__":
    sys.argv = ["neddf_b200.launch"] + sys.argv[1:]
    main()
