"""Tiny render_rays for compute-sanitizer runs (fp32 engine + K1/K3/K4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_parity import _bench_render
engine = sys.argv[1] if len(sys.argv) > 1 else "fp32"
r, cam = _bench_render(engine)
uv = torch.randint(300, 500, (40, 2)).cuda()
with torch.no_grad():
    o = r.render_rays(uv, cam)
torch.cuda.synchronize()
print(engine, "path finite:", bool(torch.isfinite(o["color"]).all()))
