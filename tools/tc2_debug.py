"""Per-layer diagnosis of a tensor-core engine against the fp32 engine: runs the training forward
(which saves every layer's pre-activations [layer][sample][row type][channel]) with both engines on
the same samples and prints where they first differ, split by sample block / row type / channel half."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import neddf_b200
from neddf_b200 import _lib as L
from oracle import neddf_oracle as orc
from tests.helpers import Case
import tests.gpu_util as G

engine = sys.argv[1] if len(sys.argv) > 1 else "tc2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
c = Case("bunny")
render = G.build_render(c, engine)
net = render.network_fine
d, o = orc.make_rays(c.t("uv"), c.cam)
pos, dd, var = orc.make_samples(c.rc, d, o, c.t("dists_fine"))
pos = pos.reshape(-1, 3)[:n].contiguous().to(G.DEV)
dd = dd.reshape(-1, 3)[:n].contiguous().to(G.DEV)
var = var.reshape(-1, 3)[:n].contiguous().to(G.DEV)
n_hidden = (net.ddf_layer_count - 1) + (net.col_layer_count - 1)


def run(eng):
    h = net._field(G.DEV)
    st = net._state_struct()
    save = torch.zeros(n_hidden, n, 4, 256, device=G.DEV)
    outs = [torch.zeros(n, device=G.DEV) for _ in range(2)] + [torch.zeros(n, 3, device=G.DEV)] + [torch.zeros(n, device=G.DEV) for _ in range(2)]
    L.check(L.lib().neddf_field_forward_train_samples(h, C.byref(st), L.ptr(pos), L.ptr(dd), L.ptr(var), n, L.ptr(outs[0]),
            L.ptr(outs[1]), L.ptr(outs[2]), L.ptr(outs[3]), L.ptr(outs[4]), L.ptr(save), L.ENGINE_IDS[eng], L.stream_ptr(G.DEV)))
    torch.cuda.synchronize()
    return save.cpu(), [t.cpu() for t in outs]


ref, ro = run("fp32")
got, go = run(engine)
for l in range(n_hidden):
    e = (got[l] - ref[l]).abs()
    scale = float(ref[l].abs().max())
    print(f"layer {l}: max err {float(e.max()) / scale:.2e} (scale {scale:.2e})")
    if float(e.max()) / scale > 1e-3:
        for blk in range(0, n, 16):
            eb = e[blk:blk + 16]
            print(f"   samples {blk:4d}..{blk + 15:4d}: " + " ".join(
                f"j{j}:[{float(eb[:, j, :128].max()) / scale:.1e},{float(eb[:, j, 128:].max()) / scale:.1e}]" for j in range(4)))
        break
for name, a, b in zip(("distance", "density", "color", "penalty", "aux"), go, ro):
    print(name, float((a - b).abs().max() / b.abs().max()))
torch.set_printoptions(precision=4, linewidth=200)
for (s, j) in ((0, 0), (0, 1), (17, 0), (40, 2)):
    if s < n:
        print(f"sample {s} type {j}: got", got[0][s, j, :6], got[0][s, j, 128:134])
        print(f"sample {s} type {j}: ref", ref[0][s, j, :6], ref[0][s, j, 128:134])
# does got match ref under a permutation of samples / types?  correlate got rows with ref rows
g0 = got[0].reshape(-1, 256)[:, :128]
r0 = ref[0].reshape(-1, 256)[:, :128]
gn = g0 / (g0.norm(dim=1, keepdim=True) + 1e-9)
rn = r0 / (r0.norm(dim=1, keepdim=True) + 1e-9)
corr = gn @ rn.T
best = corr.argmax(1)
print("best-matching ref row (sample*4+type) for got rows 0..15:", best[:16].tolist(), "corr", [round(float(corr[i, best[i]]), 3) for i in range(16)])
print("rows 128..143:", best[128:144].tolist())
os.makedirs("gpurun_out", exist_ok=True)
torch.save({"got0": got[0].clone(), "ref0": ref[0].clone(), "got1": got[1].clone(), "ref1": ref[1].clone(),
            "pos": pos.cpu(), "dir": dd.cpu(), "var": var.cpu()}, "gpurun_out/tc2_debug.pt")
