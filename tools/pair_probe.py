"""CTA-pair probes on a B200: cta_group::2 MMA self-test (operand split, commit, timing) and the
distributed-shared-memory store rate.  Output is committed under profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neddf_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
for n in (128, 32, 64):
    for k in (16, 256):
        g = torch.Generator().manual_seed(1000 * n + k)
        a = torch.randn(256, k, generator=g); b = torch.randn(n, k, generator=g)
        ad, bd = a.to(dev), b.to(dev)
        c = torch.zeros(256, n, device=dev); cyc = torch.zeros(1, dtype=torch.int64, device=dev)
        for reps in (1, 32):
            L.check(lib.neddf_tc_pair_selftest(L.ptr(ad), L.ptr(bd), n, k, L.ptr(c), L.ptr(cyc), reps, L.stream_ptr(dev)))
            torch.cuda.synchronize()
            ref = a.double() @ b.double().T
            err = float((c.cpu().double() - ref).abs().max() / ref.abs().max())
            n_mma = reps * (k // 16) * 3
            print(f"pair MMA n={n} k={k} reps={reps} rel err {err:.2e}  cycles/MMA {cyc.item() / n_mma:.1f}", flush=True)
            if err > 1e-3:
                # which permutation of the operand halves did the hardware use?
                cc = c.cpu().double()
                for name, r2 in (("B halves swapped", torch.cat([ref[:, n // 2:], ref[:, :n // 2]], 1)),
                                 ("A halves swapped", torch.cat([ref[128:], ref[:128]], 0))):
                    e2 = float((cc - r2).abs().max() / ref.abs().max())
                    print(f"    vs {name}: {e2:.2e}")
for n_clusters in (1, 74):
    for mode in (0, 1, 2, 3, 4, 5):
        bytes_, reps = 65536, 64
        cyc = torch.zeros(2 * n_clusters, dtype=torch.int64, device=dev)
        L.check(lib.neddf_dsmem_bench(mode, reps, bytes_, n_clusters, L.ptr(cyc), L.stream_ptr(dev)))
        torch.cuda.synchronize()
        cy = cyc.cpu().double()
        eff = bytes_ // 2 if mode == 4 else bytes_  # mode 4 stores 8 of every 16 bytes
        print(f"dsmem mode={mode} clusters={n_clusters}: {eff * reps / cy.mean().item():.1f} B/clk per CTA "
              f"(min {eff * reps / cy.max().item():.1f}, max {eff * reps / cy.min().item():.1f})", flush=True)
