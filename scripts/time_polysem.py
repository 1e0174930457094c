"""K8 timing split: polysemanticity at cfg-5 neuron shape for n_init = 1 and 10 (phase A + score vs the Lloyd restarts)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from semanticlens_b200 import ops  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda").manual_seed(2)
V = torch.randn(C, 256, 512, device="cuda", generator=g)
V[:, ::2] += 2 * torch.randn(C, 1, 512, device="cuda", generator=g)
for kind, X in (("planted", V), ("gaussian", torch.randn(C, 256, 512, device="cuda", generator=g))):
    for n_init in (1, 10):
        ops.polysem_2means(X, n_init=n_init)
        torch.cuda.synchronize()
        import ctypes
        from semanticlens_b200 import _native as N
        clk = (ctypes.c_uint64 * 7)()
        N.load().slb_polysem_phase_clocks(clk, 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.polysem_2means(X, n_init=n_init)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        N.load().slb_polysem_phase_clocks(clk, 1)
        per = [round(clk[c] / max(clk[6], 1) / 1.965e3, 1) for c in range(6)]  # us per neuron per CTA at 1.965 GHz
        print(json.dumps({"data": kind, "neurons": C, "n_init": n_init, "ms": round(ms, 2), "us_per_neuron": round(ms * 1e3 / C, 3),
                          "ms_at_65536": round(ms * 65536 / C, 1),
                          "us_per_neuron_per_cta": dict(zip(["gram", "means", "seed+assign", "first_sums", "lloyd", "score"], per))}), flush=True)
