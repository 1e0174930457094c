"""Micro-benchmarks of the individual kernels (CUDA events, L2-exceeding inputs). Prints one JSON line per case.

    python scripts/bench_kernels.py collect [--batch 256]
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from semanticlens_b200 import ops  # noqa: E402


def time_cuda(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def peaks():
    p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def bench_collect(batch):
    pk = peaks()
    shapes = {
        "rn50.conv1": (batch, 64, 112, 112),
        "rn50.layer1": (batch, 256, 56, 56),
        "rn50.layer2": (batch, 512, 28, 28),
        "rn50.layer3": (batch, 1024, 14, 14),
        "rn50.layer4": (batch, 2048, 7, 7),
    }
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, shp in shapes.items():
        x = torch.randn(*shp, device="cuda")
        B, C = shp[:2]
        k = 20
        vals = (-torch.zeros(C, k, dtype=torch.bfloat16)).cuda()
        ids = (-torch.ones(C, k, dtype=torch.int64)).cuda()
        scratch = torch.empty(B * C, device="cuda")
        nbytes = x.numel() * 4

        def k1():
            ops.agg_reduce(x, 0, "conv")

        def k12():
            ops.agg_topk_update(x, 0, "conv", 0, 0, vals, ids, scratch)

        cand = ops.agg_reduce(x, 0, "conv")

        def k2():
            ops.topk_update(cand, vals, ids, None, 0)

        for label, fn, by in (("K1", k1, nbytes), ("K1+K2", k12, nbytes), ("K2", k2, cand.numel() * 4)):
            med, best = time_cuda(lambda: (flush.zero_() if False else None, fn()))
            print(json.dumps({
                "kernel": label, "case": name, "shape": list(shp), "ms_median": round(med, 4), "ms_best": round(best, 4),
                "GBps": round(by / med / 1e6, 1), "frac_of_measured_hbm": round(by / med / 1e6 / pk["hbm_gbs"], 3),
                "direct": os.environ.get("SLB_AGG_DIRECT", "0"),
            }), flush=True)
        del x
    # transformer layout (ViT-B/16 block output)
    x = torch.randn(batch, 197, 768, device="cuda")
    med, best = time_cuda(lambda: ops.agg_reduce(x, 0, "tokens"))
    by = x.numel() * 4
    print(json.dumps({"kernel": "K1", "case": "vit_b16.block", "shape": list(x.shape), "ms_median": round(med, 4),
                      "GBps": round(by / med / 1e6, 1), "frac_of_measured_hbm": round(by / med / 1e6 / pk["hbm_gbs"], 3)}), flush=True)
    # reference points: torch's own reduction and a plain copy on the same tensor
    x = torch.randn(batch, 256, 56, 56, device="cuda")
    y = torch.empty_like(x)
    by = x.numel() * 4
    med, _ = time_cuda(lambda: x.flatten(2).mean(-1))
    print(json.dumps({"kernel": "torch.mean", "case": "rn50.layer1", "ms_median": round(med, 4), "GBps": round(by / med / 1e6, 1)}), flush=True)
    med, _ = time_cuda(lambda: y.copy_(x))
    print(json.dumps({"kernel": "torch.copy", "case": "rn50.layer1", "ms_median": round(med, 4), "GBps_rw": round(2 * by / med / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["collect"])
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    if a.what == "collect":
        bench_collect(a.batch)
