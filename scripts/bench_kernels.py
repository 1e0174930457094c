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


def time_graph(fn, reps=10, iters=10):
    """Kernel time without the host: `reps` calls captured in one CUDA graph, replayed `iters` times (median per call).
    The Python wrapper of a small kernel costs ~30 us per call, more than the kernel itself on the small maps."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
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
    def rotating(shp):
        """enough distinct maps of this shape to exceed the L2 (one graph replay touches each once)"""
        n = max(2, -(-(600 << 20) // (4 * int(torch.tensor(shp).prod()))))
        return [torch.randn(*shp, device="cuda") for _ in range(n)]

    for name, shp in shapes.items():
        xs = rotating(shp)
        x = xs[0]
        B, C = shp[:2]
        k = 20
        vals = (-torch.zeros(C, k, dtype=torch.bfloat16)).cuda()
        ids = (-torch.ones(C, k, dtype=torch.int64)).cuda()
        scratch = torch.empty(B * C, device="cuda")
        nbytes = x.numel() * 4
        it = [0]

        def nxt():
            it[0] = (it[0] + 1) % len(xs)
            return xs[it[0]]

        def k1():
            ops.agg_reduce(nxt(), 0, "conv")

        def k12():
            ops.agg_topk_update(nxt(), 0, "conv", 0, 0, vals, ids, scratch)

        cand = ops.agg_reduce(x, 0, "conv")

        def k2():
            ops.topk_update(cand, vals, ids, None, 0)

        for label, fn, by in (("K1", k1, nbytes), ("K1+K2", k12, nbytes), ("K2", k2, cand.numel() * 4)):
            med, best = time_graph(fn, reps=len(xs))
            print(json.dumps({
                "kernel": label, "case": name, "shape": list(shp), "ms_median": round(med, 4), "ms_best": round(best, 4),
                "GBps": round(by / med / 1e6, 1), "frac_of_measured_hbm": round(by / med / 1e6 / pk["hbm_gbs"], 3),
                "timing": f"CUDA graph of {len(xs)} calls over {len(xs)} distinct maps",
            }), flush=True)
        del xs, x
    # (B, T, F) layout: transformer block outputs and the channels-last maps of the accelerated probed forward
    for case, shp in (("vit_b16.block", (batch, 197, 768)), ("vit_l14.block", (batch // 2, 257, 1024)),
                      ("cl.conv1", (batch // 2, 12544, 64)), ("cl.layer1.conv1", (batch // 2, 3136, 64)),
                      ("cl.layer1", (batch // 2, 3136, 256)), ("cl.layer2", (batch // 2, 784, 512)),
                      ("cl.layer3", (batch // 2, 196, 1024)), ("cl.layer4", (batch // 2, 49, 2048))):
        xs = rotating(shp)
        it = [0]

        def kb():
            it[0] = (it[0] + 1) % len(xs)
            ops.agg_reduce(xs[it[0]], 0, "tokens")

        med, best = time_graph(kb, reps=len(xs))
        by = xs[0].numel() * 4
        print(json.dumps({"kernel": "K1", "case": case, "shape": list(shp), "ms_median": round(med, 4),
                          "GBps": round(by / med / 1e6, 1), "frac_of_measured_hbm": round(by / med / 1e6 / pk["hbm_gbs"], 3),
                          "timing": f"CUDA graph of {len(xs)} calls over {len(xs)} distinct maps"}), flush=True)
        del xs
    # reference points: torch's own reduction and a plain copy on the same tensor
    x = torch.randn(batch, 256, 56, 56, device="cuda")
    y = torch.empty_like(x)
    by = x.numel() * 4
    med, _ = time_cuda(lambda: x.flatten(2).mean(-1))
    print(json.dumps({"kernel": "torch.mean", "case": "rn50.layer1", "ms_median": round(med, 4), "GBps": round(by / med / 1e6, 1)}), flush=True)
    med, _ = time_cuda(lambda: y.copy_(x))
    print(json.dumps({"kernel": "torch.copy", "case": "rn50.layer1", "ms_median": round(med, 4), "GBps_rw": round(2 * by / med / 1e6, 1)}), flush=True)


def bench_gemm(batch):
    """K4 on the GEMM shapes of the ViT towers (rows = batch * tokens) and K6 at cfg-5 size.
    SLB_GEMM_KERNEL=single|pair128|pair256 forces a kernel variant, SLB_BENCH_ONLY=<substring> selects cases."""
    pk = peaks()
    from semanticlens_b200 import _native as N

    cases = {
        "vitb32.qkv": (batch * 50, 2304, 768), "vitb32.out": (batch * 50, 768, 768),
        "vitb32.fc": (batch * 50, 3072, 768), "vitb32.proj": (batch * 50, 768, 3072),
        "vitl14.qkv": (64 * 257, 3072, 1024), "vitl14.fc": (64 * 257, 4096, 1024), "vitl14.proj": (64 * 257, 1024, 4096),
        "square.8192": (8192, 8192, 8192),
    }
    import time

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from bench import Clocks

    only = os.environ.get("SLB_BENCH_ONLY")
    for name, (M, Nn, K) in cases.items():
        if only and only not in name:
            continue
        a = ops.split_planes(torch.randn(M, K, device="cuda"), scale=16.0)
        w = ops.split_planes(torch.randn(Nn, K, device="cuda") * 0.02, scale=1024.0)
        out = torch.empty(M, Nn, device="cuda")
        for passes in (3, 1):
            med, best = time_cuda(lambda: ops.gemm_split(a, w, passes=passes, out_f32=out))
            tf = 2 * M * Nn * K * passes / med / 1e9
            print(json.dumps({"kernel": "K4 gemm_split", "case": name, "M": M, "N": Nn, "K": K, "passes": passes,
                              "ms_median": round(med, 4), "ms_best": round(best, 4), "issued_TFLOPs": round(tf, 1),
                              "fp32_equiv_TFLOPs": round(tf / passes, 1),
                              "frac_of_measured_bf16": round(tf / pk["bf16_tflops"], 3)}), flush=True)
        if name == "square.8192":
            # sustained: ~1.5 s back to back with the SM clock sampled, to tell pipe utilisation from power capping
            clk = Clocks(0)
            time.sleep(0.3)
            n = int(1500 / med)
            torch.cuda.synchronize()
            clk.begin()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                ops.gemm_split(a, w, passes=3, out_f32=out)
            e1.record()
            torch.cuda.synchronize()
            clk.end()
            ms = e0.elapsed_time(e1) / n
            c = clk.stop()
            tf = 2 * M * Nn * K * 3 / ms / 1e9
            mhz = c.get("sm_mhz") or 0
            print(json.dumps({"kernel": "K4 gemm_split", "case": name + " sustained", "ms": round(ms, 4), "issued_TFLOPs": round(tf, 1),
                              "clocks": c, "tensor_pipe_util_at_clock": round(tf * 1e12 / (148 * 8192 * mhz * 1e6), 3) if mhz else None}),
                  flush=True)
        del a, w, out
    x, y = torch.randn(10000, 512, device="cuda"), torch.randn(65536, 512, device="cuda")
    med, best = time_cuda(lambda: ops.cosine_gemm(x, y), warmup=2, iters=5)
    fl = 2 * 10000 * 65536 * 512 * 3
    print(json.dumps({"kernel": "K6 cosine_gemm", "case": "cfg5 10000x65536x512", "ms_median": round(med, 3),
                      "issued_TFLOPs": round(fl / med / 1e9, 1), "out_write_GBps": round(10000 * 65536 * 4 / med / 1e6, 1)}), flush=True)


def bench_scores():
    pk = peaks()
    V = torch.randn(4096, 256, 512, device="cuda")  # 2.1 GB slab of the cfg-5 concept DB (larger than L2)
    med, best = time_cuda(lambda: ops.clarity(V), warmup=2, iters=5)
    by = V.numel() * 4
    print(json.dumps({"kernel": "K7 clarity", "case": "4096x256x512", "ms_median": round(med, 3), "GBps": round(by / med / 1e6, 1),
                      "frac_of_measured_hbm": round(by / med / 1e6 / pk["hbm_gbs"], 3)}), flush=True)
    Vp = V[:1184]
    med, best = time_cuda(lambda: ops.polysem_2means(Vp), warmup=1, iters=3)
    by = Vp.numel() * 4
    print(json.dumps({"kernel": "K8 polysem_2means", "case": "1184x256x512", "ms_median": round(med, 3),
                      "us_per_neuron": round(med * 1e3 / 1184, 2), "GBps": round(by / med / 1e6, 1),
                      "frac_of_measured_hbm": round(by / med / 1e6 / pk["hbm_gbs"], 4)}), flush=True)
    V20 = torch.randn(3904, 20, 512, device="cuda")
    med, best = time_cuda(lambda: ops.polysem_2means(V20), warmup=1, iters=3)
    print(json.dumps({"kernel": "K8 polysem_2means", "case": "cfg2 db 3904x20x512", "ms_median": round(med, 3),
                      "us_per_neuron": round(med * 1e3 / 3904, 2)}), flush=True)


def bench_embed(batch):
    from semanticlens_b200.foundation_models import OpenClip

    for url, B in (("ViT-B-32", batch), ("ViT-B-16", 128), ("ViT-L-14", 64), ("ViT-B-16-SigLIP2", 128), ("ViT-L-16-SigLIP-256", 64),
                   ("RN50", 128)):
        if os.environ.get("SLB_BENCH_ONLY") and os.environ["SLB_BENCH_ONLY"] not in url:
            continue
        fm = OpenClip(url, device="cuda", load_weights=False, seed=1)
        S = fm.cfg.image_size
        u8 = torch.randint(0, 255, (B, 3, S, S), dtype=torch.uint8, device="cuda")
        med, best = time_cuda(lambda: fm.encode_image(fm.preprocess(u8)), warmup=2, iters=5)
        print(json.dumps({"kernel": "embed tower", "case": url, "batch": B, "ms_median": round(med, 3),
                          "images_per_s": round(B / med * 1e3, 1)}), flush=True)
        del fm


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["collect", "gemm", "scores", "embed"])
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    if a.what == "collect":
        bench_collect(a.batch)
    elif a.what == "gemm":
        bench_gemm(a.batch)
    elif a.what == "scores":
        bench_scores()
    elif a.what == "embed":
        bench_embed(a.batch)
