"""Per-convolution GEMM timings of the CLIP RN50 tower at batch B (as slb_rn_forward runs them): ms, issued TFLOP/s and
the bytes each GEMM has to move (A planes + outputs + shortcut) against the HBM roofline.
python scripts/bench_rn_convs.py [B]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from semanticlens_b200 import _native as N  # noqa: E402
from semanticlens_b200 import ops  # noqa: E402
from semanticlens_b200.foundation_models import rn  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = rn.CONFIGS["RN50"]


def time_cuda(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


shapes = {}  # (M, N, K, kind) -> count


def add(M, Nn, K, kind):
    shapes[(M, Nn, K, kind)] = shapes.get((M, Nn, K, kind), 0) + 1


w = cfg.width
M = B * 112 * 112
add(M, w // 2, 64, "relu"), add(M, w // 2, ops.conv_k(w // 2, 3), "relu"), add(M, w, ops.conv_k(w // 2, 3), "relu")
M //= 4
for _p, inpl, pl, stride, ds in rn.block_plan(cfg):
    add(M, pl, inpl, "relu")
    add(M, pl, ops.conv_k(pl, 3), "relu")
    if stride == 2:
        M //= 4
    if ds:
        add(M, 4 * pl, inpl, "short")
    add(M, 4 * pl, pl, "add_relu")

total = 0.0
for (M, Nn, K, kind), cnt in shapes.items():
    a = ops.split_planes(torch.randn(M, K, device="cuda"), scale=16.0)
    wt = ops.split_planes(torch.randn(Nn, K, device="cuda") * 0.02, scale=1024.0)
    cs, bias = torch.rand(Nn, device="cuda") + 0.5, torch.randn(Nn, device="cuda")
    o32 = torch.empty(M, Nn, device="cuda") if kind in ("short", "add_relu") else None
    opl = torch.empty(2, M, Nn, dtype=torch.float16, device="cuda") if kind != "short" else None
    res = o32 if kind == "add_relu" else None
    epi = {"relu": N.EPI_RELU, "short": N.EPI_NONE, "add_relu": N.EPI_ADD_RELU}[kind]
    for passes, tag in ((N.PASSES_SPLIT_ACC, "split_acc"), (3, "fast")):
        ms = time_cuda(lambda: ops.gemm_split(a, wt, bias=bias, col_scale=cs, residual=res, epilogue=epi, passes=passes,
                                              alpha=1 / 16384, out_f32=o32 if o32 is not None else False,
                                              out_planes=opl if opl is not None else False))
        by = 4 * M * K + 4 * Nn * K + (4 * M * Nn if opl is not None else 0) + (4 * M * Nn if o32 is not None else 0) + (4 * M * Nn if res is not None else 0)
        print(json.dumps({"M": M, "N": Nn, "K": K, "kind": kind, "mode": tag, "count": cnt, "ms": round(ms, 4),
                          "issued_TFLOPs": round(6 * M * Nn * K / ms / 1e9, 1), "GBps": round(by / ms / 1e6, 1)}), flush=True)
        if tag == "split_acc":
            total += ms * cnt
    del a, wt, o32, opl
print(json.dumps({"sum_ms_split_acc": round(total, 3), "batch": B}))
