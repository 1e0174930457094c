"""Time the long-sequence attention kernel on the towers' shapes: python scripts/bench_attention.py
One JSON line per shape: CUDA-event time per launch (tcgen05 tiles + the mma.sync tail rows), issued TFLOP/s (3 plane products)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from semanticlens_b200 import ops  # noqa: E402

SHAPES = [("ViT-L-14", 64, 257, 16), ("SigLIP-L-16-256", 64, 256, 16), ("ViT-B-16", 128, 197, 12), ("ViT-L-14-336", 16, 577, 16)]
for name, B, T, H in SHAPES:
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda")
    planes = ops.split_planes(qkv, 0, 16.0)
    for _ in range(3):
        out = ops.attention_planes(planes, B, H, fmt=0)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = ops.attention_planes(planes, B, H, fmt=0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 4.0 * B * H * T * T * 64 * 3
    print(json.dumps({"shape": name, "B": B, "T": T, "H": H, "us": round(ms * 1e3, 1), "issued_TFLOP/s": round(flops / ms / 1e9, 1)}), flush=True)
