"""One launch of the long-sequence attention kernel for ncu: python scripts/ncu_attention.py [T] [B] [H]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from semanticlens_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = int(sys.argv[3]) if len(sys.argv) > 3 else 16
qkv = torch.randn(B * T, 3 * H * 64, device="cuda")
planes = ops.split_planes(qkv, 0, 16.0)
for _ in range(3):
    ops.attention_planes(planes, B, H, fmt=0)
torch.cuda.synchronize()
