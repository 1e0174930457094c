"""Timeline of one work item of one CTA of the tcgen05 attention kernel (attention_ts.cu, SLB_ATTN_TRACE=1): SM-clock
offsets of every hand-off since the CTA started.
SLB_ATTN_TRACE=1 python scripts/trace_attention.py [T] [B] [H]"""
import ctypes
import os
import sys
from pathlib import Path

os.environ["SLB_ATTN_TRACE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from semanticlens_b200 import _native as N  # noqa: E402
from semanticlens_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 257
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = int(sys.argv[3]) if len(sys.argv) > 3 else 16
qkv = torch.randn(B * T, 3 * H * 64, device="cuda")
planes = ops.split_planes(qkv, 0, 16.0)
for _ in range(3):
    ops.attention_planes(planes, B, H, fmt=0)
torch.cuda.synchronize()
ptr = N.load().slb_attention_trace()
w = (ctypes.c_uint32 * 256).from_address(ptr)
names = ["tma issue (ring item)", "mma: K landed (it)", "mma: S issue (it)", "mma: PV issue (blk)", "sm: S visible (it)",
         "sm: maxima done (it)", "sm: exps+stores (blk)", "sm: P published (blk)", "mma: V landed (blk)", "O complete / item done"]
print("SM id of the traced CTA:", w[63], "(SLB_ATTN_TRACE_CTA / SLB_ATTN_TRACE_ITEM pick the CTA and its n-th work item)")
nblk = (T + 63) // 64
for t, nm in enumerate(names):
    vals = [w[64 + 16 * t + i] for i in range(16)]
    print(f"{nm:26s}", " ".join(f"{v:7d}" for v in vals if v))
