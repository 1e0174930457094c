"""Per-kernel live profile (slb_profile_*) of one foundation-model tower forward. python scripts/profile_tower.py ViT-L-14 64"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from semanticlens_b200 import _native  # noqa: E402
from semanticlens_b200.foundation_models import OpenClip, rn, vit  # noqa: E402

url, B = sys.argv[1], int(sys.argv[2])
fm = OpenClip(url, device="cuda", load_weights=False, seed=1)
S = fm.cfg.image_size
u8 = torch.randint(0, 255, (B, 3, S, S), dtype=torch.uint8, device="cuda")
for _ in range(3):
    fm.encode_image(fm.preprocess(u8))
torch.cuda.synchronize()
n = 5
_native.profile_begin()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    fm.encode_image(fm.preprocess(u8))
b.record()
torch.cuda.synchronize()
kt = _native.profile_end()
ms = a.elapsed_time(b) / n
out = {"tower": url, "batch": B, "ms": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1),
       "issued_TFLOPs_overall": round(3 * (rn if fm.url in rn.CONFIGS else vit).flops_per_image(fm.cfg) * B / ms / 1e9, 1), "kernels": {}}
for k, d in sorted(kt.items(), key=lambda kv: -kv[1]["ms"]):
    e = {"ms": round(d["ms"] / n, 3), "share": round(d["ms"] / n / ms, 3), "launches": d["launches"] // n}
    if d["flops"]:
        e["TFLOP/s"] = round(d["flops"] / d["ms"] / 1e9, 1)
    if d["bytes"]:
        e["GB/s"] = round(d["bytes"] / d["ms"] / 1e6, 1)
    out["kernels"][k] = e
print(json.dumps(out))
