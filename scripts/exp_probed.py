"""Host-framework experiment: how fast can torch run the probed ResNet-50 forward (user code, not libslb200) in strict
fp32 on this GPU under different cuDNN settings? Prints one JSON line per setting."""
import json
import sys
from pathlib import Path

import torch
import torchvision

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def timeit(fn, warm=3, it=8):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def main():
    B = 256
    torch.manual_seed(0)
    m = torchvision.models.resnet50(weights=None).eval().cuda()
    x = torch.randn(B, 3, 224, 224, device="cuda")
    ref = None
    for name, bench, cl, tf32 in (("strict fp32 (default algos)", False, False, False), ("strict fp32 + cudnn.benchmark", True, False, False),
                                  ("strict fp32 + channels_last", False, True, False), ("strict fp32 + channels_last + benchmark", True, True, False),
                                  ("tf32 convs (torch default) + benchmark", True, False, True)):
        torch.backends.cudnn.benchmark = bench
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        mm = m.to(memory_format=torch.channels_last) if cl else m.to(memory_format=torch.contiguous_format)
        xx = x.contiguous(memory_format=torch.channels_last) if cl else x
        with torch.no_grad():
            ms = timeit(lambda: mm(xx))
            out = mm(xx).float()
        if ref is None:
            ref = out
        err = ((out - ref).abs().max() / ref.abs().max()).item()
        print(json.dumps({"setting": name, "ms_per_256": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1),
                          "max_rel_diff_vs_first": err}), flush=True)


if __name__ == "__main__":
    main()
