"""One launch of the polysemanticity kernel at cfg-5 neuron shape for ncu: python scripts/ncu_polysem.py"""
import sys, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
from semanticlens_b200 import ops
C = 296 * 8
g = torch.Generator(device="cuda").manual_seed(2)
V = torch.randn(C, 256, 512, device="cuda", generator=g)
V[:, ::2] += 2 * torch.randn(C, 1, 512, device="cuda", generator=g)
ops.polysem_2means(V); torch.cuda.synchronize()
ops.polysem_2means(V); torch.cuda.synchronize()
