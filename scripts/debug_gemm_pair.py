"""Debug harness for the CTA-pair GEMM: tiny problems first, each in a fresh subprocess with a wall-clock limit."""
import os
import subprocess
import sys

CASE = r'''
import sys, torch
sys.path.insert(0, ".")
from semanticlens_b200 import ops
M, N, K = map(int, sys.argv[1:4])
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.05
out, _ = ops.gemm_split(ops.split_planes(a), ops.split_planes(w))
torch.cuda.synchronize()
want = a.double() @ w.double().T
err = ((out.double() - want).abs().max() / want.abs().max()).item()
print("M N K", M, N, K, "rel err", err, flush=True)
'''
for shape in ((256, 128, 64), (256, 128, 256), (512, 256, 768), (12800, 2304, 768)):
    env = dict(os.environ, SLB_GEMM_DEBUG="1")
    try:
        r = subprocess.run([sys.executable, "-c", CASE, *map(str, shape)], env=env, capture_output=True, text=True, timeout=60)
        print(shape, "rc", r.returncode, r.stdout.strip()[-300:], r.stderr.strip()[-1500:], flush=True)
        if r.returncode != 0:
            break
    except subprocess.TimeoutExpired as e:
        print(shape, "TIMEOUT", (e.stdout or b"")[-300:], (e.stderr or b"")[-1500:], flush=True)
        break
