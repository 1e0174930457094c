"""BASELINE.json configs[4]: text_probing + eval_polysemanticity (+ clarity) over a 65 536-neuron x 256-example x 512-d
concept DB against 10 000 text-prompt embeddings on one B200, with the reference's CPU path timed beside it on a
bounded sample. One JSON line per score. (A parity-test configuration, not the bench.py headline.)

    python scripts/bench_cfg5.py [--neurons 65536] [--cpu-neurons 64]
"""
import argparse
import json
import os
import sys
import time
import warnings
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def cuda_time(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--neurons", type=int, default=65536)
    ap.add_argument("--cpu-neurons", type=int, default=64)
    a = ap.parse_args()
    from oracle import ref_port as rp
    from semanticlens_b200 import scores as S

    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) \
        if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    C, k, D, Q = a.neurons, 256, 512, 10000
    g = torch.Generator(device="cuda").manual_seed(2)
    V = torch.randn(C, k, D, device="cuda", generator=g)
    V[:, ::2] += 2 * torch.randn(C, 1, D, device="cuda", generator=g)  # planted 2-cluster structure
    text = torch.randn(Q, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    agg = V.mean(1)
    torch.set_num_threads(os.cpu_count() or 1)
    n_cpu = a.cpu_neurons
    Vc, tc, aggc = V[:n_cpu].cpu(), text.cpu(), agg.cpu()

    ms, sim = cuda_time(lambda: S.similarity_score(text, agg))
    t0 = time.perf_counter(); ref = rp.similarity_score(tc, aggc); cpu_s = time.perf_counter() - t0
    err = ((sim.cpu() - ref).abs().max() / ref.abs().max()).item()
    fl = 2.0 * Q * C * D
    print(json.dumps({"score": "similarity (text_probing)", "shape": [Q, C, D], "gpu_ms": round(ms, 3),
                      "fp32_equiv_TFLOPs": round(fl / ms / 1e9, 1), "issued_TFLOPs": round(3 * fl / ms / 1e9, 1),
                      "frac_of_measured_bf16_burst": round(3 * fl / ms / 1e9 / peaks["bf16_tflops"], 3), "cpu_s": round(cpu_s, 3),
                      "cpu_cores": torch.get_num_threads(), "speedup": round(cpu_s * 1e3 / ms, 1), "max_rel_err_vs_reference_port": err}), flush=True)

    ms, cl = cuda_time(lambda: S.clarity_score(V))
    t0 = time.perf_counter(); ref = rp.clarity_score(Vc); cpu_s = time.perf_counter() - t0
    err = (cl[:n_cpu].cpu() - ref).abs().max().item()
    by = V.numel() * 4
    print(json.dumps({"score": "clarity", "shape": [C, k, D], "gpu_ms": round(ms, 3), "GBps": round(by / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(by / ms / 1e6 / peaks["hbm_gbs"], 3),
                      "cpu_s_extrapolated": round(cpu_s * C / n_cpu, 1), "cpu_sample_neurons": n_cpu,
                      "speedup": round(cpu_s * C / n_cpu * 1e3 / ms, 1), "max_abs_err_vs_reference_port": err}), flush=True)

    ms, po = cuda_time(lambda: S.polysemanticity_score(V), iters=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter(); ref = rp.polysemanticity_score(Vc); cpu_s = time.perf_counter() - t0
    diff = (po[:n_cpu].cpu() - ref).abs()
    print(json.dumps({"score": "polysemanticity", "shape": [C, k, D], "gpu_ms": round(ms, 1), "us_per_neuron": round(ms * 1e3 / C, 2),
                      "GBps": round(by / ms / 1e6, 1), "cpu_s_extrapolated": round(cpu_s * C / n_cpu, 1), "cpu_sample_neurons": n_cpu,
                      "speedup": round(cpu_s * C / n_cpu * 1e3 / ms, 1), "max_abs_err_vs_reference_port": diff.max().item(),
                      "neurons_in_a_different_optimum": int((diff > 1e-6).sum())}), flush=True)


if __name__ == "__main__":
    main()
