"""The end-to-end collect contract of SURVEY.md §8(c) for a CPU sweep (the reference's op sequence on oneDNN) against a
GPU sweep (cuDNN) of the SAME probed model on the SAME images.

The two backends' fp32 activations differ at the 1e-6 level, which flips the bf16 rounding of an aggregate only when
its exact value sits next to a rounding midpoint. Instead of a percentage threshold the test ENUMERATES those elements
from a float64 forward of the model and an a-priori error budget, and demands exactness everywhere else:

  budget   tau[c] = SIGMAS * max(sigma_cpu[c], sigma_gpu[c], REL_FLOOR * scale[c]),  sigma_x[c] = rms_i(a_x32[i, c] - a64[i, c]),
           scale[c] = mean |summand of channel c| — the fp32 forward noise of channel c on either backend against the
           float64 run (a per-CHANNEL statistic over all images, 8-sigma margin; cuDNN's Winograd / implicit-GEMM
           kernels are noisier than oneDNN's direct convolution, and the noise grows with the depth of the hooked layer).
           sigma_gpu itself must stay fp32-grade (median over channels of sigma_gpu / scale <= GPU_NOISE_CEILING), so a
           systematic GPU error cannot hide behind its own budget
  excused  element (image i, channel c) iff bf16(a64 - tau) != bf16(a64 + tau)        (a64 = float64 aggregate)
  (1) every (i, c) whose bf16 candidate differs between the two backends is excused — anything else is a bug;
  (2) the GPU state is EXACTLY the canonical top-k of the GPU candidates (values and ids, bit for bit), for every row;
  (3) the CPU port's state meets the tie-aware contract against the canonical top-k of the CPU candidates;
  (4) rows without a differing candidate: GPU state vs CPU-port state under oracle.collect.check_tie_aware, no slack.
"""

import copy

import numpy as np
import torch

from oracle import collect as oc

SIGMAS = 8.0
REL_FLOOR = 1e-7
GPU_NOISE_CEILING = 2e-5


def _tap(model, layers, batches, to, forward=None):
    seen = {n: [] for n in layers}
    mods = dict(model.named_modules())
    taps = [mods[n].register_forward_hook(lambda m, i, o, n=n: seen[n].append(to(o.detach()))) for n in layers]
    with torch.no_grad():
        for x in batches:
            (forward or model)(x)
    for t in taps:
        t.remove()
    return seen


def check_collect_contract(net_cpu, layers, batches_cpu, op, kind, k, gpu_state, ref_state, device="cuda", gpu_forward=None,
                           noise_ceiling=GPU_NOISE_CEILING):
    """net_cpu: the probed model on the CPU (fp32). batches_cpu: list of fp32 input batches in sweep order.
    gpu_state / ref_state: {layer: (bits uint16 (C,k), ids int64 (C,k))}. Returns a per-layer report.
    gpu_forward: factory(model on the device) -> callable that runs the forward and fires the model's hooks (the opt-in
    accelerated forward, probed.AcceleratedResNet); default: the torch model itself. noise_ceiling: the fp32-grade bar of
    the GPU forward (22-bit-operand convolutions sit a little above cuDNN's fp32)."""
    from semanticlens_b200 import _native as N
    from semanticlens_b200 import ops

    opcode = {"mean": N.AGG_MEAN, "max": N.AGG_MAX, "absmean": N.AGG_ABSMEAN, "absmax": N.AGG_ABSMAX}[op]
    net64 = copy.deepcopy(net_cpu).double()
    maps64 = _tap(net64, layers, [b.double() for b in batches_cpu], lambda o: o.numpy())
    maps32 = _tap(net_cpu, layers, batches_cpu, lambda o: o.numpy())
    net_gpu = copy.deepcopy(net_cpu).to(device)
    agg_gpu = _tap(net_gpu, layers, [b.to(device) for b in batches_cpu], lambda o: ops.agg_reduce(o, opcode, kind).cpu().numpy(),
                   forward=gpu_forward(net_gpu) if gpu_forward else None)
    report = {}
    for name in layers:
        m64 = np.concatenate(maps64[name])
        flat = m64.reshape(m64.shape[0], m64.shape[1], -1) if kind == "conv" else m64.transpose(0, 2, 1)
        if op in ("mean", "absmean"):
            a64 = (np.abs(flat) if op == "absmean" else flat).mean(-1)
        else:
            a64 = (np.abs(flat) if op == "absmax" else flat).max(-1)
        agg_cpu = np.concatenate([oc.aggregate_exact(m, op, kind) for m in maps32[name]])  # (N, C) fp32
        agg_dev = np.concatenate(agg_gpu[name])
        scale = np.abs(flat).mean(axis=(0, 2))[None, :]
        sigma_cpu = np.sqrt(((agg_cpu.astype(np.float64) - a64) ** 2).mean(axis=0))[None, :]
        sigma_gpu = np.sqrt(((agg_dev.astype(np.float64) - a64) ** 2).mean(axis=0))[None, :]
        # layer-level sanity (the median over channels: a nearly dead post-ReLU channel has a tiny scale but inherits the
        # noise of its pre-activation, so single channels can sit far above the typical ratio)
        ratio = np.median(sigma_gpu / (scale + 1e-30))
        assert ratio <= noise_ceiling, f"{name}: the GPU forward is not fp32-grade: median noise / scale = {ratio:.2e}"
        tau = SIGMAS * np.maximum(np.maximum(sigma_cpu, sigma_gpu), REL_FLOOR * scale)  # (1, C)
        excused = oc.f32_to_bf16_bits((a64 - tau).astype(np.float32)) != oc.f32_to_bf16_bits((a64 + tau).astype(np.float32))
        cand_cpu = oc.f32_to_bf16_bits(agg_cpu)
        cand_gpu = oc.f32_to_bf16_bits(agg_dev)
        differ = ~oc.values_equal(cand_cpu, cand_gpu)
        # (1)
        unexplained = differ & ~excused
        assert not unexplained.any(), (
            f"{name}: {int(unexplained.sum())} candidate(s) differ between the backends away from any bf16 rounding "
            f"midpoint, first at (image, channel) {np.argwhere(unexplained)[0].tolist()}")
        n = cand_gpu.shape[0]
        ids = np.arange(n)
        # (2)
        canon_gpu = oc.ActMaxOracle(k)
        canon_gpu.update(oc.bf16_bits_to_f32(cand_gpu), ids)
        gb, gi = gpu_state[name]
        assert oc.values_equal(gb, canon_gpu.bits).all(), f"{name}: GPU top-k values are not the top-k of the GPU candidates"
        assert (gi == canon_gpu.ids).all(), f"{name}: GPU top-k ids are not in the canonical order of the GPU candidates"
        # (3)
        canon_cpu = oc.ActMaxOracle(k)
        canon_cpu.update(oc.bf16_bits_to_f32(cand_cpu), ids)
        rb, ri = ref_state[name]
        errs = oc.check_tie_aware(canon_cpu.bits, canon_cpu.ids, rb, ri, cand_cpu.T)
        assert not errs, f"{name}: reference port vs its own candidates: {errs[:3]}"
        # (4)
        clean = ~differ.any(axis=0)  # channels whose candidates are identical on both backends
        errs = oc.check_tie_aware(gb[clean], gi[clean], rb[clean], ri[clean], cand_gpu.T[clean])
        assert not errs, f"{name}: GPU vs reference port on rows with identical candidates: {errs[:3]}"
        report[name] = {"gpu_noise_over_scale_median": float(ratio),
                        "cpu_noise_over_scale_median": float(np.median(sigma_cpu / (scale + 1e-30))), "elements": int(differ.size), "excused": int(excused.sum()), "differing": int(differ.sum()),
                        "rows": int(clean.size), "rows_checked_exactly": int(clean.sum())}
    return report
