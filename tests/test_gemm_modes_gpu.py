"""The GEMM's two epilogues (per-lane stores after a shared-memory transpose / TMA stores of swizzled boxes) are chosen
per launch by the host; SLB_GEMM_TMA_STORE forces one for every eligible launch. The env var is read once per process, so
each mode runs the GEMM and tower parity tests in a child interpreter."""

import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["0", "2"])
def test_gemm_and_tower_parity_with_the_store_path_forced(mode):
    env = dict(os.environ, SLB_GEMM_TMA_STORE=mode)
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "tests/test_gemm_gpu.py",
         "tests/test_rn_gpu.py", "tests/test_embed_gpu.py::test_vit_tower_vs_oracle", "tests/test_embed_gpu.py::test_siglip_tower_vs_oracle"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_short_k_gemms_with_eight_epilogue_warps():
    """K <= 512 one-CTA GEMMs default to the sixteen-warp epilogue (16-column drains); SLB_GEMM_EPI_WARPS=8 keeps the
    eight-warp kernel (32-column drains). Both must pass the same parity suites."""
    env = dict(os.environ, SLB_GEMM_EPI_WARPS="8")
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "tests/test_gemm_gpu.py", "tests/test_rn_gpu.py",
         "tests/test_probed_gpu.py::test_implicit_gemm_convolution", "tests/test_probed_gpu.py::test_raw_output_rides_along_with_the_fused_epilogue"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_short_sequences_on_the_mma_sync_attention():
    """SLB_ATTN_PACK=0 keeps T < 128 attention (ViT-B/32, the text towers) on the mma.sync kernel instead of packing the images
    into slots of a tcgen05 tile: same numerics contract, whole towers and the causal text tower included."""
    env = dict(os.environ, SLB_ATTN_PACK="0")
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
         "tests/test_embed_gpu.py::test_attention_from_planes_vs_torch", "tests/test_embed_gpu.py::test_vit_tower_vs_oracle",
         "tests/test_embed_gpu.py::test_vit_tower_batch_composition_invariance", "tests/test_embed_gpu.py", "-k", "attention or tower or text"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_key_tail_as_a_block_of_its_own():
    """SLB_ATTN_KEY_TAIL=0: T = 128 n + 1 runs the last key as one more key block (the arrangement before the SIMT tail)."""
    env = dict(os.environ, SLB_ATTN_KEY_TAIL="0")
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
         "tests/test_embed_gpu.py::test_attention_from_planes_vs_torch"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
