"""bench.py's output contract on a machine without a GPU: the reference arm prints exactly ONE JSON line on stdout
(library chatter such as NCCL's version banner goes to stderr) with the keys the driver reads."""

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("concept-db images/sec")


def test_stdout_is_reserved_for_the_result_line():
    """Anything written to fd 1 after the claim (a C library's printf, a stray print) lands on stderr."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._claim_stdout(); os.write(1, b'banner\\n'); "
            "print('chatter'); bench.emit({'ok': 1})" % str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == '{"ok": 1}'
    assert "banner" in r.stderr and "chatter" in r.stderr


def test_ring_dataset_of_the_full_run_maps_global_ids_onto_the_host_ring():
    """bench.RingImages (the dataset of the cfg2_full_run sub-record): item i of a rank's shard is image (i - lo) % ring of a
    small host ring, batch-aligned ranges never wrap, ranges outside the shard are refused, both views share the pixels."""
    import pytest
    import torch

    sys.path.insert(0, str(ROOT))
    import bench

    B, ring_batches = 4, 3
    lo, hi, n_total = 40, 40 + 7 * B, 200
    dm = bench.RingImages(n_total, lo, hi, 5, "cpu", "model", B, ring_batches)
    df = bench.RingImages(n_total, lo, hi, 5, "cpu", "fm", B, ring_batches, store=dm.u8)
    assert len(dm) == len(df) == n_total and dm.ring == B * ring_batches and dm.u8.shape == (dm.ring, 3, 224, 224)
    assert dm.name != df.name and str(n_total) in dm.name
    for step in range(7):
        a = lo + step * B
        x, y = dm.get_batch(a, a + B)
        u = df.get_batch(a, a + B)
        j = (step * B) % dm.ring
        assert torch.equal(u, dm.u8[j : j + B]) and x.shape == (B, 3, 224, 224) and y.shape == (B,)
        assert torch.equal(x, bench.normalise(u))
        assert torch.equal(df[a + 1], dm.u8[(j + 1) % dm.ring]) and torch.equal(dm[a + 1][0], x[1])
    with pytest.raises(AssertionError):
        dm.get_batch(lo - B, lo)  # outside this rank's shard
    with pytest.raises(AssertionError):
        dm.get_batch(lo + 2, lo + 2 + dm.ring)  # would wrap the ring
