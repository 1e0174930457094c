"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches1.csv > profiles/rNN_launches_<what>.md
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep [regex] > profiles/rNN_<kernel>_full.md
"""

from __future__ import annotations

import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg",
]


def to_ns(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    return x * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)


def launches(path: str):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\s+", " ", row["Kernel Name"])[:96]
        d = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]])
        d[0] += 1
        d[1] += to_ns(row["Metric Value"], row["Metric Unit"])
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: `{path}`\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES).\n")
    print(f"{n} launches, {tot / 1e6:.3f} ms total device time\n")
    print("| ms | share | launches | avg us | grid | block | kernel |\n|---:|---:|---:|---:|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f}% | {v[0]} | {v[1] / v[0] / 1e3:.1f} | {v[2]} | {v[3]} | `{k}` |")
    # every kernel of libslb200 lives in an anonymous namespace; torch's own anonymous kernels carry an `at::` prefix
    ours = {k: v for k, v in agg.items() if "<unnamed>::" in k and "at::" not in k and "cudnn" not in k}
    if ours:
        t = sum(v[1] for v in ours.values())
        print(f"\nlibslb200 kernels: {t / 1e6:.3f} ms = {100 * t / tot:.2f}% of the listed device time")


def full(path: str, pat: str | None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in FULL_METRICS if m in hdr]
    print(f"# ncu --set full: `{path}`\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not re.search(pat, name):
            continue
        print(f"## `{re.sub(r'\\s+', ' ', name)[:110]}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for c in cols:
            print(f"| {hdr[c]} | {r[c]} | {units[c]} |")
        try:
            rd = to_ns(r[hdr.index("dram__bytes_read.sum")], "") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
            wr = to_ns(r[hdr.index("dram__bytes_write.sum")], "") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
            ns = to_ns(r[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
            print(f"| traffic (dram read+write) | {(rd + wr) / 1e6:.3f} | MB |")
            print(f"| traffic / duration (under profiler) | {(rd + wr) / ns:.1f} | GB/s |")
        except (ValueError, KeyError):
            pass
        print()


PROFILE_NAMES = [  # kernel function name pattern -> the name the in-library profiler (slb_profile_*) reports
    (r"gemm_split", "K4 gemm_split (tcgen05)"), (r"attention_", "K4 attention"), (r"layernorm_kernel", "K4 layernorm"),
    (r"patchify", "K4 patchify"), (r"assemble_kernel", "K4 assemble_tokens"), (r"u8_norm", "K3 u8_to_f32_norm"),
    (r"agg_rows|agg_btf|token_select", "K1 agg_reduce"), (r"topk_update", "K2 topk_update"),
    (r"topk_merge_lists", "K2 topk_merge_lists"), (r"gather_rows", "K5 gather_rows"), (r"clarity_kernel", "K7 clarity"),
    (r"polysem_kernel", "K8 polysem_2means"), (r"normalize_split", "K6 normalize_split_rows"), (r"split_planes", "split_planes"),
]


def traffic(path: str):
    """ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum CSV -> JSON on stdout:
    {profile kernel name: {launches, dram_bytes_per_launch, ...}} (what bench.py reports as roofline.traffic)."""
    import json

    lines = [l for l in open(path) if not l.startswith("==")]
    per_launch: dict = {}
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        if "<unnamed>::" not in name or "at::" in name:
            continue
        key = next((pn for pat, pn in PROFILE_NAMES if re.search(pat, name)), None)
        if key is None:
            continue
        d = per_launch.setdefault((key, row["ID"]), {})
        unit = row["Metric Unit"]
        val = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"].startswith("dram__bytes"):
            val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["dram"] = d.get("dram", 0.0) + val
        elif row["Metric Name"] == "gpu__time_duration.sum":
            d["ns"] = to_ns(row["Metric Value"], unit)
    out: dict = {}
    for (key, _), d in per_launch.items():
        o = out.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "ns": 0.0})
        o["launches"] += 1
        o["dram_bytes"] += d.get("dram", 0.0)
        o["ns"] += d.get("ns", 0.0)
    res = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
               "us_per_launch_under_ncu": v["ns"] / v["launches"] / 1e3,
               "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum ({path})"} for k, v in out.items()}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
