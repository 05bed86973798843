#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
    python tools/summarize_ncu.py full gpurun_out/prof_fused.ncu-rep > profiles/rNN_fused_full.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = r["Kernel Name"][:90]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {len(rows)} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised: compare SHARES)\n")
    print("| us total | launches | share | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1]:.1f} | {v[0]} | {100 * v[1] / tot:.1f}% | `{k}` |")
    print("\n## fused-layer launches in order (last pass)\n")
    fams = ("bt_fused", "bt_ws", "bt_direct", "bt_tma_kernel", "bt_tms_kernel", "bt_dtma_kernel")
    fused = [r for r in rows if any(t in r["Kernel Name"] for t in fams)]
    n = 21 if len(fused) >= 21 else len(fused)
    print("| # | kernel | grid | block | duration |\n|---:|---|---|---|---:|")
    import re
    for i, r in enumerate(fused[-n:]):
        nm = re.sub(r"void |<unnamed>::|\(.*", "", r["Kernel Name"])
        print(f"| {i} | `{nm}` | {r['Grid Size']} | {r['Block Size']} | {r['Metric Value']} {r['Metric Unit']} |")


def full(path):
    if path.endswith(".csv"):          # `ncu -i x.ncu-rep --page raw --csv` already run where the report was made
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full: {len(data)} captured launches of {path}\n")
    cols = [k for k in KEYS if k in ix]
    print("| metric (unit) | " + " | ".join(str(i) for i in range(len(data))) + " |")
    print("|---|" + "---:|" * len(data))
    for k in cols:
        vals = []
        for d in data:
            v = d[ix[k]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {k} ({units[ix[k]]}) | " + " | ".join(vals) + " |")
    names = [d[ix["Kernel Name"]][:48] for d in data] if "Kernel Name" in ix else []
    if names:
        print("\n| # | kernel |\n|---:|---|")
        for i, nme in enumerate(names):
            print(f"| {i} | `{nme}` |")
    # DRAM traffic of the captured launches (bench.py reads the json next to the .md: roofline.traffic)
    if "dram__bytes_read.sum" in ix and len(sys.argv) > 3:
        import json

        def to_bytes(v, u):
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        rd = sum(to_bytes(d[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) for d in data)
        wr = sum(to_bytes(d[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]) for d in data)
        ent = {"source": path, "launches": len(data), "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_step": rd + wr,
               "note": "sum over the Bayesian-layer launches of ONE bench step (ncu --set full, cold caches, serialised)"}
        key = sys.argv[4] if len(sys.argv) > 4 else None          # "fp32" | "bf16": one json carries both models
        if key:
            import os
            cur = json.load(open(sys.argv[3])) if os.path.exists(sys.argv[3]) else {}
            if "launches" in cur:
                cur = {}
            cur[key] = ent
            ent = cur
        json.dump(ent, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
