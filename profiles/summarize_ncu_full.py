"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv > raw.csv`) into a
markdown table (one row per captured launch) and the per-launch DRAM traffic json that bench.py reports as
`roofline.traffic`.

    ncu -i gpurun_out/full.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu_full.py /tmp/raw.csv "title" "command" profiles/rN_ncu_full.md profiles/rN_ncu_traffic.json
"""
import csv
import json
import re
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    ("DRAM read MB", "dram__bytes_read.sum", 1.0),
    ("DRAM write MB", "dram__bytes_write.sum", 1.0),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
    ("smem LSU wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("smem TC wavefronts %", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
]


def to_mb(v, unit):
    v = float(v)
    return {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(unit, v)


def short(name):
    name = re.sub(r"^void ", "", name).replace("said::tc::", "tc::").replace("said::", "")
    name = re.sub(r"\(int\)", "", name)
    return re.sub(r"\(.*$", "", name)[:90]


def main():
    raw, title, cmd, md_out, js_out = sys.argv[1:6]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}\n", f"Command: `{cmd}`\n", "| kernel | grid x block | " + " | ".join(c[0] for c in COLS) + " |",
             "|---|---|" + "---:|" * len(COLS)]
    agg = {}
    for r in data:
        name = short(r[ix["Kernel Name"]])
        vals = []
        for label, key, _ in COLS:
            if key not in ix:
                vals.append("-")
                continue
            v, u = r[ix[key]], units[ix[key]]
            if "MB" in label:
                vals.append(f"{to_mb(v, u):.1f}")
            elif label == "us":
                f = float(v)
                f = {"ns": f / 1e3, "us": f, "ms": f * 1e3, "usecond": f, "nsecond": f / 1e3, "msecond": f * 1e3}.get(u, f)
                vals.append(f"{f:.1f}")
            else:
                vals.append(f"{float(v):.1f}" if v not in ("", "n/a") else "-")
        grid = r[ix["Grid Size"]] if "Grid Size" in ix else "?"
        block = r[ix["Block Size"]] if "Block Size" in ix else "?"
        lines.append(f"| `{name}` | {grid} x {block} | " + " | ".join(vals) + " |")
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "bytes": 0.0})
        a["n"] += 1
        a["us"] += float(vals[0])
        a["bytes"] += (float(vals[3]) + float(vals[4])) * 1e6
    lines.append("\nPer-launch averages (traffic = dram__bytes_read.sum + dram__bytes_write.sum):\n")
    for k, a in agg.items():
        lines.append(f"* `{k}`: {a['n']} launches, {a['us'] / a['n']:.1f} us, {a['bytes'] / a['n'] / 1e6:.1f} MB DRAM traffic")
    open(md_out, "w").write("\n".join(lines) + "\n")
    gemm = [a for k, a in agg.items() if "gemm_tc_kernel" in k]
    if gemm:
        n = sum(a["n"] for a in gemm)
        js = {"source": f"{md_out} (ncu --set full, {n} gemm_tc_kernel launches of one loop iteration)",
              "kernel": "gemm_tc_kernel (all instantiations)",
              "dram_bytes_per_launch": sum(a["bytes"] for a in gemm) / n,
              "note": "launch-weighted mean of dram__bytes_read.sum + dram__bytes_write.sum over the captured gemm_tc_kernel launches"}
        json.dump(js, open(js_out, "w"), indent=1)


if __name__ == "__main__":
    main()
