"""Summarise a section-limited ncu capture (SpeedOfLight, MemoryWorkloadAnalysis, Occupancy, WarpStateStats, LaunchStats,
SchedulerStats; exported with `ncu -i X.ncu-rep --page raw --csv`) into one markdown row per kernel (mean over its launches).

    python profiles/summarize_ncu_sections.py raw.csv "title" "command" > profiles/rN_ncu_loop.md
"""
import csv
import re
import sys
from collections import OrderedDict

COLS = [
    ("us", "gpu__time_duration.sum"),
    ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("L1/TEX %", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("waves / SM", "launch__waves_per_multiprocessor"),
]


def short(name):
    name = re.sub(r"^void ", "", name).replace("said::", "")
    return re.sub(r"\(.*$", "", name)[:80]


def main():
    raw, title, cmd = sys.argv[1:4]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in data:
        a = agg.setdefault(short(r[ix["Kernel Name"]]), [])
        vals = []
        for _, key in COLS:
            if key not in ix or r[ix[key]] in ("", "n/a"):
                vals.append(None)
                continue
            v = float(r[ix[key]].replace(",", ""))
            if key.startswith("gpu__time"):
                v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(units[ix[key]], v)
            vals.append(v)
        a.append(vals)
    print(f"# {title}\n\nCommand: `{cmd}`\n\n(mean over the captured launches of each kernel; under ncu every launch is serialised and replayed, so absolute times are "
          "upper bounds -- use the ratios)\n")
    print("| kernel | launches | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    for k, ls in agg.items():
        cells = []
        for j in range(len(COLS)):
            xs = [v[j] for v in ls if v[j] is not None]
            cells.append("-" if not xs else (f"{sum(xs) / len(xs):.0f}" if COLS[j][0] in ("regs", "grid", "block") else f"{sum(xs) / len(xs):.1f}"))
        print(f"| `{k}` | {len(ls)} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
