"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) of profiles/run_profile.py into a markdown
summary: loop kernels (after prepare_latents_kernel) and once-per-batch kernels (before it), grouped by kernel.

    python profiles/summarize_launches.py gpurun_out/launches.csv "title" "command" > profiles/rN_launches_x.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = name.replace("said::tc::", "tc::")
    name = re.sub(r"\(.*$", "", name)
    return name[:120]


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
    by_id = OrderedDict()
    for r in csv.DictReader(lines):
        m = r.get("Metric Name")
        if m == "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed":
            e = by_id.setdefault(r["ID"], [short(r["Kernel Name"]), 0.0, 0.0, False])
            if len(e) < 5:
                e.append(0.0)
            e[4] = float(r["Metric Value"].replace(",", ""))
            continue
        if m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            e = by_id.setdefault(r["ID"], [short(r["Kernel Name"]), 0.0, 0.0, False])
            v = float(r["Metric Value"].replace(",", "")) * unit.get(r.get("Metric Unit", "ns"), 1e-6 if m.startswith("gpu") else 1.0)
            if m.startswith("gpu"):
                e[1] = v
            else:
                e[2] += v
                e[3] = True
    rows = [(e[0], e[1], e[2], e[4] if len(e) > 4 else None) for e in by_id.values()]
    have_dram = any(e[3] for e in by_id.values())
    have_tensor = any(len(e) > 4 for e in by_id.values())
    split = max((i for i, r_ in enumerate(rows) if "prepare_latents_kernel" in r_[0]), default=-1)
    pre, loop = rows[: split + 1], rows[split + 1:]
    print(f"# {title}\n\nCommand: `{cmd}`\n(cold-cache, serialised launches: compare shares, not absolutes.)\n")
    for head, part in (("Loop iterations", loop), ("Once per batch: audio encoder + K/V hoist + tables", pre)):
        agg = OrderedDict()
        for k, ms, by, tp in part:
            a = agg.setdefault(k, [0.0, 0, 0.0, 0.0])
            a[0] += ms
            a[1] += 1
            a[2] += by
            a[3] += (tp or 0.0) * ms           # time-weighted tensor-pipe utilisation
        tot = sum(v[0] for v in agg.values())
        totb = sum(v[2] for v in agg.values())
        extra = f", {totb / 1e6:.0f} MB of DRAM traffic (read + write)" if have_dram else ""
        print(f"## {head}\n\n{len(part)} launches, {tot:.2f} ms{extra}.\n")
        hdr = "| ms | share | launches | us / launch |" + (" DRAM MB / launch |" if have_dram else "") + (" tensor pipe % |" if have_tensor else "") + " kernel |"
        print(hdr + "\n|" + "---:|" * (hdr.count("|") - 2) + "---|")
        for k, (ms, n, by, tpw) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            mid = (f" {by / n / 1e6:.1f} |" if have_dram else "") + (f" {tpw / ms if ms > 0 else 0.0:.1f} |" if have_tensor else "")
            print(f"| {ms:.3f} | {100 * ms / tot:.1f}% | {n} | {1000 * ms / n:.1f} |{mid} `{k}` |")
        print()


if __name__ == "__main__":
    main()
