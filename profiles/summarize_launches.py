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
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e6))
    split = max((i for i, (k, _) in enumerate(rows) if "prepare_latents_kernel" in k), default=-1)
    pre, loop = rows[: split + 1], rows[split + 1:]
    print(f"# {title}\n\nCommand: `{cmd}`\n(cold-cache, serialised launches: compare shares, not absolutes.)\n")
    for head, part in (("Loop iterations", loop), ("Once per batch: audio encoder + K/V hoist + tables", pre)):
        agg = OrderedDict()
        for k, ms in part:
            a = agg.setdefault(k, [0.0, 0])
            a[0] += ms
            a[1] += 1
        tot = sum(v[0] for v in agg.values())
        print(f"## {head}\n\n{len(part)} launches, {tot:.2f} ms.\n\n| ms | share | launches | us / launch | kernel |\n|---:|---:|---:|---:|---|")
        for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"| {ms:.3f} | {100 * ms / tot:.1f}% | {n} | {1000 * ms / n:.1f} | `{k}` |")
        print()


if __name__ == "__main__":
    main()
