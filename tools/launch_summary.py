"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table (kernel, launches, mean us,
share of the summed kernel time).  Usage: python tools/launch_summary.py launches.csv > profiles/rN_launches_summary.md"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    if name.startswith("at::"):
        return "torch elementwise / reduce kernels (host-side logging and the bench's own bookkeeping)"
    m = re.match(r"([A-Za-z0-9_:]+(<[^()]*>)?)\(", name)
    name = m.group(1) if m else name
    return "`" + name + "`"


def main(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
            rows.append((short(r["Kernel Name"]), v))
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | mean us | share |")
    print("|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {t / n:.1f} | {100 * t / total:.1f} % |")
    print(f"\n{len(rows)} launches, {total / 1e3:.2f} ms of kernel time in the list.")


if __name__ == "__main__":
    main(sys.argv[1])
