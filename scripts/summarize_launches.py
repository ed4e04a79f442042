"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (sum / count / share)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        short = r["Kernel Name"].split("(")[0]
        short = re.sub(r"^void ", "", short)
        short = short.replace("pcy::<unnamed>::", "").replace("pcy::", "")
        if len(short) > 90:
            short = short[:87] + "..."
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    out_lines = [f"# {path}: {sum(v[0] for v in agg.values())} launches, {total/1e3:.3f} ms total (ncu, serialised, cold cache)",
                 f"{'kernel':92s} {'launches':>8s} {'sum_us':>12s} {'avg_us':>10s} {'share':>7s}"]
    for k in sorted(agg, key=lambda k: -agg[k][1]):
        n, s = agg[k]
        out_lines.append(f"{k:92s} {n:8d} {s:12.1f} {s/n:10.2f} {100*s/total:6.2f}%")
    txt = "\n".join(out_lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
