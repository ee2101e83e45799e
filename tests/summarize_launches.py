"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (developer tool).
    python tests/summarize_launches.py profiles/r01_v1_launches.csv
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for row in rows:
        name = row["Kernel Name"]
        if not name.startswith(("sb::", "void sb::")):
            continue                                  # torch kernels of the synthetic-input generator
        name = name[:78]
        v = float(row["Metric Value"].replace(",", ""))
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{len(rows)} launches captured; library kernels only; total {tot / 1e6:.3f} ms (cold-cache, serialised)")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / tot * 100:6.2f}%  n={c:4d}  total={v / 1e6:9.3f} ms  avg={v / c / 1e3:9.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
