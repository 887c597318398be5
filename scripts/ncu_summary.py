"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the
ordered list of tensor-core launches.  Usage: python scripts/ncu_summary.py launches.csv [--seq]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"void |plume::|at::native::", "", name)
        rows.append((name, v, row.get("Grid Size", "")))
    return rows


def main():
    rows = load(sys.argv[1])
    tot = sum(v for _, v, _ in rows)
    agg = collections.OrderedDict()
    for n, v, _ in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"{len(rows)} launches, {tot:.1f} us total (serialised, cold cache: compare shares)")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:10.1f} us {100 * v / tot:5.1f}%  n={c:3d}  {n[:80]}")
    if "--seq" in sys.argv:
        for i, (n, v, g) in enumerate(rows):
            if "igemm" in n:
                print(i, f"{v:8.1f}", n[:44], g)


if __name__ == "__main__":
    main()
