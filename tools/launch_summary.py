"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path, top=45):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        t = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        t_us = t / 1000.0 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1000.0)
        name = row["Kernel Name"].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += t_us
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"{n} launches, {tot / 1000:.2f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'share':>7} {'launches':>8} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] / tot * 100:6.2f}% {v[0]:8d} {v[1] / v[0]:9.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
