"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time per step."""
import collections
import csv
import sys


def main(path, steps):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for row in rows:
        k = row["Kernel Name"][:100]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    if steps is None:
        # one Adam launch per training iteration: the list itself says how many iterations it holds (the round-1
        # summary of a 3-iteration capture was printed with the default of 1, every us/step in it 3x too large)
        steps = max(1, sum(a[0] for k, a in agg.items() if "adam_multi" in k))
    tot = sum(a[1] for a in agg.values())
    print(f"launches/step {len(rows) / steps:.1f}   sum of kernel durations {tot / steps:.1f} us/step ({steps} steps)")
    print(f"{'us/step':>10} {'share':>6} {'n/step':>6}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1] / steps:10.1f} {100 * a[1] / tot:5.1f}% {a[0] / steps:6.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
