"""Turn an `ncu --csv` launch list (gpu__time_duration.sum [+ other metrics]) into the
per-kernel summary committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/launches.csv > profiles/launches_r1.md
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per_launch = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (int(row["ID"]), row["Kernel Name"])
        try:
            per_launch.setdefault(key, {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (_, name), m in per_launch.items():
        short = name.split("(")[0].replace("void ", "")[:72]
        a = agg[short]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"| kernel | launches | total ms | share | avg us | avg threads/inst |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.1f}% | {a[1] / a[0] / 1e3:.1f} | "
              f"{a[2] / a[0]:.1f} |")
    print(f"\ntotal device time of the captured launches: {total / 1e6:.3f} ms "
          f"(cold-cache, serialised by ncu: compare SHARES, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1])
