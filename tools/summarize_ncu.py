"""Turn an `ncu --csv` launch list (gpu__time_duration.sum [+ other metrics]) into the
per-kernel summary committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/launches.csv > profiles/launches_r1.md
"""
import collections
import csv
import sys


FAMILY = {"insert_keys_kernel": "hashagg_insert", "encode_kernel": "encode", "moments_kernel": "moments",
          "transform_kernel": "normalize", "export_kernel": "hashagg_export"}


def traffic_json(path, rows_per_gpu, skip_to_last_step=True):
    """per-kernel-family average DRAM bytes per launch (for bench.py's roofline.traffic)"""
    import json
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per_launch = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (int(row["ID"]), row["Kernel Name"])
        try:
            per_launch.setdefault(key, {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    items = list(per_launch.items())
    if skip_to_last_step:
        enc = [i for i, ((_, n), _) in enumerate(items) if "encode_kernel" in n]
        if len(enc) >= 52:
            items = items[enc[len(enc) - 27] + 2:]
    out = {}
    for pat, fam in FAMILY.items():
        ms = [m for (_, n), m in items if pat in n]
        if ms and "dram__bytes_read.sum" in ms[0]:
            out[fam] = {"rows_per_gpu": rows_per_gpu, "launches": len(ms),
                        "dram_bytes_per_launch": sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in ms) / len(ms),
                        "avg_launch_us": sum(m["gpu__time_duration.sum"] for m in ms) / len(ms) / 1e3}
    return json.dumps(out, indent=1)


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
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        print(traffic_json(sys.argv[1], int(sys.argv[3])))
    else:
        main(sys.argv[1])
