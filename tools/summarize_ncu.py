"""Turn an `ncu --csv` launch list (gpu__time_duration.sum [+ other metrics]) into the
per-kernel summary committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/launches.csv > profiles/launches_r1.md
"""
import collections
import csv
import sys


# kernel-name substring -> bench.py kernel family (one family "launch" = one C-ABI call,
# e.g. one column's nvtb_hashagg_insert = 1 kernel in direct mode, 4 when partitioned)
FAMILY = [
    ("fold_i32_kernel", "hashagg_insert"), ("part_hist_kernel", "hashagg_insert"),
    ("part_scan_kernel", "hashagg_insert"), ("part_scatter_kernel", "hashagg_insert"),
    ("insert_keys_kernel", "hashagg_insert"), ("insert_agg_kernel", "hashagg_insert"),
    ("arm_launch_kernel", "hashagg_insert"),
    ("encode_smem_kernel", "encode"), ("encode_kernel", "encode"),
    ("moments_kernel", "moments"), ("moments_reduce_kernel", "moments"), ("moments_init_kernel", "moments"),
    ("transform_kernel", "normalize"), ("export_kernel", "hashagg_export"),
    ("DeviceRadixSort", "vocab_build"), ("small_vocab_kernel", "vocab_build"), ("lookup_", "vocab_build"),
    ("vocab_scalars_kernel", "vocab_build"), ("xor_copy_kernel", "vocab_build"), ("count_ge_kernel", "vocab_build"),
]
CALLS_PER_STEP = {"moments": 1, "normalize": 1}      # every other family: one call per categorical column


def _family(name):
    for pat, fam in FAMILY:
        if pat in name:
            return fam
    return None


def _launches(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per_launch = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (int(row["ID"]), row["Kernel Name"])
        try:
            per_launch.setdefault(key, {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    return list(per_launch.items())


def traffic_json(path, rows_per_gpu, n_cat_cols=26):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per family CALL over the LAST
    bench step of the capture (a step ends with the Normalize transform_kernel), for
    bench.py's roofline.traffic."""
    import json
    items = _launches(path)
    ends = [i for i, ((_, n), _) in enumerate(items) if "transform_kernel" in n]
    if len(ends) >= 2:
        items = items[ends[-2] + 1:ends[-1] + 1]
    out = {}
    for (_, name), m in items:
        fam = _family(name)
        if fam is None or "dram__bytes_read.sum" not in m:
            continue
        d = out.setdefault(fam, {"rows_per_gpu": rows_per_gpu, "kernels": 0, "dram_bytes": 0.0, "kernel_us": 0.0})
        d["kernels"] += 1
        d["dram_bytes"] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        d["kernel_us"] += m.get("gpu__time_duration.sum", 0.0) / 1e3
    for fam, d in out.items():
        calls = CALLS_PER_STEP.get(fam, n_cat_cols)
        d["calls_per_step"] = calls
        d["dram_bytes_per_launch"] = d["dram_bytes"] / calls
        d["avg_call_us_under_ncu"] = d["kernel_us"] / calls
    return json.dumps(out, indent=1)


def main(path):
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    fam_ms = collections.defaultdict(float)
    items = _launches(path)
    # the capture starts with bench.py generating the synthetic table (torch RNG / elementwise
    # kernels): the bench steps begin at the first launch of one of the engine's kernels
    first = next((i for i, ((_, n), _) in enumerate(items) if "nvtb::" in n), 0)
    skipped = sum(m.get("gpu__time_duration.sum", 0.0) for _, m in items[:first])
    items = items[first:]
    for (_, name), m in items:
        fam_ms[_family(name) or "other"] += m.get("gpu__time_duration.sum", 0.0)
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
    print("\n| family (bench.py `kernels`) | total ms | share |\n|---|---:|---:|")
    for k, v in sorted(fam_ms.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {v / 1e6:.3f} | {100 * v / total:.1f}% |")
    print(f"\n({first} launches / {skipped / 1e6:.1f} ms of synthetic-table generation before the first engine kernel are not listed)")
    print(f"\ntotal device time of the captured launches: {total / 1e6:.3f} ms "
          f"(cold-cache, serialised by ncu: compare SHARES, not absolutes)")


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        print(traffic_json(sys.argv[1], int(sys.argv[3])))
    else:
        main(sys.argv[1])
