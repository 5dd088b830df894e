"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list, with the
bench.py kernel family of every kernel and the families' shares (the check the profiling recipe
asks for: ncu's per-launch times are cold-cache and serialised, so the SHARE of a family in the
step must agree with bench.py's CUDA-event numbers, not the absolute time).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/launches_r2_real_1gpu.md"""
import collections
import csv
import re
import sys

FAMILY = [
    ("bk_", "hashagg_insert"), ("fold_i32_kernel", "hashagg_insert"), ("part_", "hashagg_insert"),
    ("rx_", "hashagg_insert / vocab_build"), ("rle_", "hashagg_insert"), ("merge_", "hashagg_insert"),
    ("scan_tiles_kernel", "hashagg_insert"), ("insert_keys_kernel", "hashagg_insert"), ("insert_agg_kernel", "hashagg_insert"),
    ("arm_launch_kernel", "hashagg_insert"), ("table_init_kernel", "hashagg_insert"), ("special_init_kernel", "hashagg_insert"),
    ("table_to_pairs_kernel", "hashagg_insert"), ("rehash_kernel", "hashagg_insert"),
    ("encode_smem_kernel", "encode"), ("encode_kernel", "encode"),
    ("moments_kernel", "moments"), ("moments_reduce_kernel", "moments"), ("moments_init_kernel", "moments"),
    ("transform_kernel", "normalize"), ("export_kernel", "hashagg_export"),
    ("lookup_", "vocab_build"), ("packed_", "vocab_build"), ("small_vocab_kernel", "vocab_build"),
    ("vocab_scalars_kernel", "vocab_build"), ("pack32_kernel", "vocab_build"), ("unpack32_kernel", "vocab_build"),
    ("count_ge_kernel", "vocab_build"), ("slice_", "vocab_build"),
]


def family(name):
    for pat, fam in FAMILY:
        if pat in name:
            return fam
    return "other"


def main(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg, fam = collections.OrderedDict(), collections.Counter()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("nvtb::", "")
        if "at::" in name or name.startswith("cub::") or "elementwise" in name:
            continue                                   # torch's data-generation kernels: outside the step
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
        d = agg.setdefault(name[:70], [0, 0.0])
        d[0] += 1
        d[1] += us
        fam[family(name)] += us
    tot = sum(v[1] for v in agg.values())
    print(f"engine kernels: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms under ncu (cold cache, serialised)\n")
    print("| family | ms | share |\n|---|---:|---:|")
    for k, v in fam.most_common():
        print(f"| {k} | {v / 1e3:.1f} | {100 * v / tot:.1f} % |")
    print("\n| kernel | family | launches | total ms | share | avg us |\n|---|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {family(k)} | {n} | {t / 1e3:.2f} | {100 * t / tot:.1f} % | {t / n:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
