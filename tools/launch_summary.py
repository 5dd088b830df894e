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


_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, traffic=None):
    """traffic = (fits, calls_per_fit, rows_per_gpu): also print the DRAM bytes per hashagg_insert call"""
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg, fam, fam_bytes = collections.OrderedDict(), collections.Counter(), collections.Counter()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("nvtb::", "")
        if "at::" in name or name.startswith("cub::") or "elementwise" in name:
            continue                                   # torch's data-generation kernels: outside the step
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            d = agg.setdefault(name[:70], [0, 0.0, 0.0])
            d[2] += v * _B.get(u, 1.0)
            fam_bytes[family(name)] += v * _B.get(u, 1.0)
            continue
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
        d = agg.setdefault(name[:70], [0, 0.0, 0.0])
        d[0] += 1
        d[1] += us
        fam[family(name)] += us
    tot = sum(v[1] for v in agg.values())
    if traffic:
        import json
        fits, calls, rows = traffic
        b = (fam_bytes["hashagg_insert"] + 0.0) / fits / calls
        print(json.dumps({"criteo": {"hashagg_insert": {"rows_per_gpu": rows, "dram_bytes_per_launch": b,
                                                        "how": f"DRAM read+write bytes of the family's kernels over {fits} fits / {calls} calls per fit"}}}))
        return
    print(f"engine kernels: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms under ncu (cold cache, serialised)\n")
    print("| family | ms | share |\n|---|---:|---:|")
    for k, v in fam.most_common():
        print(f"| {k} | {v / 1e3:.1f} | {100 * v / tot:.1f} % |")
    print("\n| kernel | family | launches | total ms | share | avg us | DRAM GB (rd+wr) |\n|---|---|---:|---:|---:|---:|---:|")
    for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {family(k)} | {n} | {t / 1e3:.2f} | {100 * t / tot:.1f} % | {t / max(n, 1):.1f} | {b / 1e9:.2f} |")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
        main(sys.argv[1], (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])))
    else:
        main(sys.argv[1])
