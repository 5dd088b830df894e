"""Compact per-kernel table from an `ncu --set full` report (the summaries committed under
profiles/).   python tools/ncu_kernels_md.py report.ncu-rep > profiles/xyz.md"""
import csv
import subprocess
import sys

COLS = [
    ("us", "gpu__time_duration.sum", lambda v: f"{v:.0f}"),
    ("DRAM rd MB", "dram__bytes_read.sum", lambda v: f"{v:.0f}"),
    ("DRAM wr MB", "dram__bytes_write.sum", lambda v: f"{v:.0f}"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", lambda v: f"{v:.0f}"),
    ("L1/TEX %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", lambda v: f"{v:.0f}"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", lambda v: f"{v:.0f}"),
    ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", lambda v: f"{v:.0f}"),
    ("warp Minst", "smsp__inst_executed.sum", lambda v: f"{v / 1e6:.1f}"),
    ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio", lambda v: f"{v:.1f}"),
    ("regs", "launch__registers_per_thread", lambda v: f"{v:.0f}"),
    ("grid", "launch__grid_size", lambda v: f"{v:.0f}"),
]
STALL = "smsp__average_warps_issue_stalled_"


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


_TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
_BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main(path):
    """path: a .ncu-rep (needs ncu) or the `ncu -i rep --page raw --csv` output saved as .csv"""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
    print("| # | kernel | " + " | ".join(c[0] for c in COLS) + " | smem bank-conflict wavefronts | top stalls (warps per issue slot) |")
    print("|---|---|" + "---:|" * len(COLS) + "---:|---|")
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("nvtb::", "")
        cells = []
        for _, key, fmt in COLS:
            v = num(r[ix[key]]) if key in ix else float("nan")
            if key in ix:
                u = units[ix[key]]
                if key == "gpu__time_duration.sum":
                    v = v * _TIME.get(u, 1.0)            # -> us
                elif key.startswith("dram__bytes"):
                    v = v * _BYTES.get(u, 1e-6)          # -> MB
            cells.append(fmt(v))
        conf = num(r[ix.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0)])
        wav = num(r[ix.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0)])
        st = sorted(((num(r[ix[h]]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)
        st = [f"{n} {v:.1f}" for v, n in st if n != "selected"][:3]
        print(f"| {r[ix['ID']]} | `{name}` | " + " | ".join(cells) +
              f" | {conf / 1e6:.1f}M of {wav / 1e6:.1f}M | {', '.join(st)} |")


if __name__ == "__main__":
    main(sys.argv[1])
