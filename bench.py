#!/usr/bin/env python
"""bench.py — rows/sec through the Categorify + FillMissing + Normalize workflow
on a synthetic Criteo-1TB-shaped table (13 int + 26 categorical columns),
BASELINE.json configs[1] on 1 GPU and its key-hash-sharded form on N GPUs.

    python bench.py --gpus N --steps K --warmup W          # this repo's engine
    python bench.py --impl reference --gpus N ...          # the reference's CPU path (oracle port)

A "step" = Workflow.fit(dataset) + Workflow.transform(dataset) over the whole
resident table (every kernel, NCCL call and host sync of fit and transform).
Prints ONE JSON line (see the contract in the task statement): `value` is
device-resident throughput, `e2e` the same workflow fed from pinned HOST buffers
with the H2D / D2H copies inside the timed region, `roofline` the dominant
kernel against the measured HBM peak, `cpu_baseline` the CPU oracle timed on
this box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_ROW = 641.75      # fit 160.875 + transform 480.875 (int64 labels, f64 outputs)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(kernel, rows):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed ncu capture (profiles/traffic_r1.json; produced by tools/summarize_ncu.py).
    Only valid for the row count it was captured at; null otherwise."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r1.json")) as f:
            d = json.load(f)
        e = d.get(kernel)
        if e and int(e.get("rows_per_gpu", -1)) == int(rows):
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  ONE nvidia-smi
    process, started before the warm-up (its start-up takes driver-wide locks for hundreds
    of ms and would otherwise land inside a ~150 ms timed region) and left polling every
    50 ms; stop(t0, t1) keeps the samples whose own timestamps fall inside the region."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("NVTB_SMI_MS", "50"),
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                rows.append((self._epoch(f[0]), float(f[2]), float(f[3]), f[6:10]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and r[0] is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        window = "timed region"
        if not inside:          # region shorter than the polling period: nearest samples
            inside, window = rows[-3:], "nearest samples (region shorter than the 50 ms poll)"
        sm, mx, reasons = [r[1] for r in inside], [r[2] for r in inside], set()
        for r in inside:
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def build_workflow(nvt, out_path, int32_outputs=False):
    from nvtabular_b200.synth import CAT_NAMES, CONT_NAMES
    ops = nvt.ops
    cat_kw = {"dtype": "int32"} if int32_outputs else {}
    norm_kw = {"out_dtype": "float32"} if int32_outputs else {}
    cats = CAT_NAMES >> ops.Categorify(out_path=out_path, **cat_kw)
    conts = CONT_NAMES >> ops.FillMissing() >> ops.Normalize(**norm_kw)
    return nvt.Workflow(cats + conts + ["label"])


def device_partitions(frame, nparts):
    """the resident table cut into `nparts` row partitions (views, 64-row aligned)"""
    rows = len(frame)
    chunk = ((rows + nparts - 1) // nparts + 63) // 64 * 64
    return [frame.slice_rows(s, min(rows, s + chunk)) for s in range(0, rows, chunk)]


def run_step(nvt, wf, parts):
    """one pass of the hot path over the resident table: fit (accumulated over the
    partitions), then transform partition by partition (outputs of one partition live at a time)"""
    ds = nvt.Dataset(list(parts))
    wf.fit(ds)
    out = None
    for part in wf.transform(ds).partitions():
        out = part
    return out


def host_partitions(frame, nparts):
    """pinned-host mirror of a device frame, cut into `nparts` row partitions"""
    rows = len(frame)
    chunk = ((rows + nparts - 1) // nparts + 63) // 64 * 64
    return [frame.slice_rows(s, min(rows, s + chunk)).pin() for s in range(0, rows, chunk)]


def run_step_e2e(nvt, wf, host_parts, out_host):
    """The same step from HOST buffers through the public API: Dataset of pinned host
    partitions -> Workflow.fit -> Workflow.transform -> pinned host results.  Every input
    byte crosses PCIe once (partitions are prefetched one ahead and stay in HBM between
    fit and transform), every output byte crosses it once (D2H overlapped with the next
    partition's kernels)."""
    trace = os.environ.get("NVTB_BENCH_DUMP")
    t0 = time.perf_counter()
    ds = nvt.Dataset(list(host_parts))
    wf.fit(ds)
    if trace:
        import torch
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    tds = wf.transform(ds)
    res = tds.to_host(out_host if out_host else None)
    if trace:
        sys.stderr.write("[bench dump] e2e step: fit %.1f ms, transform+to_host %.1f ms\n"
                         % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
    return ds.h2d_bytes, tds.d2h_bytes, res


def cpu_reference(rows, workers, seed=1234, steps=1, warmup=0):
    """The reference's CPU path (oracle port, oracle/parallel.py) on a bounded sample of the
    same synthetic workload.  Data is generated with torch on the CPU (same generator code)."""
    import pandas as pd  # noqa: F401
    from nvtabular_b200.synth import CAT_NAMES, CONT_NAMES, criteo_frame, frame_to_pandas_nullable
    from oracle.parallel import run_criteo_workflow
    frame = criteo_frame(rows, total_rows=rows, seed=seed, device="cpu")
    df = frame_to_pandas_nullable(frame)
    # what pandas itself holds for a nullable int column read from parquet: float64 + NaN
    df = df.astype({c: "float64" for c in CAT_NAMES + CONT_NAMES})
    times = []
    for i in range(warmup + steps):
        tf, tt, *_ = run_criteo_workflow(df, CAT_NAMES, CONT_NAMES, workers)
        if i >= warmup:
            times.append(tf + tt)
    sec = sum(times) / len(times)
    return rows / sec, sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=250_000_000, help="rows resident per GPU (SURVEY 8d C2: 2.5e8)")
    ap.add_argument("--parts", type=int, default=4, help="device-resident partitions the table is cut into")
    ap.add_argument("--profile-rows", type=int, default=4_370_000_000,
                    help="row count the categorical cardinalities are scaled to (4.37e9 = the full Criteo-1TB profile)")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows per e2e step (default: same table)")
    ap.add_argument("--e2e-parts", type=int, default=8, help="host partitions per e2e step")
    ap.add_argument("--cpu-rows", type=int, default=1 << 20, help="rows of the bounded CPU sample")
    ap.add_argument("--int32-outputs", action="store_true", help="Categorify(dtype=int32), Normalize(float32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eager-artifacts", action="store_true",
                    help="write the vocabulary parquet files inside fit (default: deferred until read)")
    ap.add_argument("--pyprofile", action="store_true", help="cProfile one extra step to stderr")
    args = ap.parse_args()
    if not args.eager_artifacts:
        os.environ["NVTB_ARTIFACTS"] = "lazy"
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        rows = max(args.cpu_rows * 8, 1 << 23)
        value, sec = cpu_reference(rows, cores, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
        line = {
            "impl": "reference", "metric": "rows/sec Criteo-1TB-shaped Categorify+FillMissing+Normalize",
            "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "criteo-shaped 13 int + 26 cat, Categorify+FillMissing+Normalize, fit+transform",
                       "rows_per_step": rows, "impl": "oracle port of the reference pandas path, partition-parallel"},
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": f"{rows} rows of the same synthetic table, fit+transform"},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import nvtabular_b200 as nvt
    from nvtabular_b200 import engine
    from nvtabular_b200.synth import criteo_frame

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    rows = args.rows
    total_rows = rows * world
    frame = criteo_frame(rows, total_rows=args.profile_rows, device=dev, rank=rank)
    frame = device_partitions(frame, args.parts)
    out_dir = f"/tmp/nvtb_bench_rank{rank}"
    wf = build_workflow(nvt, out_dir, args.int32_outputs)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.pyprofile and rank == 0:
        import cProfile
        import pstats
        run_step(nvt, wf, frame)
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        run_step(nvt, wf, frame)
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(35)

    # ---------------- device-resident timing --------------------------------------
    sampler = ClockSampler(local_rank)
    t_sampler = time.time()
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        out = run_step(nvt, wf, frame)
        del out
    if rank == 0 and time.time() - t_sampler < 1.5:      # nvidia-smi start-up must be over
        time.sleep(1.5 - (time.time() - t_sampler))
    sync_all()
    engine.profile = []
    t_wall0 = time.time()
    launches0 = engine.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = run_step(nvt, wf, frame)
        del out
    ev1.record()
    sync_all()
    prof = engine.profile
    engine.profile = None
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    launches = engine.kernel_launches - launches0
    ms_per_step = elapsed_ms / args.steps
    value = total_rows / (ms_per_step / 1e3)

    # per-kernel-family device time (CUDA events on the launching stream)
    fam = {}
    for family, s, e, nbytes in prof:
        d = fam.setdefault(family, {"ms": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += s.elapsed_time(e)
        d["bytes"] += nbytes
        d["launches"] += 1
    if os.environ.get("NVTB_BENCH_DUMP") and rank == 0:
        per_step = len(prof) // max(1, args.steps)
        for st in range(args.steps):
            sys.stderr.write("[bench dump] step %d: %s\n" % (st, " ".join(
                "%s:%.0f" % (f[:3] + f[-3:], s.elapsed_time(e) * 1e3)
                for f, s, e, _ in prof[st * per_step:(st + 1) * per_step])))
    peak, peak_src = _peaks()
    kernels = {}
    for k, d in fam.items():
        gbs = d["bytes"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else 0.0
        kernels[k] = {"ms_per_step": d["ms"] / args.steps, "launches_per_step": d["launches"] / args.steps,
                      "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    dominant = max(fam, key=lambda k: fam[k]["ms"]) if fam else None
    roofline = None
    if dominant:
        d = fam[dominant]
        ach = d["bytes"] / (d["ms"] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": _traffic(dominant, rows), "peak_source": peak_src,
                    "avg_launch_ms": d["ms"] / d["launches"],
                    "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
                    "share_of_step": (d["ms"] / args.steps) / ms_per_step,
                    "whole_step_gbs": ALGO_BYTES_PER_ROW * rows / (ms_per_step / 1e3) / 1e9,
                    "whole_step_frac": ALGO_BYTES_PER_ROW * rows / (ms_per_step / 1e3) / 1e9 / peak}

    # ---------------- the same step with the reference's eager artefacts ----------------
    # `value` runs with NVTB_ARTIFACTS=lazy (config.artifacts); this leg reports what the step
    # costs when every fit also writes unique.<col>.parquet / meta.<col>.parquet like the
    # reference does (host-side pandas/pyarrow work, GPU idle meanwhile).  Single GPU only;
    # a failure here never affects the other numbers.
    eager = None
    if world == 1 and not args.eager_artifacts and not args.no_e2e:
        try:
            os.environ["NVTB_ARTIFACTS"] = "eager"
            out = run_step(nvt, wf, frame)
            del out
            torch.cuda.synchronize()
            n_eager = max(1, min(args.steps, 3))
            ev0.record()
            for _ in range(n_eager):
                out = run_step(nvt, wf, frame)
                del out
            ev1.record()
            torch.cuda.synchronize()
            e_ms = ev0.elapsed_time(ev1) / n_eager
            eager = {"ms_per_step": e_ms, "value": total_rows / (e_ms / 1e3), "unit": "rows/s", "steps": n_eager,
                     "what": "library default: every meta.<col>.parquet and the unique.<col>.parquet of every "
                             "vocabulary up to 2^20 keys written during fit"}
        except Exception as exc:          # noqa: BLE001 - reported, not fatal
            eager = {"error": repr(exc)[:200]}
        finally:
            os.environ["NVTB_ARTIFACTS"] = "lazy"

    # ---------------- end to end from pinned host buffers ---------------------------
    e2e = None
    if not args.no_e2e:
        e_rows = args.e2e_rows or rows
        src = frame if e_rows == rows else criteo_frame(e_rows, total_rows=total_rows, device=dev, rank=rank)
        host = host_partitions(src, args.e2e_parts)
        del src
        if not os.environ.get("NVTB_BENCH_KEEP_FRAME"):
            # the device-resident table and the allocator blocks cached by the timed region above
            # are not part of the e2e leg: it starts from host buffers and a clean device pool
            frame = None
            torch.cuda.empty_cache()
        out_host = None
        # W >= 3 warm-up steps here too: the first e2e step pins ~21 GB of result buffers (seconds),
        # the second still grows the device allocator's pools
        for _ in range(max(3, args.warmup)):
            h2d, d2h, out_host = run_step_e2e(nvt, wf, host, out_host)
        sync_all()
        e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(e_steps):
            h2d, d2h, out_host = run_step_e2e(nvt, wf, host, out_host)
        ev1.record()
        sync_all()
        e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item()) / e_steps
        e2e = {"value": e_rows * world / (e_ms / 1e3), "unit": "rows/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": e_ms, "rows_per_step": e_rows * world,
               "host_partitions": len(host),
               "warmup": max(3, args.warmup), "steps": e_steps,
               # fit needs every partition before the first label exists, so H2D and D2H of one
               # step cannot overlap: the bound is their SUM at the ~55 GB/s one PCIe 5 x16 sustains
               "pcie_serial_bound_ms": (h2d + d2h) / 55e9 * 1e3}
        del host, out_host

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec = cpu_reference(args.cpu_rows, 1)
        cpu_baseline = {"value": v, "unit": "rows/s", "cores": 1, "kind": "port",
                        "sample": f"{args.cpu_rows} rows of the same synthetic table, fit+transform, "
                                  f"{sec:.1f} s on 1 core (host has {os.cpu_count()})"}

    line = {
        "metric": "rows/sec Criteo-1TB-shaped Categorify+FillMissing+Normalize",
        "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64" if not args.int32_outputs else "int32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: Criteo-1TB-shaped synthetic (13 int + 26 cat int32, "
                               "nullable), Categorify+FillMissing+Normalize, one step = Workflow.fit + "
                               "Workflow.transform over the HBM-resident table",
                   "rows_per_gpu": rows, "total_rows": total_rows,
                   "outputs": "int32 labels, float32 conts" if args.int32_outputs else "int64 labels, float64 conts",
                   "algorithmic_bytes_per_row": ALGO_BYTES_PER_ROW if not args.int32_outputs else 485.75,
                   "cache": "inputs (%.1f GB/GPU) larger than L2; no explicit flush" % (rows * 160.9 / 1e9),
                   "artifacts": "eager" if args.eager_artifacts else
                                "unique./meta. parquet files deferred until read (NVTB_ARTIFACTS=lazy)",
                   "parallelism": f"row-sharded x{world}, key-hash owner merge over NCCL" if world > 1 else "single GPU"},
        "roofline": roofline, "kernels": kernels, "e2e": e2e, "eager_artifacts": eager,
        "cpu_baseline": cpu_baseline,
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
