#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r2_criteo_1gpu.json 2> gpurun_out/bench_r2_criteo_1gpu.err; echo criteo rc=$?; tail -3 gpurun_out/bench_r2_criteo_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_criteo_reference.json 2>/dev/null; echo ref rc=$?
timeout 400 python bench.py --workload hashbucket --sweep 1e7,1e8,2.5e8 > gpurun_out/bench_r2_hashbucket_1gpu.json 2> gpurun_out/bench_r2_hashbucket.err; echo hashbucket rc=$?; tail -2 gpurun_out/bench_r2_hashbucket.err
timeout 200 python bench.py --workload hashbucket --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_hashbucket_reference.json 2>/dev/null
timeout 400 python bench.py --workload movielens > gpurun_out/bench_r2_movielens_1gpu.json 2> gpurun_out/bench_r2_movielens.err; echo movielens rc=$?; tail -2 gpurun_out/bench_r2_movielens.err
timeout 200 python bench.py --workload movielens --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_movielens_reference.json 2>/dev/null
python - <<'PY'
import json
for f in ["criteo_1gpu","criteo_reference","hashbucket_1gpu","hashbucket_reference","movielens_1gpu","movielens_reference"]:
    try:
        d=json.load(open(f"gpurun_out/bench_r2_{f}.json"))
        print(f, "value %.4g" % d["value"], "ms %.1f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "first_fit", d.get("first_fit_ms"),
              "roofline", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
        if d.get("kernels"): print("   ", {k: round(v["ms_per_step"],1) for k,v in d["kernels"].items()})
        if d.get("sweep"): print("   sweep", [(s["rows_per_gpu"], round(s["frac_of_hbm_peak"],3)) for s in d["sweep"]])
        if d.get("artifact_policies"): print("   ", d["artifact_policies"])
    except Exception as e: print(f, "ERR", e)
PY
echo "== ncu launch list with DRAM bytes (criteo, real scale)"
timeout 280 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"_kernel" -c 6000 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-gate > gpurun_out/launches_bench.log 2>&1; echo ncu rc=$?; wc -l gpurun_out/launches_r2_final.csv
