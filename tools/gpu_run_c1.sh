#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/pytest_gpu.log
NVTB_BENCH_DUMP=1 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_criteo.json 2> gpurun_out/bench_criteo.err; echo criteo rc=$?; grep -v "bench dump\] step" gpurun_out/bench_criteo.err | tail -12
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_criteo.json'))
    for k in ['value','ms_per_step','first_fit_ms','parity_gate','artifact_policies','e2e','gpu_launches']:
        print(k, d.get(k))
    for k,v in d['kernels'].items(): print(k, {a:round(b,2) for a,b in v.items()})
except Exception as e: print("no json", e)
PY
timeout 600 python bench.py --workload hashbucket --steps 5 --warmup 3 --sweep 1e7,1e8,2.5e8 > gpurun_out/bench_hashbucket.json 2> gpurun_out/bench_hashbucket.err; echo hashbucket rc=$?; tail -c 600 gpurun_out/bench_hashbucket.err; cat gpurun_out/bench_hashbucket.json | cut -c1-3000
timeout 600 python bench.py --workload movielens --steps 5 --warmup 3 > gpurun_out/bench_movielens.json 2> gpurun_out/bench_movielens.err; echo movielens rc=$?; tail -c 600 gpurun_out/bench_movielens.err; cat gpurun_out/bench_movielens.json | cut -c1-3000
