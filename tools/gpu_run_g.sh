#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
for g in 128 64 32; do
  echo "== L2 fetch granularity $g"
  NVTB_L2_FETCH=$g NVTB_BENCH_DUMP=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-gate --no-cpu-baseline > gpurun_out/bench_l2_$g.json 2> gpurun_out/bench_l2_$g.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_l2_$g.json'))
print("ms_per_step", round(d['ms_per_step'],1), {k: round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
PY
done
timeout 600 python -m pytest tests/test_boundary_gpu.py tests/test_shape_gpu.py -q > gpurun_out/pytest_gpu_b.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_b.log
