#!/bin/bash
# round-2 run B1: full GPU test suite + the default bench (all legs)
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
NVTB_BENCH_DUMP=1 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_criteo.json 2> gpurun_out/bench_criteo.err; echo criteo rc=$?; grep -v "bench dump\] step" gpurun_out/bench_criteo.err | tail -20; cat gpurun_out/bench_criteo.json
free -g | head -2
