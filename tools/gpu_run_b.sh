#!/bin/bash
# round-2 run B: full GPU test suite + the three bench workloads + their reference arms
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
NVTB_BENCH_DUMP=1 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_criteo.json 2> gpurun_out/bench_criteo.err; echo criteo rc=$?; tail -c 1500 gpurun_out/bench_criteo.err | grep -v "bench dump\] step"; cat gpurun_out/bench_criteo.json
timeout 600 python bench.py --workload hashbucket --steps 5 --warmup 3 --sweep 1e7,1e8,2.5e8 --e2e-rows 50000000 > gpurun_out/bench_hashbucket.json 2> gpurun_out/bench_hashbucket.err; echo hashbucket rc=$?; tail -c 800 gpurun_out/bench_hashbucket.err; cat gpurun_out/bench_hashbucket.json
timeout 600 python bench.py --workload movielens --steps 5 --warmup 3 > gpurun_out/bench_movielens.json 2> gpurun_out/bench_movielens.err; echo movielens rc=$?; tail -c 800 gpurun_out/bench_movielens.err; cat gpurun_out/bench_movielens.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_criteo.json 2>/dev/null; echo ref rc=$?; cat gpurun_out/ref_criteo.json | cut -c1-300
timeout 300 python bench.py --impl reference --workload hashbucket --steps 3 --warmup 1 > gpurun_out/ref_hashbucket.json 2>/dev/null; cat gpurun_out/ref_hashbucket.json | cut -c1-300
timeout 300 python bench.py --impl reference --workload movielens --steps 3 --warmup 1 > gpurun_out/ref_movielens.json 2>/dev/null; cat gpurun_out/ref_movielens.json | cut -c1-300
