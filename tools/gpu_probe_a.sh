#!/bin/bash
# round-2 probe A: e2e jitter (allocator), cluster-fold microbench, launch list at real scale
mkdir -p gpurun_out tools/_bin
cd $GRAFT_REPO_ROOT
echo "== e2e default allocator"
NVTB_BENCH_DUMP=1 timeout 600 python bench.py --rows 67108864 --steps 3 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step"
echo "== e2e expandable segments"
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True NVTB_BENCH_DUMP=1 timeout 600 python bench.py --rows 67108864 --steps 3 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "e2e step"
echo "== cluster fold microbench"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/microbench_cluster_fold.cu -o tools/_bin/mb_cluster && timeout 300 tools/_bin/mb_cluster 2>&1 | tee gpurun_out/mb_cluster.txt
echo "== ncu launch list, real scale"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r2_real.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
echo ncu rc=$?
wc -l gpurun_out/launches_r2_real.csv
