#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload movielens --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --pyprofile > gpurun_out/bench_movielens.json 2> gpurun_out/movielens_pyprofile.txt; echo movielens rc=$?; head -60 gpurun_out/movielens_pyprofile.txt | cut -c1-200
echo "== ncu launch list (criteo, real scale)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r2_real.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-gate > gpurun_out/launches_bench.log 2>&1; echo ncu rc=$?; wc -l gpurun_out/launches_r2_real.csv
echo "== ncu --set full on the hot kernels at real per-column sizes"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"_kernel" -c 120 -o gpurun_out/r2_kernels python tools/profile_kernels.py --rows 62500000 --cards 290000000,39043,1543 --batches 2 --reps 1 > gpurun_out/profile_kernels.log 2>&1; echo ncu-full rc=$?; tail -3 gpurun_out/profile_kernels.log
ls -la gpurun_out/*.ncu-rep
ncu -i gpurun_out/r2_kernels.ncu-rep --page raw --csv > gpurun_out/r2_kernels_raw.csv 2>/dev/null; wc -c gpurun_out/r2_kernels_raw.csv
python tools/ncu_kernels_md.py gpurun_out/r2_kernels.ncu-rep > gpurun_out/r2_kernels.md 2>/dev/null; head -5 gpurun_out/r2_kernels.md | cut -c1-300
# keep the report only if it fits the 64 MiB return budget
sz=$(stat -c %s gpurun_out/r2_kernels.ncu-rep); if [ "$sz" -gt 45000000 ]; then rm gpurun_out/r2_kernels.ncu-rep; echo "report too large ($sz), kept csv + md"; fi
