#!/bin/bash
# 8-GPU check of the default bench (no e2e leg) with the merge trace
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l; cat /sys/fs/cgroup/memory.max
NVTB_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo bench rc=$?
grep -v "^\[nvtb trace\]\|OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_8gpu.err | tail -25
grep "nvtb trace" gpurun_out/bench_8gpu.json | tail -7 | cut -c1-300
python - <<'PY'
import json
try:
    lines=[l for l in open('gpurun_out/bench_8gpu.json') if l.startswith('{')]
    d=json.loads(lines[-1])
    for k in ['value','ms_per_step','first_fit_ms','parity_gate','gpu_launches','clocks']:
        print(k, d.get(k))
    for k,v in d['kernels'].items(): print(k, {a:round(b,2) for a,b in v.items()})
except Exception as e: print("no json", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv | tail -8
