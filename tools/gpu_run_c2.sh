#!/bin/bash
# 2-GPU run: distributed parity test + boundary tests + the default bench on 2 ranks with the merge trace
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
nvidia-smi -L; cat /sys/fs/cgroup/memory.max
timeout 900 python -m pytest tests/test_dist_gpu.py tests/test_boundary_gpu.py tests/test_shape_gpu.py -q -x > gpurun_out/pytest_2gpu.log 2>&1; echo pytest rc=$?; tail -25 gpurun_out/pytest_2gpu.log
NVTB_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo bench rc=$?
grep -v "^\[nvtb trace\]" gpurun_out/bench_2gpu.err | tail -15
grep "nvtb trace" gpurun_out/bench_2gpu.json gpurun_out/bench_2gpu.err | tail -14 | cut -c1-400
python - <<'PY'
import json
try:
    lines=[l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')]
    d=json.loads(lines[-1])
    for k in ['value','ms_per_step','first_fit_ms','parity_gate','e2e','gpu_launches']:
        print(k, d.get(k))
    for k,v in d['kernels'].items(): print(k, {a:round(b,2) for a,b in v.items()})
except Exception as e: print("no json", e)
PY
