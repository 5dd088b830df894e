#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 300 python tools/debug_te_reload.py 2>&1 | tail -14
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_boundary_gpu.py::test_save_load_roundtrip_in_process_and_fresh_process > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/pytest_gpu.log
NVTB_BENCH_DUMP=1 timeout 1200 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_criteo.json 2> gpurun_out/bench_criteo.err; echo criteo rc=$?; grep -v "bench dump\] step" gpurun_out/bench_criteo.err | tail -8
python - <<'PY'
import json,re
try:
    d=json.load(open('gpurun_out/bench_criteo.json'))
    for k in ['value','ms_per_step','first_fit_ms','artifact_policies','gpu_launches']:
        print(k, d.get(k))
    for k,v in d['kernels'].items(): print(k, {a:round(b,2) for a,b in v.items()})
except Exception as e: print("no json", e)
for line in open('gpurun_out/bench_criteo.err'):
    m=re.match(r"\[bench dump\] step (\d+): (.*)", line)
    if not m or m.group(1) not in ("0","4"): continue
    fam={}
    for it in m.group(2).split():
        k,v=it.split(':'); fam.setdefault(k,[]).append(int(v))
    print("step",m.group(1), {k:(round(sum(v)/1e3,1)) for k,v in fam.items()}, "big vocab:", fam.get('vocild',[])[-5:], "enc max", sorted(fam['encode'])[-6:])
PY
