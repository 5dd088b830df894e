import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nvtabular_b200 as nvt
from nvtabular_b200 import engine as eng
from nvtabular_b200.synth import criteo_frame, CAT_NAMES
rows = 1 << 26
frame = criteo_frame(rows, device='cuda')
aggs = {n: eng.HashAgg(0) for n in CAT_NAMES}
for step in range(4):
    line = []
    t_all0 = time.perf_counter()
    for n in CAT_NAMES:
        h = aggs[n]
        if step: h.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        h.insert(frame[n])
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        line.append((n, round(e0.elapsed_time(e1) * 1e3), round((t1 - t0) * 1e6)))
    torch.cuda.synchronize()
    print('step', step, 'wall ms', round((time.perf_counter() - t_all0) * 1e3, 1), 'gpu us / host us:', ' '.join(f'{n}:{g}/{c}' for n, g, c in line), flush=True)
