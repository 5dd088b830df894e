"""GPU timeline of one bench step (torch.profiler / CUPTI activity records): busy time, idle
gaps and which kernels surround the largest gaps.  python tools/timeline.py [rows]"""
import os, sys, collections
os.environ.setdefault("NVTB_ARTIFACTS", "lazy")     # what bench.py runs with by default
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import nvtabular_b200 as nvt
from nvtabular_b200.synth import criteo_frame
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
frame = criteo_frame(rows, device="cuda")
wf = bench.build_workflow(nvt, "/tmp/nvtb_timeline")
for _ in range(3):
    bench.run_step(nvt, wf, frame)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(int(os.environ.get("TIMELINE_STEPS", "3"))):      # back to back, as bench.py times them
        out = bench.run_step(nvt, wf, frame)
        del out
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy, cur_end, gaps = 0.0, t0, []
prev = None
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, prev.name[:50] if prev else "-", e.name[:50]))
        busy += en - s
        cur_end = en
    elif en > cur_end:
        busy += en - cur_end
        cur_end = en
    if prev is None or en >= cur_end:
        prev = e
print(f"span {(t1 - t0) / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms, {len(evs)} device activities")
print("largest single gaps (us): " + "; ".join(f"{g:.0f} [{a.split('(')[0][-28:]} -> {b.split('(')[0][-28:]}]" for g, a, b in sorted(gaps, key=lambda x: -x[0])[:12]))
by = collections.defaultdict(lambda: [0, 0.0])
for g, a, b in gaps:
    k = (a.split("(")[0][:34], b.split("(")[0][:34])
    by[k][0] += 1
    by[k][1] += g
print("largest idle contributors (prev kernel -> next kernel): count, total us")
for k, (c, tot) in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {k[0]:36s} -> {k[1]:36s} {c:5d} {tot:9.1f}")
kt = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    kt[e.name.split("(")[0][:48]][0] += 1
    kt[e.name.split("(")[0][:48]][1] += e.time_range.end - e.time_range.start
print("device time by activity:")
for k, (c, tot) in sorted(kt.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {k:50s} {c:5d} {tot / 1e3:9.3f} ms")
