import os, sys, time, torch
os.environ["NVTB_ARTIFACTS"] = "lazy"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nvtabular_b200 as nvt
from nvtabular_b200.synth import criteo_frame
import bench
rows = 1 << 26
dev = torch.device("cuda", 0)
frame = criteo_frame(rows, device="cuda")
wf = bench.build_workflow(nvt, "criteo", "/tmp/nvtb_e2e")
host = bench.host_partitions(frame, 8)
del frame
torch.cuda.synchronize()
def T(): torch.cuda.synchronize(); return time.perf_counter()
# raw H2D
for rep in range(2):
    t0 = T()
    d = [p.to(dev) for p in host]
    t1 = T()
    nb = sum(p.nbytes() for p in host)
    print(f"raw H2D {nb/1e9:.2f} GB in {(t1-t0)*1e3:.1f} ms = {nb/(t1-t0)/1e9:.1f} GB/s", flush=True)
    del d
# big single-buffer copies for reference
big_h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
big_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for rep in range(2):
    t0 = T(); big_d.copy_(big_h, non_blocking=True); t1 = T()
    big_h.copy_(big_d, non_blocking=True); t2 = T()
    print(f"1 GiB H2D {1.0737/(t1-t0):.1f} GB/s, D2H {1.0737/(t2-t1):.1f} GB/s", flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
big_h2 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
big_d2 = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
t0 = T()
with torch.cuda.stream(s1): big_d.copy_(big_h, non_blocking=True)
with torch.cuda.stream(s2): big_h2.copy_(big_d2, non_blocking=True)
t1 = T()
print(f"concurrent 1 GiB each way: {(t1-t0)*1e3:.1f} ms -> {2*1.0737/(t1-t0):.1f} GB/s total", flush=True)
del big_h, big_d, big_h2, big_d2
out_host = None
for step in range(3):
    ds = nvt.Dataset(list(host))
    t0 = T()
    wf.fit(ds)
    t1 = T()
    tds = wf.transform(ds)
    res = tds.to_host(out_host if out_host else None)
    t2 = T()
    out_host = res
    print(f"step {step}: fit {1e3*(t1-t0):.1f} ms (h2d {ds.h2d_bytes/1e9:.2f} GB -> {ds.h2d_bytes/(t1-t0)/1e9:.1f} GB/s), "
          f"transform+to_host {1e3*(t2-t1):.1f} ms (d2h {tds.d2h_bytes/1e9:.2f} GB -> {tds.d2h_bytes/(t2-t1)/1e9:.1f} GB/s)", flush=True)
