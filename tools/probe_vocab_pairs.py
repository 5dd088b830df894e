"""time engine.Vocab.build_from_pairs / build_from_agg at the vocabulary sizes of the N=1 and N=8
Criteo runs (1.67e8 and 2.9e8 keys)"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nvtabular_b200 import engine
from nvtabular_b200.column import Column
torch.cuda.set_device(0)
for n in (167_000_000, 290_000_000):
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.randperm(n, device="cuda", generator=g).to(torch.int64) * 7 + 3          # distinct
    keys = (keys & 0x7FFFFFFF) | ((torch.arange(n, device="cuda") % 2) << 31)             # spread over 32 bits (still distinct: odd multiplier)
    cnt = torch.randint(1, 50, (n,), device="cuda", generator=g, dtype=torch.int64)
    cnt, _ = torch.sort(cnt, descending=True)
    pairs = ((keys & 0xFFFFFFFF) << 32) | cnt
    del keys, cnt
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        v = engine.Vocab.build_from_pairs(pairs, 0)
        e1.record()
        torch.cuda.synchronize()
        print(f"n={n} rep {rep}: device {e0.elapsed_time(e1):.2f} ms, host {1e3 * (time.perf_counter() - t0):.2f} ms, kept {v.n_kept}", flush=True)
        del v
    del pairs
    torch.cuda.empty_cache()
