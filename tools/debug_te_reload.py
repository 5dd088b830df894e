import os, sys, numpy as np, pandas as pd, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nvtabular as nvt
rng = np.random.default_rng(3)
n = 20000
df = pd.DataFrame({"a": rng.integers(0, 500, n).astype(np.int32), "s": rng.choice(["x", "yy", "zzz", "w"], n),
                   "t": rng.integers(0, 2, n).astype(np.float32)})
te = ["a", ["a", "s"]] >> nvt.ops.TargetEncoding("t", kfold=3, p_smooth=10, out_path="/tmp/te_dbg")
wf = nvt.Workflow(te)
exp = wf.fit_transform(nvt.Dataset(df, npartitions=2)).to_ddf().compute()
wf.save("/tmp/te_dbg_saved")
wf2 = nvt.Workflow.load("/tmp/te_dbg_saved")
got = wf2.transform(nvt.Dataset(df, npartitions=2)).to_ddf().compute()
op1 = [x.op for x in wf.output_node.topo_order() if x.kind == "op"][0]
op2 = [x.op for x in wf2.output_node.topo_order() if x.kind == "op"][0]
print("means", op1.means, op2.means, "stats2", op2.stats)
for name in ["a", "a_s"]:
    g1, g2 = op1._groups[name], op2._groups[name]
    for tag, (x1, x2) in {"all": (g1._all, g2._all), "fold": (g1._fold, g2._fold)}.items():
        k1, k2 = x1[0].cpu().numpy(), x2[0].cpu().numpy()
        o1, o2 = np.argsort(k1), np.argsort(k2)
        print(name, tag, "n", len(k1), len(k2), "keys equal", np.array_equal(k1[o1], k2[o2]))
        c1, c2 = x1[1].cpu().numpy(), x2[1].cpu().numpy()
        if tag == "all":
            print("   counts equal", np.array_equal(c1[:-1][o1], c2[:-1][o2]), "null", c1[-1], c2[-1],
                  "sums", np.allclose(x1[2].cpu().numpy()[:-1][o1], x2[2].cpu().numpy()[:-1][o2]))
        else:
            print("   counts equal", np.array_equal(c1[o1], c2[o2]), "sums", np.allclose(x1[2].cpu().numpy()[o1], x2[2].cpu().numpy()[o2]))
            if not np.array_equal(k1[o1], k2[o2]):
                print("   first keys", k1[o1][:5], k2[o2][:5], "last", k1[o1][-3:], k2[o2][-3:])
for c in exp.columns:
    bad = np.nonzero(got[c].to_numpy() != exp[c].to_numpy())[0]
    print(c, "mismatch", len(bad), bad[:10], "parts split at", -(-n // 2))
fdf = pd.read_parquet(op2.stats["__fold___a"]); print(fdf.dtypes, fdf.head(3), fdf["__fold__"].value_counts().to_dict())

# ---- record what transform feeds pack_keys2 per partition for both workflows
from nvtabular_b200 import engine
calls = []
orig = engine.pack_keys2
def rec(a, b):
    out = orig(a, b)
    calls.append((a.data.cpu().numpy().copy(), b.data.cpu().numpy().copy(), out.data.cpu().numpy().copy()))
    return out
engine.pack_keys2 = rec
import nvtabular_b200.ops.target_encoding as te_mod
te_mod.engine.pack_keys2 = rec
calls.clear(); _ = wf.transform(nvt.Dataset(df, npartitions=2)).to_ddf().compute(); c1 = list(calls)
calls.clear(); _ = wf2.transform(nvt.Dataset(df, npartitions=2)).to_ddf().compute(); c2 = list(calls)
print("calls", len(c1), len(c2))
for i, (x, y) in enumerate(zip(c1, c2)):
    print(i, "fold eq", np.array_equal(x[0], y[0]), "gid eq", np.array_equal(x[1], y[1]), "packed eq", np.array_equal(x[2], y[2]),
          "n", len(x[0]), "gid1", x[1][:6], "gid2", y[1][:6])
