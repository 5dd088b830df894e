"""N>1 host-side logic on CPU: two gloo ranks run the product's key-hash owner
exchange (all-to-all -> owner merge -> all-gather) and must both end with the
exact global group-by of the union of their rows."""
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_owner_exchange_two_ranks_gloo(tmp_path):
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = [json.load(open(tmp_path / f"rank{i}.json")) for i in range(2)]
    # identical on every rank
    for key in ("keys", "sizes", "null_size", "allgather_var"):
        assert res[0][key] == res[1][key], key
    np.testing.assert_allclose(res[0]["vals"], res[1]["vals"], rtol=0, atol=0)
    # equals the global groupby over the union of both ranks' rows
    df = pd.DataFrame({"k": res[0]["local_keys"] + res[1]["local_keys"], "x": res[0]["local_x"] + res[1]["local_x"]})
    df["x2"] = df["x"] ** 2
    g = df.groupby("k").agg(size=("x", "size"), s=("x", "sum"), s2=("x2", "sum"), mn=("x", "min"), mx=("x", "max"))
    assert res[0]["keys"] == g.index.tolist() and res[0]["sizes"] == g["size"].tolist()
    v = np.array(res[0]["vals"])
    np.testing.assert_allclose(v[:, 0], g["s"].to_numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(v[:, 1], g["s2"].to_numpy(), rtol=1e-12)
    np.testing.assert_array_equal(v[:, 2], g["mn"].to_numpy())
    np.testing.assert_array_equal(v[:, 3], g["mx"].to_numpy())
    assert res[0]["null_size"] == 7 + 8
    np.testing.assert_allclose(res[0]["null_vals"], [[3.0, 4.0, -4.0, 5.0]])
    assert res[0]["allgather_var"] == [0, 1, 0, 1, 2]
    # batched multi-table merge: identical on both ranks and equal to the union's value_counts
    assert res[0]["many"] == res[1]["many"]
    for c in range(5):
        vc = pd.Series(res[0]["many_local"][c] + res[1]["many_local"][c]).value_counts().sort_index()
        assert res[0]["many"][c]["keys"] == vc.index.tolist()
        assert res[0]["many"][c]["sizes"] == vc.tolist()
        assert res[0]["many"][c]["null"] == 2 * c + 1
    # key-range exchange of sorted pairs: identical on both ranks, and exactly the union's
    # value_counts in (count desc, key asc) order — the order the vocabulary is built in
    assert res[0]["sorted"] == res[1]["sorted"]
    for c in range(2):
        vc = pd.Series(res[0]["sorted_local"][c] + res[1]["sorted_local"][c]).value_counts(sort=False)
        exp = pd.DataFrame({"k": vc.index.to_numpy(), "s": vc.to_numpy()}).sort_values(
            ["s", "k"], ascending=[False, True], kind="stable")
        assert res[0]["sorted"][c]["keys"] == exp["k"].tolist()
        assert res[0]["sorted"][c]["sizes"] == exp["s"].tolist()
        assert res[0]["sorted"][c]["null"] == 2 * (3 + c) + 1


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [3, 4, 8])
def test_sorted_exchange_world_sizes(tmp_path, world):
    """dist.global_merge_sorted on 3 (odd merge tree, one empty shard) 4 and 8 ranks (what the scaling run
    uses): every rank ends with the union's value_counts in (count desc, key asc) order."""
    port = 29900 + (os.getpid() + world) % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_sorted_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = [json.load(open(tmp_path / f"rank{i}.json")) for i in range(world)]
    for i in range(1, world):
        assert res[i]["sorted"] == res[0]["sorted"]
    for c in range(3):
        allk = sum((res[i]["local"][c] for i in range(world)), [])
        vc = pd.Series(allk).value_counts(sort=False)
        exp = pd.DataFrame({"k": vc.index.to_numpy(), "s": vc.to_numpy()}).sort_values(
            ["s", "k"], ascending=[False, True], kind="stable")
        assert res[0]["sorted"][c]["keys"] == exp["k"].tolist()
        assert res[0]["sorted"][c]["sizes"] == exp["s"].tolist()
        assert res[0]["sorted"][c]["null"] == sum(2 + i + c for i in range(world))
