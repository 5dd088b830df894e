"""GPU tests of the drop-in boundary beyond fit/transform:
Workflow.save / Workflow.load in the reference's graph.json + artifacts/node_<id>/ layout
(nvtabular/workflow/workflow.py:256-348, graph_serializer.py:1077-1165), vocabulary files written
the way the reference writes them (categorify.py:731-822) loading here, Clip / LogOp
(ops/clip.py:46-53, ops/logop.py:47-56) and the inference hooks (categorify.py:589-609)."""
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nvt():
    import nvtabular
    return nvtabular


def _frame(n=20000, seed=3):
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({
        "a": rng.integers(0, 500, n).astype(np.int32), "b": rng.integers(0, 40, n).astype(np.int64),
        "s": rng.choice(["x", "yy", "zzz", "w"], n), "x": rng.normal(2, 3, n), "y": rng.integers(-5, 50, n).astype(np.int32),
        "t": rng.integers(0, 2, n).astype(np.float32)})
    df.loc[rng.random(n) < 0.05, "x"] = np.nan
    return df


def _workflow(nvt, path):
    ops = nvt.ops
    cat = ["a", "b", "s"] >> ops.Categorify(out_path=path, freq_threshold=2)
    combo = [["a", "b"]] >> ops.Categorify(out_path=path + "_combo", encode_type="combo")
    cont = ["x"] >> ops.FillMissing() >> ops.Normalize()
    logc = ["y"] >> ops.Clip(min_value=0) >> ops.LogOp()
    jg = ["a", ["a", "b"]] >> ops.JoinGroupby(out_path=path, cont_cols=["x"], stats=["count", "sum", "mean", "std"])
    te = ["a", ["a", "s"]] >> ops.TargetEncoding("t", kfold=3, p_smooth=10, out_path=path)
    hb = ["b"] >> ops.HashBucket(17)
    out = cat + combo + cont + logc + jg + te
    return nvt.Workflow(out), hb


def test_save_load_roundtrip_in_process_and_fresh_process(nvt, tmp_path):
    df = _frame()
    wf, _ = _workflow(nvt, str(tmp_path / "fit"))
    exp = wf.fit_transform(nvt.Dataset(df, npartitions=2)).to_ddf().compute()
    save_dir = str(tmp_path / "saved")
    wf.save(save_dir)
    # layout of the reference (graph_serializer.py:16-29)
    assert os.path.exists(os.path.join(save_dir, "metadata.json"))
    graph = json.load(open(os.path.join(save_dir, "graph.json")))
    assert graph["format_version"] == 1 and {"id", "op_class", "op_params", "op_state", "parent_ids", "dependency_ids",
                                             "selector", "input_schema", "output_schema"} <= set(graph["nodes"][0])
    classes = {n["op_class"] for n in graph["nodes"]}
    assert "nvtabular.ops.categorify.Categorify" in classes and "merlin.dag.ops.selection.SelectionOp" in classes
    cat_node = [n for n in graph["nodes"] if n["op_class"].endswith("Categorify")][0]
    rel = cat_node["op_state"]["categories"][0]["path"]
    assert os.path.exists(os.path.join(save_dir, "artifacts", f"node_{cat_node['id']}", rel))
    # saving must not re-point the live op at the save directory (a later re-fit would overwrite it)
    live = [n.op for n in wf.output_node.topo_order() if type(n.op).__name__ == "Categorify"][0]
    assert str(tmp_path / "fit") in live.categories["a"] and "saved" not in live.out_path
    # in-process load: bit-identical transform, embedding sizes from the files
    wf2 = nvt.Workflow.load(save_dir)
    got = wf2.transform(nvt.Dataset(df, npartitions=2)).to_ddf().compute()
    assert list(got.columns) == list(exp.columns)
    for c in exp.columns:
        assert got[c].dtype == exp[c].dtype, c
        np.testing.assert_array_equal(got[c].to_numpy(), exp[c].to_numpy(), err_msg=c)
    assert nvt.ops.get_embedding_sizes(wf2) == nvt.ops.get_embedding_sizes(wf)
    # fresh process
    df.to_parquet(tmp_path / "in.parquet")
    exp.to_parquet(tmp_path / "exp.parquet")
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); import pandas as pd, numpy as np, nvtabular as nvt\n"
            f"wf = nvt.Workflow.load({save_dir!r})\n"
            f"df = pd.read_parquet({str(tmp_path / 'in.parquet')!r}); exp = pd.read_parquet({str(tmp_path / 'exp.parquet')!r})\n"
            "got = wf.transform(nvt.Dataset(df, npartitions=2)).to_ddf().compute()\n"   # TE folds are drawn per partition

            "assert list(got.columns) == list(exp.columns)\n"
            "for c in exp.columns: np.testing.assert_array_equal(got[c].to_numpy(), exp[c].to_numpy(), err_msg=c)\n"
            "print('RELOAD_OK')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "RELOAD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_reference_format_vocabulary_files_load_here(nvt, tmp_path):
    """unique.<col>.parquet / meta.<col>.parquet written with pandas exactly as the reference's
    _save_encodings does (categorify.py:731-822: RangeIndex starting at the first label, a
    `<col>_size` column) drive a transform here."""
    base = tmp_path / "categories"
    os.makedirs(base)
    uniq = pd.DataFrame({"brand": ["acme", "zeta", "beta"], "brand_size": [5, 3, 1]})
    uniq.index = pd.RangeIndex(3, 6)
    uniq.to_parquet(base / "unique.brand.parquet")
    pd.DataFrame({"kind": ["pad", "null", "oov", "unique"], "offset": [0, 1, 2, 3], "num_indices": [1, 1, 1, 3],
                  "num_observed": [0, 1, 0, 9]}).to_parquet(base / "meta.brand.parquet")
    ints = pd.DataFrame({"item": np.array([40, 7, 19], dtype=np.int64), "item_size": [9, 4, 4]})
    ints.index = pd.RangeIndex(3, 6)
    ints.to_parquet(base / "unique.item.parquet")
    op = nvt.ops.Categorify(vocabs={"brand": str(base / "unique.brand.parquet"), "item": str(base / "unique.item.parquet")},
                            out_path=str(tmp_path))
    wf = nvt.Workflow(["brand", "item"] >> op)
    df = pd.DataFrame({"brand": ["zeta", "acme", None, "nope", "beta"], "item": [19, 40, 7, 8, 19]})
    out = wf.fit_transform(nvt.Dataset(df)).to_ddf().compute()
    assert out["brand"].tolist() == [4, 3, 1, 2, 5]
    assert out["item"].tolist() == [5, 3, 4, 2, 5]


def test_clip_logop_vs_numpy(nvt):
    rng = np.random.default_rng(5)
    n = 100_003
    df = pd.DataFrame({"i": rng.integers(-10, 1000, n).astype(np.int32), "f": rng.normal(3, 20, n).astype(np.float32),
                       "d": rng.normal(3, 20, n), "k": rng.integers(-3, 9, n).astype(np.int64)})
    df.loc[rng.random(n) < 0.1, "d"] = np.nan
    ops = nvt.ops
    clipped = ["i", "f", "d", "k"] >> ops.Clip(min_value=0, max_value=500)
    out = nvt.Workflow(clipped).fit_transform(nvt.Dataset(df)).to_ddf().compute()
    for c in df.columns:
        exp = df[c].copy()
        exp[exp < 0] = 0
        exp[exp > 500] = 500
        assert out[c].dtype == df[c].dtype, c
        np.testing.assert_array_equal(out[c].to_numpy(), exp.to_numpy(), err_msg=c)
    # the published Criteo continuous pipeline: FillMissing >> Clip(min_value=0) >> LogOp (benchmark.py:201-204)
    logged = ["i", "d"] >> ops.FillMissing() >> ops.Clip(min_value=0) >> ops.LogOp()
    out = nvt.Workflow(logged).fit_transform(nvt.Dataset(df)).to_ddf().compute()
    for c in ["i", "d"]:
        x = df[c].fillna(0).to_numpy()
        x = np.where(x < 0, 0, x).astype(df[c].dtype)
        exp = np.log(x.astype(np.float32) + 1)
        assert out[c].dtype == np.float32
        # float32 log: numpy's vectorised logf is within 1 ulp of the correctly rounded value computed here
        np.testing.assert_allclose(out[c].to_numpy(), exp, rtol=2.5e-7, atol=0, err_msg=c)
    with pytest.raises(ValueError):
        ops.Clip()


def test_inference_hooks(nvt, tmp_path):
    """inference_initialize / supported_formats / compute_selector / dict-of-arrays transform
    (categorify.py:589-609, normalize.py:92-108, fill.py:59-65)."""
    df = _frame(5000)
    ops = nvt.ops
    cat_op = ops.Categorify(out_path=str(tmp_path))
    norm_op = ops.Normalize()
    fill_op = ops.FillMissing(fill_val=-1)
    wf = nvt.Workflow((["a", "b"] >> cat_op) + (["x"] >> fill_op >> norm_op))
    exp = wf.fit_transform(nvt.Dataset(df)).to_ddf().compute()
    sel = nvt.ColumnSelector(["a", "b"])
    inf = cat_op.inference_initialize(sel, {})
    assert inf is not None and hasattr(inf, "transform")
    arrays = {"a": df["a"].to_numpy(), "b": df["b"].to_numpy()}
    got = inf.transform(sel, arrays)
    np.testing.assert_array_equal(np.asarray(got["a"]), exp["a"].to_numpy())
    np.testing.assert_array_equal(np.asarray(got["b"]), exp["b"].to_numpy())
    xs = nvt.ColumnSelector(["x"])
    filled = fill_op.inference_initialize(xs, {}).transform(xs, {"x": df["x"].to_numpy()})
    got = norm_op.transform(xs, {"x": filled["x"]})
    np.testing.assert_allclose(np.asarray(got["x"].cpu() if hasattr(got["x"], "cpu") else got["x"]), exp["x"].to_numpy(),
                               rtol=1e-12)
    assert norm_op.supported_formats is not None and cat_op.supported_formats is not None


def test_parquet_ingest_transform_egress_with_shuffle(nvt, tmp_path):
    """parquet -> device partitions -> Workflow -> to_parquet: the row-group
    reader keeps nullable int32 as int32 + bitmask; PER_PARTITION / PER_WORKER shuffles write
    a permutation of exactly the unshuffled rows (reference bench/examples/MultiGPUBench.md:75-89)."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    rng = np.random.default_rng(8)
    n = 30_000
    a = pd.array(rng.integers(0, 300, n), dtype="Int32")
    a[rng.random(n) < 0.1] = pd.NA
    df = pd.DataFrame({"a": a, "x": rng.normal(0, 1, n), "rid": np.arange(n, dtype=np.int64)})
    src = tmp_path / "in.parquet"
    pq.write_table(pa.Table.from_pandas(df, preserve_index=False), src, row_group_size=7000)
    ops = nvt.ops
    wf = nvt.Workflow((["a"] >> ops.Categorify(out_path=str(tmp_path / "c"))) + (["x"] >> ops.Normalize()) + ["rid"])
    ds = nvt.Dataset(str(src))
    assert ds.npartitions == 5                                   # one partition per row group
    wf.fit(ds)
    base = wf.transform(ds).to_ddf().compute()
    assert base["rid"].tolist() == list(range(n))
    exp = wf.transform(nvt.Dataset(df)).to_ddf().compute()      # the pandas path sees the same rows
    np.testing.assert_array_equal(base["a"].to_numpy(), exp["a"].to_numpy())
    for mode, kw in [(None, {}), ("PER_PARTITION", {}), (nvt.Shuffle.PER_WORKER, {"out_files_per_proc": 3})]:
        out_dir = tmp_path / f"out_{mode}"
        wf.transform(ds).to_parquet(str(out_dir), shuffle=mode, **kw)
        files = sorted(os.listdir(out_dir))
        assert len(files) == (3 if mode == "PER_WORKER" else 5)
        got = pd.concat([pd.read_parquet(out_dir / f) for f in files], ignore_index=True)
        assert len(got) == n
        if mode is None:
            assert got["rid"].tolist() == list(range(n))
        else:
            assert got["rid"].tolist() != list(range(n))
        got = got.sort_values("rid", ignore_index=True)
        np.testing.assert_array_equal(got["a"].to_numpy(), base["a"].to_numpy())
        np.testing.assert_array_equal(got["x"].to_numpy(), base["x"].to_numpy())
