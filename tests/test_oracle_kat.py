"""Pins the CPU oracle with the reference's own known-answer tests (SURVEY.md
§8c).  Each test cites the reference test it re-types.  CPU only."""
import math

import numpy as np
import pandas as pd
import pytest

import oracle
from oracle.categorify import CategorifyOracle
from oracle.groupby import groupby_stats, join_groupby_transform, target_encoding


def _fit_transform(df, groups, **kw):
    o = CategorifyOracle(groups, **kw).fit(df)
    return o, o.transform(df)


# reference tests/unit/ops/test_categorify.py:124-157
@pytest.mark.parametrize("freq_threshold", [0, 1, 2])
@pytest.mark.parametrize("dtype", [None, np.int32, np.int64])
@pytest.mark.parametrize("use_vocab", [False, True])
def test_categorify_lists(freq_threshold, dtype, use_vocab):
    df = pd.DataFrame({
        "Authors": [["User_A"], ["User_A", "User_E"], ["User_B", "User_C"], ["User_C"]],
        "Engaging User": ["User_B", "User_B", "User_A", "User_D"],
        "Post": [1, 2, 3, 4],
    })
    vocabs = {"Authors": pd.Series([f"User_{x}" for x in "ACBE"])} if use_vocab else None
    _, out = _fit_transform(df, ["Authors", "Engaging User"], freq_threshold=freq_threshold,
                            dtype=dtype, vocabs=vocabs)
    assert out["Authors"][0].dtype == (np.dtype(dtype) if dtype else np.dtype("int64"))
    compare = [list(r) for r in out["Authors"].tolist()]
    if freq_threshold < 2 or use_vocab:
        assert compare == [[3], [3, 6], [5, 4], [4]]
    else:
        assert compare == [[3], [3, 2], [2, 4], [4]]


# reference tests/unit/ops/test_categorify.py:160-216
@pytest.mark.parametrize("cat_names", [[["Author", "Engaging User"]], ["Author", "Engaging User"]])
@pytest.mark.parametrize("kind", ["joint", "combo"])
def test_categorify_multi(cat_names, kind):
    df = pd.DataFrame({
        "Author": ["User_A", "User_E", "User_B", "User_C"],
        "Engaging User": ["User_B", "User_B", "User_A", "User_D"],
        "Post": [1, 2, 3, 4],
    })
    _, out = _fit_transform(df, cat_names, encode_type=kind)
    if len(cat_names) == 1:
        if kind == "joint":
            assert out["Author"].tolist() == [4, 7, 3, 5]
            assert out["Engaging User"].tolist() == [3, 3, 4, 6]
        else:
            assert out["Author_Engaging User"].tolist() == [3, 6, 4, 5]
    else:
        assert out["Author"].tolist() == [3, 6, 4, 5]
        assert out["Engaging User"].tolist() == [3, 3, 4, 5]


_COMBO_CASES = [
    ({"Author": ["User_B", "User_E", "User_B", "User_C"], "Engaging User": ["User_C", "User_B", "User_A", "User_D"]},
     [3, 5, 3, 4], [5, 4, 3, 6], [4, 6, 3, 5]),
    ({"Author": ["User_A", "User_E", "User_B", "User_C"], "Engaging User": ["User_B", "User_B", "User_A", "User_D"]},
     [3, 6, 4, 5], [3, 3, 4, 5], [3, 6, 4, 5]),
    ({"Author": ["User_C", "User_E", "User_B", "User_C"], "Engaging User": ["User_B", "User_B", "User_A", "User_D"]},
     [3, 5, 4, 3], [3, 3, 4, 5], [4, 6, 3, 5]),
    ({"Author": ["User_A", "User_B", "User_C", "User_C"], "Engaging User": ["User_A", "User_B", "User_C", "User_C"]},
     [4, 5, 3, 3], [4, 5, 3, 3], [4, 5, 3, 3]),
    ({"Author": ["User_C", "User_E", "User_B", "User_A"], "Engaging User": ["User_C", "User_B", "User_A", "User_D"]},
     [5, 6, 4, 3], [5, 4, 3, 6], [5, 6, 4, 3]),
    ({"Author": [np.nan, "User_E", "User_B", "User_A"], "Engaging User": ["User_C", "User_B", "User_A", "User_D"]},
     [1, 5, 4, 3], [5, 4, 3, 6], [3, 6, 5, 4]),
]


# reference tests/unit/ops/test_categorify.py:219-323
@pytest.mark.parametrize("case", _COMBO_CASES)
def test_categorify_multi_combo(case):
    data, exp_a, exp_e, exp_ae = case
    df = pd.DataFrame({**data, "Post": [1, 2, 3, 4]})
    _, out = _fit_transform(df, [["Author", "Engaging User"], "Author", "Engaging User"], encode_type="combo")
    assert out["Author"].tolist() == exp_a
    assert out["Engaging User"].tolist() == exp_e
    assert out["Author_Engaging User"].tolist() == exp_ae


# reference tests/unit/ops/test_categorify.py:99-121
def test_na_value_count():
    df = pd.DataFrame({
        "productID": ["B00406YHLI"] * 5 + ["B002YXS8E6"] * 5 + ["B00011KM38"] * 2 + [np.nan] * 3,
        "brand": ["Coby"] * 5 + [np.nan] * 5 + ["Cooler Master"] * 2 + ["Asus"] * 3,
    })
    o, _ = _fit_transform(df, ["brand", "productID"])
    assert o.categories["brand"].meta["kind"].iloc[1] == "null"
    assert o.categories["brand"].meta["num_observed"].iloc[1] == 5
    assert o.categories["productID"].meta["num_observed"].iloc[1] == 3


# reference tests/unit/ops/test_categorify.py:326-421 (merge path: not search_sorted)
@pytest.mark.parametrize("freq_limit", [{"Author": 3, "Engaging User": 4}])
@pytest.mark.parametrize("buckets", [None, 10, {"Author": 10, "Engaging User": 20}])
def test_categorify_freq_limit(freq_limit, buckets):
    df = pd.DataFrame({
        "Author": ["User_A", "User_E", "User_B", "User_C", "User_A", "User_E", "User_B", "User_C", "User_B", "User_C"],
        "Engaging User": ["User_B", "User_B", "User_A", "User_D", "User_B", "User_c", "User_A", "User_D", "User_D", "User_D"],
    })
    # string OOV hashing is outside the numeric oracle hash: use integer ids
    ids = {s: i for i, s in enumerate(sorted(set(df["Author"]) | set(df["Engaging User"])))}
    dfi = df.replace(ids).astype("int64")
    o, out = _fit_transform(dfi, ["Author", "Engaging User"], freq_threshold=freq_limit, num_buckets=buckets)
    for col in ["Author", "Engaging User"]:
        assert o.categories[col].meta["num_observed"].sum() == len(df)
    freq_limited = {"Author": 2, "Engaging User": 1}
    if not buckets:
        assert out["Author"].max() == 1 + 1 + freq_limited["Author"]
        assert out["Engaging User"].max() == 1 + 1 + freq_limited["Engaging User"]


# reference tests/unit/ops/test_categorify.py:424-447
def test_categorify_hash_bucket_only():
    df = pd.DataFrame({"Authors": [0, 0, 4, 1, 2], "Engaging_User": [1, 1, 0, 3, 3], "Post": [1, 2, 3, 4, 5]})
    buckets = 10
    max_size = buckets + 2
    o, out = _fit_transform(df, ["Authors", "Engaging_User"], num_buckets=buckets, max_size=max_size)
    assert out["Authors"].max() <= max_size
    assert out["Engaging_User"].max() <= max_size
    assert o.embedding_sizes()["Authors"][0] == max_size


# reference tests/unit/ops/test_categorify.py:450-506
@pytest.mark.parametrize("max_emb_size", [6, {"Author": 8, "Engaging_User": 7}])
def test_categorify_max_size(max_emb_size):
    a = "A E B C A E B C D F F".split()
    e = "B B A D B M A D N F E".split()
    ids = {s: i for i, s in enumerate(sorted(set(a) | set(e)))}
    df = pd.DataFrame({"Author": [ids[x] for x in a], "Engaging_User": [ids[x] for x in e]})
    o, out = _fit_transform(df, ["Author", "Engaging_User"], max_size=max_emb_size, num_buckets=3)
    if isinstance(max_emb_size, int):
        max_emb_size = {n: max_emb_size for n in ["Author", "Engaging_User"]}
    for n in ["Author", "Engaging_User"]:
        assert out[n].max() <= max_emb_size[n] + 1
        assert o.embedding_sizes()[n][0] <= max_emb_size[n] + 1


# reference tests/unit/ops/test_categorify.py:509-529
def test_categorify_single_table():
    df = pd.DataFrame({
        "Authors": [None, "User_A", "User_A", "User_E", "User_B", "User_C"],
        "Engaging_User": [None, "User_B", "User_B", "User_A", "User_D", "User_D"],
    })
    _, out = _fit_transform(df, ["Authors", "Engaging_User"], single_table=True)
    old_max = 1
    for name in ["Authors", "Engaging_User"]:
        assert old_max <= out[name].min()
        old_max += out[name].max()


# reference tests/unit/ops/test_categorify.py:543-556 and :615-633
def test_categorify_null_meta():
    df = pd.DataFrame({"user_id": [1, 2, 3, 4, 6, 8, 5, 3] * 10, "item_id": [2, 4, 4, 7, 5, 2, 5, 2] * 10})
    o, _ = _fit_transform(df, ["user_id", "item_id"])
    assert o.categories["user_id"].meta["num_observed"].iloc[1] == 0
    df = pd.DataFrame({"C1": [1, np.nan, 3, 4, 3] * 5, "C2": [1, 1, 2, 3, 6] * 5})
    o, _ = _fit_transform(df, ["C1", "C2"], max_size=4)
    assert o.categories["C1"].meta["num_observed"].iloc[1] == 5
    assert o.categories["C2"].meta["num_observed"].iloc[1] == 0


# reference tests/unit/ops/test_categorify.py:636-665
def test_categorify_joint_list():
    df = pd.DataFrame({
        "Author": ["User_A", "User_E", "User_B", "User_C"],
        "Engaging User": [["User_B", "User_C"], [], ["User_A", "User_D"], ["User_A"]],
        "Post": [1, 2, 3, 4],
    })
    _, out = _fit_transform(df, ["Post", ["Author", "Engaging User"]], encode_type="joint")
    assert out["Author"].tolist() == [3, 7, 4, 5]
    assert [x for r in out["Engaging User"] for x in r] == [4, 5, 3, 6, 3]


# reference tests/unit/ops/test_categorify.py:38-96
@pytest.mark.parametrize("include_nulls", [True, False])
def test_categorify_size(include_nulls):
    rng = np.random.RandomState(0)
    ids = list(range(10)) + ([None] if include_nulls else [])
    df = pd.DataFrame({"session_id": [ids[i] for i in rng.randint(0, len(ids), 50)]})
    o, _ = _fit_transform(df, ["session_id"])
    vals = df["session_id"].value_counts()
    v = o.categories["session_id"].unique
    assert dict(zip(v["session_id"], v["session_id_size"])) == dict(zip(vals.index, vals))


def test_partitions_equal_single():
    """tree reduction over partitions == single pass (reference test_categorify.py:668-704 spirit)."""
    rng = np.random.RandomState(1)
    df = pd.DataFrame({"a": rng.randint(0, 50, 1000), "b": rng.randint(0, 7, 1000)})
    o1, t1 = _fit_transform(df, ["a", "b"])
    parts = [df.iloc[i:i + 100] for i in range(0, 1000, 100)]
    o2 = CategorifyOracle(["a", "b"]).fit(parts)
    for c in ["a", "b"]:
        pd.testing.assert_frame_equal(o1.categories[c].unique, o2.categories[c].unique)
    pd.testing.assert_frame_equal(t1, o2.transform(df))


# reference tests/unit/ops/test_normalize.py:60-84, 110-139
def test_normalize():
    rng = np.random.RandomState(2)
    df = pd.DataFrame({"x": rng.rand(1000), "y": rng.randint(-5, 100, 1000)})
    means, stds = oracle.normalize_fit([df.iloc[:300], df.iloc[300:]], ["x", "y"])
    for c in ["x", "y"]:
        assert math.isclose(df[c].mean(), means[c], rel_tol=1e-4)
        assert math.isclose(df[c].std(), stds[c], rel_tol=1e-4)
    out = oracle.normalize_transform(df, ["x", "y"], means, stds)
    assert np.all(((df["x"] - df["x"].mean()) / df["x"].std() - out["x"]).abs() <= 1e-2)
    df0 = pd.DataFrame({"a": 7 * [10]})
    m, s = oracle.normalize_fit(df0, ["a"])
    assert (oracle.normalize_transform(df0, ["a"], m, s)["a"] == 0).all()
    big = pd.DataFrame({"x": [1.9e10, 2.3e16, 3.4e18, 1.6e19]})
    m, s = oracle.normalize_fit(big, ["x"])
    assert math.isclose(big["x"].mean(), m["x"], rel_tol=1e-4)
    assert math.isclose(big["x"].std(), s["x"], rel_tol=1e-4)


# reference tests/unit/ops/test_fill.py:61-85
@pytest.mark.parametrize("add_binary_cols", [True, False])
def test_fill_missing(add_binary_cols):
    rng = np.random.RandomState(3)
    df = pd.DataFrame({"x": rng.rand(100), "y": rng.rand(100)})
    df.loc[rng.choice(100, 20), "x"] = None
    out = oracle.fill_missing(df, ["x", "y"], 42, add_binary_cols)
    assert out["x"].isna().sum() == 0
    assert np.all((df["x"].fillna(42) - out["x"]).abs() <= 1e-2)
    assert ("x_filled" in out) == add_binary_cols
    if add_binary_cols:
        assert df["x"].isna().sum() == out["x_filled"].sum()


# reference tests/unit/ops/test_join.py:32-92
def test_joingroupby():
    df = pd.DataFrame({"Author": ["User_A", "User_A", "User_A", "User_B"],
                       "Engaging-User": ["User_B", "User_B", "User_C", "User_C"],
                       "Cost": [100.0, 200.0, 300.0, 400.0]})
    t = groupby_stats(df, ["Author", "Engaging-User"], ["Cost"], ["sum"])
    out = join_groupby_transform(df, [["Author", "Engaging-User"]], {"Author_Engaging-User": t})
    assert out["Author_Engaging-User_Cost_sum"].tolist() == [300.0, 300.0, 300.0, 400.0]
    t = groupby_stats(df, ["Author"], ["Cost"], ["sum"])
    out = join_groupby_transform(df, ["Author"], {"Author": t})
    assert out["Author_Cost_sum"].tolist() == [600.0, 600.0, 600.0, 400.0]
    df = pd.DataFrame({"Author": ["User_A"] * 3 + ["User_B"] * 2, "Cost": [0.0, 1 / 3, 2 / 3, 1.0, 1.0]})
    t = groupby_stats(df, ["Author"], ["Cost"], ["sum"])
    out = join_groupby_transform(df, ["Author"], {"Author": t})
    assert out["Author_Cost_sum"].tolist() == [1.0, 1.0, 1.0, 2.0, 2.0]


# reference tests/unit/ops/test_target_encode.py:111-147
def test_target_encode_multi():
    cat_1 = np.asarray(["baaaa"] * 12)
    cat_2 = np.asarray(["baaaa"] * 6 + ["bbaaa"] * 3 + ["bcaaa"] * 3)
    num_1 = np.asarray([1, 1, 2, 2, 2, 1, 1, 5, 4, 4, 4, 4])
    df = pd.DataFrame({"cat": cat_1, "cat2": cat_2, "num": num_1, "num_2": num_1 * 2})
    for parts in ([df], [df.iloc[:6], df.iloc[6:]]):
        outs, _, _ = target_encoding(parts, ["cat", "cat2", ["cat", "cat2"]], ["num", "num_2"],
                                     kfold=1, p_smooth=5, out_dtype="float32")
        out = pd.concat(outs)
        np.testing.assert_array_equal(out["TE_cat2_num"].values, out["TE_cat_cat2_num"].values)
        assert out["TE_cat_num"].iloc[0] != out["TE_cat2_num"].iloc[0]
        assert math.isclose(out["TE_cat_num"].iloc[0], num_1.mean(), abs_tol=1e-4)
        assert math.isclose(out["TE_cat_num_2"].iloc[0], (num_1 * 2).mean(), abs_tol=1e-3)


# SURVEY.md §8c golden vectors of pandas.util.hash_array(index=False)
def test_hash_golden_inline():
    h = oracle.hash_values(np.array([0, 1, 2, -1, 2**31, 2**40], dtype="int64"))
    assert [hex(int(x)) for x in h] == ["0x0", "0x5692161d100b05e5", "0xdbd238973a2b148a",
                                        "0xb4d055fcf2cbbd7b", "0xec105bf588587c9f", "0xab4daf7c2673f8"]
    assert hex(int(oracle.hash_values(np.array([-1], dtype="int32"))[0])) == "0x8b32c408e8c2c97c"
    assert oracle.hash_bucket(np.array([1, 2, 3]), 10).tolist() == [9, 0, 6]
    for dt in ["int32", "int64", "float32", "float64"]:
        arr = np.arange(-50, 50).astype(dt)
        np.testing.assert_array_equal(oracle.hash_values(arr), pd.util.hash_array(arr))


def test_hash_golden_file():
    """oracle hash == golden vectors generated from pandas.util.hash_array
    (tests/golden/make_hash_golden.py), for every numeric dtype incl. specials."""
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "hash_golden.json")
    g = json.load(open(path))
    assert g["mod10_of_1_2_3_int64"] == [9, 0, 6]
    for case in g["cases"]:
        dt = case["dtype"]
        bits = np.array(case["bits"], dtype=np.uint64)
        if dt == "bool":
            arr = bits.astype(bool)
        else:
            width = np.dtype(dt).itemsize
            arr = bits.astype(f"u{width}").view(dt)
        got = oracle.hash_values(arr) if dt not in ("float32", "float64") else \
            oracle.hashing._mix(bits)   # raw-bit hash (the oracle maps NaN to the null pattern)
        np.testing.assert_array_equal(got, np.array(case["hash"], dtype=np.uint64), err_msg=dt)


# reference tests/unit/ops/test_categorify.py:543-557 (issue 1325: the null row exists with 0 observations)
def test_categorify_no_nulls():
    df = pd.DataFrame({"user_id": [1, 2, 3, 4, 6, 8, 5, 3] * 10, "item_id": [2, 4, 4, 7, 5, 2, 5, 2] * 10})
    o, out = _fit_transform(df, ["user_id", "item_id"])
    meta = o.categories["user_id"].meta
    assert meta["kind"].iloc[1] == "null" and meta["num_observed"].iloc[1] == 0
    # pad, null and oov rows precede the uniques: the most frequent user (3) gets the first free index
    assert out["user_id"].min() == 3 and set(out.loc[df["user_id"] == 3, "user_id"]) == {3}


# reference tests/unit/ops/test_categorify.py:615-634
def test_categorify_max_size_null_iloc_check():
    df = pd.DataFrame({"C1": [1, np.nan, 3, 4, 3] * 5, "C2": [1, 1, 2, 3, 6] * 5})
    o, _ = _fit_transform(df, ["C1", "C2"], max_size=4)
    m1, m2 = o.categories["C1"].meta, o.categories["C2"].meta
    assert m1["kind"].iloc[1] == "null" and m1["num_observed"].iloc[1] == 5
    assert m2["kind"].iloc[1] == "null" and m2["num_observed"].iloc[1] == 0


# reference tests/unit/ops/test_target_encode.py:37-84: with unique categories the per-fold stat
# table holds exactly the (fold, key) pairs of the transformed rows
@pytest.mark.parametrize("cat_groups", ["Author", [["Author", "Engaging-User"]]])
@pytest.mark.parametrize("kfold", [1, 3])
@pytest.mark.parametrize("fold_seed", [None, 42])
def test_target_encode_fold_mapping(cat_groups, kfold, fold_seed):
    import string
    from oracle.groupby import add_fold
    df = pd.DataFrame({"Author": list(string.ascii_uppercase), "Engaging-User": list(string.ascii_lowercase),
                       "Cost": range(26), "Post": [0, 1] * 13})
    parts = [df.iloc[:9], df.iloc[9:18], df.iloc[18:]]
    groups = [cat_groups] if isinstance(cat_groups, str) else cat_groups
    outs, tables, y_mean = target_encoding(parts, groups, ["Cost"], kfold=kfold, fold_seed=fold_seed,
                                           out_dtype="float32")
    out = pd.concat(outs)
    assert len(out) == 26 and out.dtypes.iloc[0] == np.float32
    assert math.isclose(y_mean["Cost"], 12.5)
    if kfold > 1:
        g = [cat_groups] if isinstance(cat_groups, str) else cat_groups[0]
        cols = ["__fold__"] + g
        check = tables["_".join(cols)][cols].sort_values(cols).reset_index(drop=True)
        folded = pd.concat([p.assign(__fold__=add_fold(len(p), kfold, fold_seed)) for p in parts])
        got = folded[cols].sort_values(cols).reset_index(drop=True)
        pd.testing.assert_frame_equal(check, got, check_dtype=False)
        # every key is alone in its group: the out-of-fold estimate is the smoothed prior
        np.testing.assert_allclose(out.iloc[:, 0].values, 12.5, rtol=1e-6)
    else:
        te = out.iloc[:, 0].values
        np.testing.assert_allclose(te, (df["Cost"].values + 20 * 12.5) / 21.0, rtol=1e-6)


# reference tests/unit/ops/test_target_encode.py:87-110 (kfold=1; values from the op's formula
# target_encoding.py:376-380: (sum + p_smooth * mean) / (count + p_smooth))
def test_target_encode_group():
    df = pd.DataFrame({"Cost": range(15), "Post": [1, 2, 3, 4, 5] * 3,
                       "Author": ["A"] * 5 + ["B"] * 5 + ["C"] * 2 + ["D"] * 3,
                       "Engaging_User": ["A"] * 5 + ["B"] * 3 + ["E"] * 2 + ["D"] * 3 + ["G"] * 2})
    df["label"] = (df["Post"] > 3).astype("int8")
    outs, _, y_mean = target_encoding([df], ["Author", "Engaging_User"], ["label"], kfold=1, out_dtype="float32")
    out = outs[0]
    assert math.isclose(y_mean["label"], 0.4)
    exp = {"A": (2 + 8.0) / 25, "B": (2 + 8.0) / 25, "C": (0 + 8.0) / 22, "D": (2 + 8.0) / 23}
    np.testing.assert_allclose(out["TE_Author_label"].values, df["Author"].map(exp).values, rtol=1e-6)
    exp_u = {"A": 10.0 / 25, "B": (0 + 8.0) / 23, "E": (2 + 8.0) / 22, "D": (0 + 8.0) / 23, "G": (2 + 8.0) / 22}
    np.testing.assert_allclose(out["TE_Engaging_User_label"].values, df["Engaging_User"].map(exp_u).values, rtol=1e-6)


# reference tests/unit/ops/test_hash_bucket.py:36-57: buckets in range, deterministic
def test_hash_bucket_range_and_determinism():
    rng = np.random.RandomState(5)
    v = rng.randint(-2**62, 2**62, 10000)
    b = oracle.hash_bucket(v, 10)
    assert b.min() >= 0 and b.max() <= 9 and len(np.unique(b)) == 10
    np.testing.assert_array_equal(b, oracle.hash_bucket(v.copy(), 10))
    np.testing.assert_array_equal(b, (pd.util.hash_array(v) % np.uint64(10)).astype(b.dtype))
