"""Oracle restatement of Categorify — TEST INFRASTRUCTURE, never imported by
the product (see oracle/__init__.py).

Follows the reference nvtabular/ops/categorify.py on plain pandas:
  _top_level_groupby     :955-1051   groupby(dropna=False) per column group
  _mid_level_groupby     :1054-1070  concat partials + groupby again
  _bottom_level_groupby  :1073-1137  final merge (+ derived stats, see groupby.py)
  _write_uniques         :1149-1337  null peel, ordering
  _save_encodings        :719-822    threshold / max_size cut, label index, meta
  _encode                :1558-1807  left merge on key + sort by order
  _hash_bucket           :1837-1852
  _emb_sz_rule           :687-688
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

import numpy as np
import pandas as pd

from .hashing import hash_values

PAD_OFFSET, NULL_OFFSET, OOV_OFFSET = 0, 1, 2  # categorify.py:51-55


def _make_name(*args, sep="_"):  # categorify.py:691-692
    return sep.join(args)


def emb_sz_rule(n_cat: int, minimum_size=16, maximum_size=512):  # categorify.py:687-688
    return n_cat, min(max(minimum_size, round(1.6 * n_cat**0.56)), maximum_size)


def _is_list_series(ser: pd.Series) -> bool:
    if ser.dtype != object:
        return False
    for v in ser:
        if v is None or (isinstance(v, float) and np.isnan(v)):
            continue
        return isinstance(v, (list, tuple, np.ndarray))
    return False


def _flatten(ser: pd.Series) -> pd.Series:
    """dispatch.flatten_list_column_values: leaf values of a list column."""
    vals = [x for row in ser for x in (row if row is not None else [])]
    return pd.Series(vals)


def _maybe_flatten(ser: pd.Series) -> pd.Series:
    return _flatten(ser) if _is_list_series(ser) else ser


# --------------------------------------------------------------------------
# fit
# --------------------------------------------------------------------------
def top_level_groupby(df: pd.DataFrame, group: List[str], concat_groups: bool, name_sep="_"):
    """categorify.py:955-1051 for ONE column group with agg_list=["size"]."""
    sel_str = _make_name(*group, sep=name_sep)
    if concat_groups and len(group) > 1:
        # categorify.py:972-981: joint encoding concatenates the columns' values
        df_gb = pd.DataFrame({sel_str: pd.concat(
            [_maybe_flatten(df[c]) for c in group], ignore_index=True)})
        names = [sel_str]
    else:
        df_gb = df[group].copy(deep=False)
        names = list(group)
        if len(names) == 1:
            df_gb = pd.DataFrame({names[0]: _maybe_flatten(df_gb[names[0]])})
    # categorify.py:1018  (the "size" agg is attached to the first key column)
    gb = df_gb.groupby(names, dropna=False).agg({names[0]: ["size"]})
    gb.columns = [_make_name(*(tuple(names) + ("size",)), sep=name_sep)]
    gb.reset_index(inplace=True, drop=False)
    return gb, names


def mid_level_groupby(dfs: List[pd.DataFrame], names: List[str]):
    """categorify.py:1054-1070: concat partial frames, groupby again, sum sizes."""
    df = pd.concat(dfs, ignore_index=True)
    gb = df.groupby(names, dropna=False).agg(
        {c: "sum" for c in df.columns if c not in names})
    gb.reset_index(drop=False, inplace=True)
    return gb


@dataclass
class Vocab:
    """unique.<name>.parquet + meta.<name>.parquet of the reference."""
    name: str
    names: List[str]                 # key column names inside `unique`
    unique: pd.DataFrame             # index = label; columns = keys + <name>_size
    meta: pd.DataFrame
    num_buckets: Optional[int] = None


def write_uniques(gb: pd.DataFrame, names: List[str], name_sep="_", freq_threshold=0,
                  max_size=0, num_buckets=None, has_size=True) -> Vocab:
    """categorify.py:1149-1337 (+ _save_encodings :719-822), single partition."""
    field_name = _make_name(*names, sep=name_sep)
    oov_count = num_buckets or 1                                   # :1194
    if max_size and max_size < oov_count + 2:                      # :1206-1211
        raise ValueError(
            "`max_size` can never be less than the maximum of `num_buckets + 2` and `3`, "
            "because we must always reserve pad, null and at least 1 oov-bucket index.")
    name_size = "_".join(names + ["size"])
    null_size = None
    df = gb
    if len(df):
        # :1300 make sure the first category is null
        df = df.sort_values(names, na_position="first", ignore_index=True)
        has_size = name_size in df
        has_nans = bool(df[names].iloc[0].transpose().isnull().all())  # :1305
        if has_nans:
            if has_size:
                null_size = df[name_size].iloc[0]
            df = df.iloc[1:]
        else:
            null_size = 0
        if has_size:
            # :1316 — deviation 2: kind="stable" => ties keep key-ascending order
            df = df.sort_values(name_size, ascending=False, ignore_index=True, kind="stable")
        df_write = df
    else:
        # :1318-1324 empty input: a single null row
        df_write = pd.DataFrame({c: pd.Series([None], dtype=gb[c].dtype if c in gb else object)
                                 for c in names})
    return save_encodings(df_write.copy(), field_name, names, first_n=max_size or None,
                          freq_threshold=freq_threshold or None, oov_count=oov_count,
                          null_size=null_size, num_buckets=num_buckets)


def save_encodings(df: pd.DataFrame, field_name: str, names: List[str], first_n=None,
                   freq_threshold=None, oov_count=1, null_size=None, num_buckets=None,
                   preserve_index=False) -> Vocab:
    """categorify.py:719-822 for a single in-memory partition."""
    record_size_meta = True
    oov_size = 0
    unique_count = 0
    unique_size = 0
    size = oov_count + OOV_OFFSET                                  # :754
    _df = df
    _len = len(_df)
    size_col = f"{field_name}_size"
    if _len:
        if size_col not in _df.columns:
            record_size_meta = False
        if record_size_meta:
            first_n_local = first_n - size if first_n is not None else _len   # :768-771
            if first_n or freq_threshold:
                removed = None
                if freq_threshold:
                    sizes = _df[size_col]
                    removed = df[(sizes < freq_threshold) & (sizes > 0)]      # :778
                    _df = _df[(sizes >= freq_threshold) | (sizes == 0)]       # :779
                if first_n and _len > first_n_local:
                    removed = _df.iloc[first_n_local:]                        # :781
                    _df = _df.iloc[:first_n_local]
                if removed is not None:
                    oov_size += removed[size_col].sum()
                    _len = len(_df)
            unique_size += _df[size_col].sum()                                # :788
        if not preserve_index:
            _df = _df.copy()
            _df.index = pd.RangeIndex(start=size, stop=size + _len, step=1)   # :795-803
        size += _len
        unique_count += _len
    else:
        _df = _df.iloc[:0]
    meta = {
        "kind": ["pad", "null", "oov", "unique"],
        "offset": [PAD_OFFSET, NULL_OFFSET, OOV_OFFSET, OOV_OFFSET + oov_count],
        "num_indices": [1, 1, oov_count, unique_count],
    }                                                                          # :812-816
    if record_size_meta:
        meta["num_observed"] = [0, null_size, oov_size, unique_size]          # :818
    return Vocab(field_name, names, _df, pd.DataFrame(meta), num_buckets)


def _resolve(opt, name):
    if isinstance(opt, dict):
        return opt.get(name)
    return opt


def categorify_fit(partitions: Union[pd.DataFrame, List[pd.DataFrame]], col_groups,
                   encode_type="joint", freq_threshold=0, max_size=0, num_buckets=None,
                   name_sep="_", split_every=8) -> Dict[str, Vocab]:
    """Categorify.fit -> _category_stats -> _groupby_to_disk (categorify.py:345-402,
    1344-1540) on a list of pandas partitions: per-partition groupby, tree
    reduction with fan-in `split_every`, then _write_uniques."""
    if isinstance(partitions, pd.DataFrame):
        partitions = [partitions]
    out = {}
    for group in col_groups:
        group = [group] if isinstance(group, str) else list(group)
        concat = encode_type == "joint"
        level = []
        names = None
        for df in partitions:
            gb, names = top_level_groupby(df, group, concat, name_sep)
            level.append(gb)
        while len(level) > 1:                                     # :1425-1469 tree
            level = [mid_level_groupby(level[i:i + split_every], names)
                     for i in range(0, len(level), split_every)]
        gb = mid_level_groupby(level, names)                      # _bottom_level_groupby
        field_name = _make_name(*names, sep=name_sep)
        out[field_name] = write_uniques(
            gb, names, name_sep,
            freq_threshold=_resolve(freq_threshold, field_name) or 0,
            max_size=_resolve(max_size, field_name) or 0,
            num_buckets=_resolve(num_buckets, field_name))
    return out


def vocab_from_series(col_name: str, vocab: pd.Series, num_buckets=None) -> Vocab:
    """Categorify.process_vocabs (categorify.py:421-454) for a user Series."""
    oov_count = num_buckets or 1
    col_df = pd.DataFrame({col_name: vocab}).dropna()
    col_df.index = col_df.index + NULL_OFFSET + oov_count
    return save_encodings(col_df, col_name, [col_name], oov_count=oov_count,
                          num_buckets=num_buckets)


# --------------------------------------------------------------------------
# transform
# --------------------------------------------------------------------------
def hash_bucket_oov(df, num_buckets, cols, encode_type="joint"):
    """categorify.py:1837-1852."""
    if encode_type == "joint":
        return hash_values(df[cols[0]]) % np.uint64(num_buckets)
    val = np.zeros(len(df), dtype=np.uint64)
    for c in cols:
        val ^= hash_values(df[c])
    return val % np.uint64(num_buckets)


def categorify_encode(df: pd.DataFrame, name, vocab: Vocab, num_buckets=None,
                      encode_type="joint", dtype=None, single_table=False):
    """_encode (categorify.py:1558-1807), merge path.  `name` is a column name or
    a list of names (combo).  Returns a numpy array (or a list-of-arrays Series
    for a list column)."""
    sel_l = list(name) if isinstance(name, (list, tuple)) else [name]
    sel_r = list(name) if isinstance(name, (list, tuple)) else vocab.names
    num_oov_buckets = num_buckets if num_buckets else 1           # :1615-1619
    value = vocab.unique[sel_r].copy()
    value.index = value.index.rename("labels")
    value.reset_index(drop=False, inplace=True)
    if len(value) == 0 and len(vocab.unique) == 0:
        value = pd.DataFrame({**{c: pd.Series([None], dtype=object) for c in sel_r},
                              "labels": [0]})[["labels"] + sel_r]
    null_off = value["labels"].head(1).iloc[0] if single_table else NULL_OFFSET   # :1683
    bucket_off = null_off + 1
    list_col = _is_list_series(df[sel_l[0]])
    if list_col and len(sel_l) != 1:
        raise ValueError("Can't categorical encode multiple list columns")     # :1823-1824
    expr = df[sel_l[0]].isna() if not list_col else pd.Series(False, index=df.index)
    for n in sel_l[1:]:
        expr = expr & df[n].isna()                                             # :1689-1691
    if list_col:
        flat = _flatten(df[sel_l[0]])
        codes = pd.DataFrame({"order": np.arange(len(flat)), sel_l[0]: flat})
        nulls = np.flatnonzero(flat.isna().to_numpy())
    else:
        codes = pd.DataFrame({"order": np.arange(len(df))}, index=df.index)
        for cl in sel_l:
            codes[cl] = df[cl].copy()
        nulls = np.flatnonzero(expr.to_numpy())
    for cl, cr in zip(sel_l, sel_r):
        if len(value) and value[cr].dtype != codes[cl].dtype:
            try:
                codes[cl] = codes[cl].astype(value[cr].dtype)                  # :1707
            except (TypeError, ValueError):
                pass
    indistinct = bucket_off
    if num_buckets:
        src = codes if list_col else df
        indistinct = hash_bucket_oov(src, num_buckets, sel_l, encode_type).astype(np.int64) + bucket_off
    merged = codes.merge(value, left_on=sel_l, right_on=sel_r, how="left").sort_values("order")  # :1774-1776
    labels = merged["labels"].to_numpy(dtype="float64", copy=True)             # deviation 1
    miss = np.isnan(labels)
    if np.isscalar(indistinct) or np.ndim(indistinct) == 0:
        labels[miss] = indistinct
    else:
        labels[miss] = np.asarray(indistinct)[miss]
    labels = labels.astype(np.int64)
    if len(nulls):
        labels[nulls] = null_off                                               # :1799-1800
    out_dtype = np.dtype(dtype) if dtype else np.dtype("int64")
    labels = labels.astype(out_dtype)
    if list_col:
        lens = [len(r) if r is not None else 0 for r in df[sel_l[0]]]
        off = np.concatenate([[0], np.cumsum(lens)])
        return pd.Series([labels[off[i]:off[i + 1]] for i in range(len(lens))], index=df.index)
    return labels


@dataclass
class CategorifyOracle:
    """fit + transform with the Categorify kwargs (categorify.py:206-343)."""
    col_groups: list
    encode_type: str = "joint"
    freq_threshold: Union[int, dict] = 0
    max_size: Union[int, dict] = 0
    num_buckets: Union[int, dict, None] = None
    dtype: Optional[object] = None
    name_sep: str = "_"
    single_table: bool = False
    vocabs: Optional[dict] = None
    categories: Dict[str, Vocab] = field(default_factory=dict)
    storage_name: Dict[str, str] = field(default_factory=dict)

    def fit(self, partitions):
        if self.freq_threshold and self.max_size:
            raise ValueError("cannot use freq_threshold param together with max_size param")
        groups = []
        for g in self.col_groups:
            g = [g] if isinstance(g, str) else list(g)
            if len(g) > 1:
                nm = _make_name(*g, sep=self.name_sep)
                for c in g:
                    self.storage_name[c] = nm                     # categorify.py:359-365
            nm = _make_name(*g, sep=self.name_sep)
            if self.vocabs and nm in self.vocabs:
                self.categories[nm] = vocab_from_series(nm, self.vocabs[nm], _resolve(self.num_buckets, nm))
            else:
                groups.append(g)
        self.categories.update(categorify_fit(
            partitions, groups, self.encode_type, self.freq_threshold, self.max_size,
            self.num_buckets, self.name_sep))
        if self.single_table:                                     # :410-415, 1884-1897
            idx = 0
            for nm, v in self.categories.items():
                v.unique = v.unique.copy()
                v.unique.index = v.unique.index + idx
                idx += v.unique.shape[0]
        return self

    def transform(self, df: pd.DataFrame) -> pd.DataFrame:
        new_df = df.copy(deep=False)
        if self.encode_type == "combo":                           # column_mapping :539-553
            mapping = {}
            for g in self.col_groups:
                g = [g] if isinstance(g, str) else list(g)
                mapping[_make_name(*g, sep=self.name_sep)] = g
        else:
            flat = []
            for g in self.col_groups:
                flat += [g] if isinstance(g, str) else list(g)
            mapping = {c: [c] for c in flat}
        for out_name, use in mapping.items():
            use_name = use[0] if len(use) == 1 else use
            if use_name != out_name or self.encode_type == "joint":
                storage = self.storage_name.get(out_name, out_name)   # :501-504
            else:
                storage = out_name
            vocab = self.categories[storage]
            nb = _resolve(self.num_buckets, storage) if not isinstance(self.num_buckets, int) \
                else self.num_buckets
            new_df[out_name] = categorify_encode(
                df, use_name, vocab, nb, self.encode_type, self.dtype, self.single_table)
        return new_df

    def embedding_sizes(self):
        """_get_embeddings_dask (categorify.py:666-684)."""
        out = {}
        for nm, v in self.categories.items():
            num_rows = OOV_OFFSET + len(v.unique)
            b = _resolve(self.num_buckets, nm)
            if isinstance(self.num_buckets, dict):
                bucket_size = b or 0
            elif isinstance(self.num_buckets, int):
                bucket_size = self.num_buckets
            else:
                bucket_size = 1
            out[nm] = emb_sz_rule(num_rows + bucket_size)
        return out
