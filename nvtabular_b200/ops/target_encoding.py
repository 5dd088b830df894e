"""TargetEncoding (reference nvtabular/ops/target_encoding.py:35-439).

    TE(g, f) = (sum_all(g) - sum_fold(g,f) + p * ybar) / (count_all(g) - count_fold(g,f) + p)

fit        global target mean (K1) + per group two resident hash tables with
           {size, sum} payload: key = group, and key = (fold, group-id) (K3)
transform  per row one probe + gather of the pre-combined TE value (K7);
           unseen (fold, group) -> ybar (target_encoding.py:378-380).
The fold id of a row is RandomState(fold_seed).choice(kfold, len(partition)),
drawn per partition with the same seed (target_encoding.py:182-188,427-439).
"""
import os
from typing import Dict, List

import numpy as np
import torch

from .. import engine
from ..column import Column, DeviceFrame
from ..dist import global_merge
from ..graph import ColumnSelector, Node, Tags
from .base import StatOperator
from .keyspace import ComboKeySpace, KeySpace, _leaf


def _make_name(*args, sep="_"):
    return sep.join(args)


_FOLD_CACHE = {}


def _add_fold(n, kfold, fold_seed=None, device="cuda") -> Column:
    """target_encoding.py:427-439.  The fold of row i of a partition is a function of the
    partition length only (the reference reseeds RandomState(fold_seed) per partition), so fit
    and transform of equally long partitions share ONE host draw + upload."""
    key = (int(n), int(kfold), fold_seed, str(device))
    hit = _FOLD_CACHE.get(key)
    if hit is not None:
        return Column(hit)
    col = _draw_fold(n, kfold, fold_seed, device)
    if len(_FOLD_CACHE) >= 8:
        _FOLD_CACHE.pop(next(iter(_FOLD_CACHE)))
    _FOLD_CACHE[key] = col.data
    return col


def _draw_fold(n, kfold, fold_seed=None, device="cuda") -> Column:
    typ = np.min_scalar_type(kfold * 2)
    if fold_seed is None:
        fold = np.arange(n, dtype=np.int64) % kfold
    else:
        state = np.random.RandomState(fold_seed)
        fold = state.choice(np.arange(kfold, dtype=typ), n)
    return Column(torch.from_numpy(fold.astype(np.int32)).to(device))


class _TEGroup:
    def __init__(self, names, space):
        self.names = names
        self.space = space
        self.all_vocab = None      # group key -> row of the "all" table
        self.n_groups = 0
        self.has_null = False
        self.handle = None         # key -> TE values
        self.fold_table = None     # host frame for the artefact / tests

    def key_for(self, df):
        return self.space.keys_for([df[n] for n in self.names]) if len(self.names) > 1 \
            else self.space.keys_for(df[self.names[0]])

    def gid_for(self, key: Column) -> Column:
        # position in the "all" table; the null group is its last row
        g = self.all_vocab.encode(key, null_label=self.n_groups, oov_label=-2, first_label=0,
                                  out_dtype=np.int32)
        return Column(g)


class TargetEncoding(StatOperator):
    def __init__(self, target, target_mean=None, kfold=None, fold_seed=42, p_smooth=20, out_col=None,
                 out_dtype=None, split_out=None, split_every=None, cat_cache="host", out_path=None,
                 on_host=True, name_sep="_", drop_folds=True, tree_width=None):
        super().__init__()
        self.target = Node.construct_from(target)
        self.target_mean = target_mean
        self.kfold = kfold or 3
        self.fold_seed = fold_seed
        self.p_smooth = p_smooth
        self.out_col = [out_col] if isinstance(out_col, str) else out_col
        self.out_dtype = out_dtype
        self.out_path = out_path or "./"
        self.name_sep = name_sep
        self.drop_folds = drop_folds
        self.fold_name = "__fold__"
        self.stats: Dict[str, str] = {}
        self.means: Dict[str, float] = {}
        self._groups: Dict[str, _TEGroup] = {}

    @property
    def dependencies(self):
        return self.target

    @property
    def target_columns(self) -> List[str]:
        return self.target.output_columns.names

    def _group_names(self, col_selector):
        return [list(g) if isinstance(g, tuple) else [g] for g in col_selector.grouped_names]

    # ------------------------------------------------------------------------ fit
    def fit(self, col_selector: ColumnSelector, ddf):
        parts = list(ddf)
        targets = self.target_columns
        means = None
        if self.target_mean is None:                                  # target_encoding.py:174-176
            m = engine.Moments(len(targets))
            for df in parts:
                m.accumulate([_leaf(df[t]) for t in targets])
            m.allreduce()
            means = m.result()["mean"]
        folds = [_add_fold(len(df), self.kfold, self.fold_seed, df[targets[0]].data.device)
                 for df in parts] if self.kfold > 1 else None
        groups = {}
        for names in self._group_names(col_selector):
            groups[_make_name(*names, sep=self.name_sep)] = self._fit_group(names, parts, folds, targets)
        return groups, means

    def _fit_group(self, names, parts, folds, targets) -> _TEGroup:
        if len(names) > 1:
            space = ComboKeySpace.fit([[_leaf(df[n]) for n in names] for df in parts])
        else:
            space = KeySpace.for_columns([_leaf(df[names[0]]) for df in parts])
        g = _TEGroup(names, space)
        nt = len(targets)
        agg_all = engine.HashAgg(nt)
        keys_per_part = []
        for df in parts:
            key = g.key_for(df)
            keys_per_part.append(key)
            agg_all.insert(key, [_leaf(df[t]) for t in targets])
        keys, sizes, vals, null_size, null_vals = global_merge(agg_all)
        dev = keys.device
        U = keys.numel()
        g.n_groups = U
        g.has_null = null_size > 0
        g.all_vocab = engine.Vocab.from_arrays(keys)
        sizes_f = sizes.to(torch.float64)
        if isinstance(space, ComboKeySpace) and U:
            fn = space.first_component_null_t(keys)
            count_all = torch.where(fn, torch.zeros_like(sizes_f), sizes_f)
        else:
            count_all = sizes_f
        sum_all = vals[:, :, 0] if U else torch.zeros((0, nt), dtype=torch.float64, device=dev)
        # row U = the null group (count of the first key column is 0 there)
        count_all = torch.cat([count_all, torch.zeros(1, dtype=torch.float64, device=dev)])
        nsum = torch.tensor(null_vals[:, 0] if null_vals is not None else np.zeros(nt), dtype=torch.float64, device=dev)
        sum_all = torch.cat([sum_all, nsum.reshape(1, nt)], dim=0)
        g._all = (keys, count_all, sum_all)
        if folds is not None:
            agg_f = engine.HashAgg(nt)
            for df, key, fold in zip(parts, keys_per_part, folds):
                fkey = engine.pack_keys2(fold, g.gid_for(key))
                agg_f.insert(fkey, [_leaf(df[t]) for t in targets])
            fk, fs, fv, _, _ = global_merge(agg_f)
            g._fold = (fk, fs.to(torch.float64), fv[:, :, 0] if fk.numel() else
                       torch.zeros((0, nt), dtype=torch.float64, device=dev))
        return g

    def fit_finalize(self, stats):
        groups, means = stats
        if means is not None:
            for t, m in zip(self.target_columns, means):
                self.means[t] = float(m)
        for name, g in groups.items():
            self._finalize_group(name, g)

    def _finalize_group(self, name, g: _TEGroup):
        """pre-combine the TE value of every (fold, group) — target_encoding.py:341-349, 360-363"""
        y_mean = self.target_mean or self.means
        targets = self.target_columns
        p = float(self.p_smooth)
        keys, count_all, sum_all = g._all
        ym = torch.tensor([float(y_mean[t]) for t in targets], dtype=torch.float64, device=keys.device)
        if self.kfold > 1:
            fk, count_f, sum_f = g._fold
            # second component of the packed (fold, group id) key, on the device (engine.unpack_keys2:
            # b = low word ^ 2^31; group ids are >= 0, so the unsigned value is the id)
            gid_t = (fk & 0xFFFFFFFF) ^ 0x80000000
            te = (sum_all[gid_t] - sum_f + p * ym) / ((count_all[gid_t] - count_f)[:, None] + p)
            g.handle = engine.GroupStats(fk, te, -1)
            self.stats.setdefault(_make_name(self.fold_name, *g.names, sep=self.name_sep), name)
        else:
            te = (sum_all + p * ym) / (count_all[:, None] + p)
            g.handle = engine.GroupStats(keys, te, g.n_groups)
        self.stats.setdefault(name, name)
        self._groups[name] = g

    # ------------------------------------------------------------- artefacts (cat_stats files)
    def export_tables(self, new_path) -> Dict[str, str]:
        """write the cat_stats files of every group under new_path/categories WITHOUT re-pointing
        this op (Workflow.save); -> {stats name: path}"""
        out = {}
        names_of = {}
        for key in list(self.stats):
            if not key.startswith(self.fold_name + self.name_sep):
                names_of[key] = None
        for name in names_of:
            g = self._groups.get(name)
            if g is None:
                continue
            self._write_group(name, g, os.path.join(new_path, "categories"), out)
        return out

    def _write_group(self, name, g: _TEGroup, base, record=None):
        """cat_stats.<name>.parquet (+ cat_stats.__fold___<name>.parquet): group keys, count and
        per-target sums — the reference's TargetEncoding state (target_encoding.py:190-214 via
        categorify.py:1543-1555), enough to rebuild the op after Workflow.load"""
        import pandas as pd
        from ._tables import key_columns
        os.makedirs(base, exist_ok=True)
        targets = self.target_columns
        keys, count_all, sum_all = g._all
        k = keys.cpu().numpy()
        data = key_columns(g.space, g.names, k, with_null_row=True)
        data[f"{name}_count"] = count_all.cpu().numpy().astype(np.int64)
        sa = sum_all.cpu().numpy()
        for j, t in enumerate(targets):
            data[f"{name}_{t}_sum"] = sa[:, j]
        path = os.path.join(base, f"cat_stats.{name}.parquet")
        pd.DataFrame(data).to_parquet(path)
        (self.stats if record is None else record)[name] = path
        if self.kfold > 1:
            fname = _make_name(self.fold_name, *g.names, sep=self.name_sep)
            fk, count_f, sum_f = g._fold
            a, gid = engine.unpack_keys2(fk.cpu().numpy())      # host copies only when the file is written
            kk = np.concatenate([k, np.zeros(1, dtype=k.dtype)])        # row U = the null group
            fdata = {self.fold_name: a.astype(np.int64)}
            cols = key_columns(g.space, g.names, kk[gid.astype(np.int64)])
            isnull = gid.astype(np.int64) >= len(k)
            for n_, v in cols.items():
                if isnull.any():
                    if pd.api.types.is_integer_dtype(v.dtype) and not pd.api.types.is_extension_array_dtype(v.dtype):
                        v = v.astype(v.dtype.name.replace("int", "Int"))
                    v = v.mask(isnull)
                fdata[n_] = v
            fdata[f"{fname}_count"] = count_f.cpu().numpy().astype(np.int64)
            sf = sum_f.cpu().numpy()
            for j, t in enumerate(targets):
                fdata[f"{fname}_{t}_sum"] = sf[:, j]
            fpath = os.path.join(base, f"cat_stats.{fname}.parquet")
            pd.DataFrame(fdata).to_parquet(fpath)
            (self.stats if record is None else record)[fname] = fpath

    def _group(self, name, names) -> _TEGroup:
        g = self._groups.get(name)
        if g is not None:
            return g
        # a workflow reloaded from disk: rebuild the tables from the cat_stats files
        import pandas as pd
        from ._tables import keys_from_frame
        from ..column import DeviceFrame
        path = self.stats.get(name)
        if not isinstance(path, str) or not os.path.exists(path):
            raise KeyError(name)
        targets = self.target_columns
        df = pd.read_parquet(path)
        space, keys, isnull = keys_from_frame(df, names)
        dev = keys.device
        cnt = torch.from_numpy(df[f"{name}_count"].to_numpy(dtype=np.float64)).to(dev)
        sums = torch.from_numpy(np.stack([df[f"{name}_{t}_sum"].to_numpy(dtype=np.float64) for t in targets],
                                         axis=1)).to(dev)
        keep = ~isnull
        g = _TEGroup(names, space)
        g.n_groups = int(keep.sum().item())
        g.has_null = bool(isnull.any().item())
        kk = keys[keep].contiguous()
        g.all_vocab = engine.Vocab.from_arrays(kk)
        nt = len(targets)
        if g.has_null:
            ncnt, nsum = cnt[isnull][:1] * 0.0, sums[isnull][:1]
        else:
            ncnt, nsum = torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros((1, nt), dtype=torch.float64, device=dev)
        g._all = (kk, torch.cat([cnt[keep], ncnt]), torch.cat([sums[keep], nsum], dim=0))
        if self.kfold > 1:
            fname = _make_name(self.fold_name, *names, sep=self.name_sep)
            fdf = pd.read_parquet(self.stats[fname])
            frame = DeviceFrame.from_pandas(fdf[names].reset_index(drop=True))
            key = space.keys_for([frame[n] for n in names]) if len(names) > 1 else space.keys_for(frame[names[0]])
            fold = Column(torch.from_numpy(fdf[self.fold_name].to_numpy(dtype=np.int32)).to(dev))
            fk = engine.pack_keys2(fold, g.gid_for(key)).data
            g._fold = (fk, torch.from_numpy(fdf[f"{fname}_count"].to_numpy(dtype=np.float64)).to(dev),
                       torch.from_numpy(np.stack([fdf[f"{fname}_{t}_sum"].to_numpy(dtype=np.float64)
                                                  for t in targets], axis=1)).to(dev))
        self._finalize_group(name, g)
        return g

    # ------------------------------------------------------------------ transform
    def transform(self, col_selector: ColumnSelector, df: DeviceFrame) -> DeviceFrame:
        targets = self.target_columns
        y_mean = self.target_mean or self.means
        fit_folds = self.kfold > 1
        n = len(df)
        dev = df[col_selector.names[0]].data.device
        fold = _add_fold(n, self.kfold, self.fold_seed, dev) if fit_folds else None
        new_df = DeviceFrame()
        out_dtype = np.dtype(self.output_dtype)
        for ind, names in enumerate(self._group_names(col_selector)):
            name = _make_name(*names, sep=self.name_sep)
            g = self._group(name, names)
            if isinstance(self.out_col, list):
                if ind >= len(self.out_col):
                    raise ValueError("out_col and cat_groups are different sizes.")
                out_col = self.out_col[ind]
                out_col = [out_col] if isinstance(out_col, str) else out_col
                if len(out_col) != len(targets):
                    raise ValueError("out_col and target are different sizes.")
            else:
                out_col = [f"TE_{name}_{x}" for x in targets]
            key = g.key_for(df)
            if fit_folds:
                key = engine.pack_keys2(fold, g.gid_for(key))
            outs = g.handle.gather(key, list(range(len(targets))), [float(y_mean[t]) for t in targets],
                                   [out_dtype] * len(targets))
            for c, o in zip(out_col, outs):
                new_df[c] = Column(o)
        if fit_folds and not self.drop_folds:
            new_df[self.fold_name] = Column(fold.data.to(torch.uint8))
        return new_df

    def column_mapping(self, col_selector):
        column_mapping = {}
        for group in col_selector.grouped_names:
            names = list(group) if isinstance(group, tuple) else [group]
            tag = _make_name(*names, sep=self.name_sep)
            for target_name in self.target_columns:
                column_mapping[f"TE_{tag}_{target_name}"] = [target_name, *names]
        if self.kfold > 1 and not self.drop_folds:
            column_mapping[self.fold_name] = []
        return column_mapping

    def _compute_dtype(self, col_schema, input_schema):
        if col_schema.name == self.fold_name:
            return col_schema.with_dtype(np.uint8, False, False)
        return col_schema.with_dtype(np.dtype(self.output_dtype), False, False)

    @property
    def output_dtype(self):
        return self.out_dtype or np.float32

    @property
    def output_tags(self):
        return [Tags.CONTINUOUS]

    def set_storage_path(self, new_path, copy=False):
        if copy:
            for name, g in list(self._groups.items()):
                self._write_group(name, g, os.path.join(new_path, "categories"))
        self.out_path = new_path

    def clear(self):
        self.stats = {}
        self.means = {}
        self._groups = {}
