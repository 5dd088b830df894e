"""The partition-parallel oracle runner (bench.py's reference arm) gives the single-process
oracle's answer: merging per-partition partials is exact (categorify.py:1054-1137,
moments.py:80-116)."""
import numpy as np
import pandas as pd

import oracle
from oracle.categorify import CategorifyOracle
from oracle.parallel import run_criteo_workflow


def test_parallel_runner_matches_single_process_oracle():
    rng = np.random.default_rng(11)
    n = 160_000                      # > 3 x 50 000 rows: three worker processes
    df = pd.DataFrame({
        "C1": rng.integers(0, 5000, n).astype("float64"),
        "C2": rng.integers(0, 7, n).astype("float64"),
        "I1": rng.integers(-3, 100, n).astype("float64"),
        "I2": rng.normal(size=n),
    })
    for c in df.columns:
        df.loc[rng.random(n) < 0.05, c] = np.nan
    cats, conts = ["C1", "C2"], ["I1", "I2"]
    tf, tt, vocabs, means, stds = run_criteo_workflow(df, cats, conts, workers=3)
    assert tf > 0 and tt > 0
    single = CategorifyOracle(cats).fit([df])
    for c in cats:
        a, b = vocabs[c].unique, single.categories[c].unique
        np.testing.assert_array_equal(a[c].to_numpy(), b[c].to_numpy())
        np.testing.assert_array_equal(a[f"{c}_size"].to_numpy(), b[f"{c}_size"].to_numpy())
        np.testing.assert_array_equal(a.index.to_numpy(), b.index.to_numpy())
    filled = oracle.fill_missing(df, conts, 0)
    m1, s1 = oracle.normalize_fit(filled, conts)
    for c in conts:
        assert abs(means[c] - m1[c]) <= 1e-12 * max(1.0, abs(m1[c]))
        assert abs(stds[c] - s1[c]) <= 1e-12 * max(1.0, abs(s1[c]))
    # one worker = the reference without a Dask client
    _, _, v1, mm, ss = run_criteo_workflow(df, cats, conts, workers=1)
    for c in cats:
        np.testing.assert_array_equal(v1[c].unique[c].to_numpy(), vocabs[c].unique[c].to_numpy())
    assert mm == means and ss == stds or all(abs(mm[c] - means[c]) < 1e-12 for c in conts)
