"""Generates tests/golden/hash_golden.json from pandas ITSELF (pandas.util.hash_array,
the function the reference's CPU branch of merlin.core.dispatch.hash_series resolves
to).  Run in the build container; the JSON is committed and travels to the GPU box.

    python tests/golden/make_hash_golden.py
"""
import json
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260922)
    out = {"pandas": pd.__version__, "numpy": np.__version__, "cases": []}
    specials = {
        "int32": [0, 1, 2, -1, 2**31 - 1, -2**31, 12345678],
        "int64": [0, 1, 2, -1, 2**31, 2**40, -2**63, 2**63 - 1],
        "float32": [0.0, -0.0, 1.0, -1.5, 3.4e38, 1e-45, float("inf")],
        "float64": [0.0, -0.0, 1.0, -1.5, 1.7e308, 5e-324, float("inf"), float("nan")],
        "uint8": [0, 1, 255],
        "bool": [False, True],
    }
    for dt, vals in specials.items():
        arr = np.array(vals, dtype=dt)
        rnd = (rng.integers(-2**31, 2**31 - 1, 64).astype(dt) if dt.startswith("int") or dt == "uint8"
               else rng.standard_normal(64).astype(dt) if dt.startswith("float") else rng.random(64) < 0.5)
        arr = np.concatenate([arr, rnd.astype(dt)])
        h = pd.util.hash_array(arr)
        out["cases"].append({
            "dtype": dt,
            "bits": [int(x) for x in (arr.view(f"u{arr.dtype.itemsize}") if dt != "bool" else arr.astype("u1"))],
            "hash": [int(x) for x in h],
        })
    # bucket KATs
    out["mod10_of_1_2_3_int64"] = [int(x) for x in pd.util.hash_array(np.array([1, 2, 3])) % np.uint64(10)]
    with open(os.path.join(HERE, "hash_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
