"""Loads the committed reference fixtures written by tests/golden/make_golden.py."""
import glob
import json
import os

import numpy as np

from msolve_b200.flat import FlatMatrix, FlatResult

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_cases(path):
    """-> list of (FlatMatrix, reference FlatResult, meta dict)"""
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    out = []
    for k in range(int(z["n"])):
        fc, nru, nrl, ncl, ncr, flags = (int(x) for x in z[f"m{k}_dims"])
        m = FlatMatrix(fc, nru, nrl, ncl, ncr, z[f"m{k}_row_ptr"], z[f"m{k}_cols"], z[f"m{k}_row_cf"],
                       z[f"m{k}_cf_ptr"], z[f"m{k}_cf"], flags=flags)
        r = FlatResult(z[f"r{k}_row_ptr"], z[f"r{k}_cols"], z[f"r{k}_cf"])
        out.append((m, r, meta[k]))
    return out


def gb_digests():
    with open(os.path.join(GOLDEN, "gb_digests.json")) as f:
        return json.load(f)
