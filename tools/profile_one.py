"""Runs the device pipeline a few times on one saved matrix (for ncu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from msolve_b200.backend import Backend  # noqa: E402
from msolve_b200.flat import FlatMatrix, FlatResult  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    m = FlatMatrix.load(a.matrix)
    be = Backend(0)
    rm = be.upload(m)
    for _ in range(a.reps):
        out, st = be.reduce_resident(rm)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
    ref = a.matrix + ".ref.npz"
    if os.path.exists(ref):
        z = np.load(ref)
        assert out.same_rows(FlatResult(z["row_ptr"], z["cols"], z["cf"])), "result differs from the recorded one"
        print("result identical to the recorded rows")
    rm.release()
    be.close()


if __name__ == "__main__":
    main()
