"""Saves one F4 linear-algebra matrix of a run (the one with most reference
units / non-zeros) as .npz, for profiling with tools/profile_one.py.
Runs the reference F4 driver with the GPU backend (needs a GPU)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("system")
    ap.add_argument("--prime", type=int, default=bench.P31)
    ap.add_argument("--index", type=int, default=-1, help="matrix index, default: the one with most non-zeros x rows")
    ap.add_argument("--out", default="/tmp/matrix.npz")
    a = ap.parse_args()
    gb, recs, f4_s, tot = bench.generate_workload(a.system, a.prime, os.cpu_count() or 8)
    k = a.index if a.index >= 0 else max(range(len(recs)), key=lambda i: recs[i].matrix.nnz * recs[i].matrix.nrl)
    m = recs[k].matrix
    m.save(a.out)
    recs[k].result  # noqa
    import numpy as np
    np.savez(a.out + ".ref.npz", row_ptr=recs[k].result.row_ptr, cols=recs[k].result.cols, cf=recs[k].result.cf)
    print(f"saved matrix {k}: {m.nru}+{m.nrl} x {m.ncl}+{m.ncr}, nnz {m.nnz} -> {a.out}")


if __name__ == "__main__":
    main()
