"""Generates the committed fixtures tests/golden/*.npz with the UNMODIFIED
reference (oracle/_ref/libneogb_ref.so, built from /root/reference):

  * every linear-algebra call of the reference's F4 on a few small systems is
    recorded: the input matrix (flat form) and the reference's output rows;
  * the reduced Groebner basis digests of larger runs (final outputs).

Run from the repo root where /root/reference exists:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from msolve_b200 import systems  # noqa: E402
from oracle import refharness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
P31, P16, P8 = 1073741827, 65521, 251

MATRIX_CASES = [("cyclic-5", P31), ("cyclic-5", P16), ("cyclic-5", P8),
                ("cyclic-6", P31), ("katsura-5", P31), ("katsura-5", P16), ("eco-6", P31), ("eco-6", P8)]
GB_CASES = [("cyclic-5", P31), ("cyclic-6", P31), ("cyclic-7", P31), ("cyclic-7", P16), ("cyclic-7", P8),
            ("katsura-7", P31), ("katsura-8", P31), ("katsura-8", P16), ("eco-8", P31), ("eco-9", P31),
            ("randquad-6", P31), ("cyclic-8", P31), ("cyclic-8", P16), ("cyclic-8", P8), ("katsura-10", P31)]


def main():
    for name, p in MATRIX_CASES:
        gb, recs = rh.record_f4(systems.by_name(name, p), p, threads=1, laopt=2)
        out = {}
        meta = []
        for k, r in enumerate(recs):
            m, res = r.matrix, r.result
            for f in ("row_ptr", "cols", "row_cf", "cf_ptr", "cf"):
                out[f"m{k}_{f}"] = getattr(m, f)
            out[f"m{k}_dims"] = np.asarray([m.fc, m.nru, m.nrl, m.ncl, m.ncr, m.flags], dtype=np.int64)
            out[f"r{k}_row_ptr"] = res.row_ptr
            out[f"r{k}_cols"] = res.cols
            out[f"r{k}_cf"] = res.cf
            meta.append(dict(kind=r.kind, units=r.units, pivot_apps=r.pivot_apps))
        out["n"] = np.asarray(len(recs))
        out["meta"] = np.asarray(json.dumps(meta))
        out["gb_digest"] = np.asarray(gb.digest())
        path = os.path.join(HERE, f"{name}_p{p}.npz")
        np.savez_compressed(path, **out)
        print(path, len(recs), "matrices", os.path.getsize(path), "bytes")
    digests = {}
    for name, p in GB_CASES:
        rh.set_mode(rh.MODE_REFERENCE)
        gb = rh.run_f4(systems.by_name(name, p), p, threads=8, laopt=2)
        digests[f"{name}@{p}"] = dict(digest=gb.digest(), nelts=int(len(gb.lens)), nterms=int(gb.lens.sum()))
        print(name, p, digests[f"{name}@{p}"])
    with open(os.path.join(HERE, "gb_digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
