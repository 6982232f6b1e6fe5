"""Runs whole F4 computations through the reference driver with every linear-algebra
pointer on the GPU backend, checks the reduced Groebner bases against the reference's
digests (tests/golden/gb_digests.json) and prints timings.  Optionally also runs the
reference's own CPU linear algebra for comparison (--cpu)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_io  # noqa: E402
import gpu_helpers  # noqa: E402
from msolve_b200 import systems  # noqa: E402
from oracle import refharness as rh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("systems", nargs="+")
    ap.add_argument("--prime", type=int, default=1073741827)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--cpu", action="store_true", help="also time the reference's CPU linear algebra")
    ap.add_argument("--laopt", type=int, default=2)
    a = ap.parse_args()
    want = golden_io.gb_digests()
    rows = []
    for name in a.systems:
        sysm = systems.by_name(name, a.prime)
        L = gpu_helpers.install_backend()
        try:
            t0 = time.time()
            gb = rh.run_f4(sysm, a.prime, threads=a.threads, laopt=a.laopt)
            t_gpu = time.time() - t0
            tot = gpu_helpers.neogb_totals(L)
        finally:
            gpu_helpers.uninstall_backend()
        key = f"{name}@{a.prime}"
        ok = None
        if key in want:
            ok = gb.digest() == want[key]["digest"]
        row = {"system": name, "p": a.prime, "gb_elements": int(len(gb.lens)), "gb_terms": int(gb.lens.sum()),
               "bit_exact_vs_reference": ok, "f4_seconds_gpu_la": round(t_gpu, 3),
               "la_seconds_gpu": round(gb.la_seconds, 3), "la_calls": gb.la_calls,
               "device_ms": {k: round(tot[k], 1) for k in ("ms_flatten", "ms_h2d", "ms_prep", "ms_reduce", "ms_echelon",
                                                            "ms_d2h", "ms_materialize")},
               "reference_units_known_pivots": tot["units"], "host_threads": a.threads}
        if a.cpu:
            rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
            t0 = time.time()
            gbr = rh.run_f4(sysm, a.prime, threads=a.threads, laopt=a.laopt)
            row["f4_seconds_reference"] = round(time.time() - t0, 3)
            row["la_seconds_reference"] = round(gbr.la_seconds, 3)
            row["same_as_reference_run"] = gbr.digest() == gb.digest()
        rows.append(row)
        print(json.dumps(row), flush=True)
        assert ok is not False, f"{name}: reduced Groebner basis differs from the reference"


if __name__ == "__main__":
    main()
