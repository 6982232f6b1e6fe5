"""Times the tensor-core product mod p (f4la_b200_modp_gemm -> k_tc_gemm_modp) on the shapes of the
verification products of the BASELINE workloads and on a large square case, and prints the int8
tensor-pipe rate: one residue product = 16 limb products (4 x 4 unsigned 8-bit limbs)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msolve_b200.backend import Backend  # noqa: E402

P = 1073741827
SHAPES = [("katsura-12 round 9 verification", 4992, 4346, 330), ("randquad-12 verification", 3072, 4237, 154),
          ("eco-14 verification", 1792, 4096, 300), ("square", 4096, 4096, 4096), ("deep", 2048, 2048, 8192),
          # one K chunk (8192 of 58704 pivots) of the dense formulation D' = D - M B of katsura-12 round 9
          ("katsura-12 round 9 dense Schur update, 1 of 7.17 K chunks", 4992, 4346, 8192)]


def main():
    be = Backend(0)
    rng = np.random.default_rng(1)
    rows = []
    for name, M, N, K in SHAPES:
        A = rng.integers(0, P, size=(M, K), dtype=np.int64)
        B = rng.integers(0, P, size=(K, N), dtype=np.int64)
        best = None
        for _ in range(4):
            _, ms = be.modp_gemm(A, B, P)
            best = ms if best is None else min(best, ms)
        macs = 16.0 * M * N * K
        rows.append({"shape": name, "M": M, "N": N, "K": K, "kernel_ms": best,
                     "int8_TOPS": 2 * macs / (best * 1e-3) / 1e12, "residue_GMAC_per_s": M * N * K / (best * 1e-3) / 1e9})
        print(json.dumps(rows[-1]), flush=True)
    be.close()


if __name__ == "__main__":
    main()
