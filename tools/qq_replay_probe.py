#!/usr/bin/env python
"""Where the time of a structure-cached QQ replay goes: replays a batch of primes with 1, 2, 4 ... host threads on one GPU
and prints primes/s; F4LA_B200_TAPE_DEBUG=1 adds the per-prime breakdown of the library.  Test infrastructure (drives the
reference harness under oracle/ for the tape and the initial bases)."""
import argparse, os, sys, time, threading
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="katsura-10")
    ap.add_argument("--primes", type=int, default=64)
    ap.add_argument("--threads", default="1,4,16")
    ap.add_argument("--blocking", type=int, default=0)
    ap.add_argument("--mode", default="replay", choices=["replay", "enqueue"],
                    help="replay: host-assembled, one prime at a time; enqueue: device-resident, 16 primes per wait")
    ap.add_argument("--torch", type=int, default=0, help="initialise torch's CUDA context first (as bench.py does)")
    args = ap.parse_args()
    if args.torch:
        import torch
        torch.cuda.init(); torch.zeros(1, device="cuda")
    import gpu_helpers
    from msolve_b200 import systems
    from msolve_b200.backend import Backend, Tape
    from oracle import refharness as rh
    system = systems.by_name(args.system)
    tl = [int(x) for x in args.threads.split(",")]
    primes = rh.primes_above(1 << 30, 2 + args.primes)
    L = gpu_helpers.install_backend()
    rh.qq_setup(system, threads=1)
    learn = rh.qq_learn(primes[0])
    words = learn["nterms"]
    Tape.begin(L)
    rec, _, rr = rh.qq_apply([primes[1]], threads=1, ngpus=0, residue_words=words)
    tape = Tape.end(L)
    assert tape is not None
    t0 = time.perf_counter()
    bases = [rh.qq_initial_basis(p, cap=1 << 16) for p in primes[2:]]
    print(f"initial bases: {(time.perf_counter() - t0) / len(bases) * 1e3:.2f} ms/prime, tape calls {tape.calls}", flush=True)
    import torch
    outs_t = torch.empty((len(bases), tape.residue_words), dtype=torch.int32).pin_memory()
    outs = outs_t.numpy().view(np.uint32)
    for nthr in tl:
        backends = [Backend(0) for _ in range(nthr)]
        for b in backends:
            b.set_host_mode(1, bool(args.blocking))
        def work(k):
            ms = 0.0
            if args.mode == "enqueue":
                mine = list(range(k, len(bases), nthr))
                for a0 in range(0, len(mine), Tape.MAX_INFLIGHT):
                    part = mine[a0:a0 + Tape.MAX_INFLIGHT]
                    for i in part:
                        tape.enqueue(backends[k], primes[2 + i], bases[i][0], bases[i][1], outs[i])
                    rcs = tape.collect(backends[k])
                    assert rcs == [0] * len(part), rcs
                return 0.0
            for i in range(k, len(bases), nthr):
                rc, resid, m = tape.replay(backends[k], primes[2 + i], *bases[i])
                assert rc == 0, rc
                ms += m
            return ms
        for rep in range(2):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(nthr) as ex:
                ms = sum(ex.map(work, range(nthr)))
            dt = time.perf_counter() - t0
            print(f"threads {nthr:2d} pass {rep}: {len(bases) / dt:8.1f} primes/s, {ms / len(bases):7.2f} ms inside replay per prime",
                  flush=True)
        for b in backends:
            b.close() if hasattr(b, "close") else None
    tape.free()


if __name__ == "__main__":
    main()
