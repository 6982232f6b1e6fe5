#!/usr/bin/env python
"""bench.py -- F4 linear algebra over GF(p) on B200: mod-p Gop/s and LA seconds.

Workload (BASELINE.json configs[1]): katsura-12 (13 variables, formula
convention) modulo the 31-bit prime 1073741827, degree-reverse-lex, `-l 2`:
every linear-algebra matrix the F4 run builds (14 rounds).  One bench "step" =
one pass of the hot path over all those matrices.

  value : whole-job throughput with the matrices already resident in HBM
          (f4la_b200_reduce_resident), Gop/s = 2 * units / seconds where
          `units` is the REFERENCE algorithm's count of coefficient
          applications for the same matrices (its own counters at -t 1,
          tests/golden/workload_katsura-12_t1.json; msolve.c:2666-2673).
  e2e   : the same passes through f4la_b200_reduce() from pinned HOST buffers,
          host->device copy of every matrix and device->host copy of every
          result inside the timed region.

The matrices are produced by running the unmodified reference F4 driver
(oracle/_ref/libneogb_ref.so = reference neogb compiled as is) with its
`linear_algebra` pointer set to the GPU backend -- the drop-in scenario -- and
capturing what it hands to the backend.  Only this library's calls are timed.

  --impl reference : the reference's own CPU linear algebra (all host threads)
                     on a bounded sample of the same matrices.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

P31 = 1073741827
CPU_SAMPLE_UNITS = 4.5e10     # bounded CPU sample: rounds while cumulative reference units stay below


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_workload_stats(system: str):
    path = os.path.join(ROOT, "tests", "golden", f"workload_{system}_t1.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def generate_workload(system_name: str, p: int, threads: int):
    """Reference F4 driver + GPU backend; returns the captured LA inputs."""
    import gpu_helpers
    from msolve_b200 import systems
    from oracle import refharness as rh
    L = gpu_helpers.install_backend(mode=rh.MODE_RECORD_BACKEND)
    t0 = time.time()
    gb = rh.run_f4(systems.by_name(system_name, p), p, threads=threads, laopt=2)
    f4_s = time.time() - t0
    tot = gpu_helpers.neogb_totals(L)
    recs = [r for r in rh.records() if r.kind == 0]
    # second run in the same process, nothing recorded: the drop-in path in steady state
    # (staging buffers already sized); this is the reference's own "F4 LA time" metric (st->la_rtime)
    rh.reset(); rh.set_mode(rh.MODE_BACKEND)
    t0 = time.time()
    gb2 = rh.run_f4(systems.by_name(system_name, p), p, threads=threads, laopt=2)
    tot["warm_f4_seconds"] = time.time() - t0
    tot["warm_la_seconds"] = gb2.la_seconds
    tot["warm"] = gpu_helpers.neogb_totals(L)
    assert gb2.digest() == gb.digest()
    gpu_helpers.uninstall_backend()
    return gb, recs, f4_s, tot


def reference_units(recs, stats):
    """Per-matrix reference unit counts (from the committed -t 1 run) matched by shape."""
    if stats:
        ref_rounds = [r for r in stats["rounds"] if r["kind"] == 0]
        if len(ref_rounds) == len(recs) and all(
                (a["nru"], a["nrl"], a["ncr"], a["nnz"]) == (b.matrix.nru, b.matrix.nrl, b.matrix.ncr, b.matrix.nnz)
                for a, b in zip(ref_rounds, recs)):
            return [float(r["units"]) for r in ref_rounds], "reference counters, -t 1 run (tests/golden)"
    return None, None


def run_gpu(args, rank, world):
    import torch
    import torch.distributed as dist
    from msolve_b200.backend import Backend
    from oracle import refharness as rh

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=180))
    os.environ["F4LA_B200_DEVICE"] = str(local)

    host_threads = max(1, (os.cpu_count() or 8) // world)
    log(f"[bench] rank {rank}: generating {args.workload} mod {args.prime} matrices (reference F4 driver + GPU backend)")
    gb, recs, f4_s, tot = generate_workload(args.workload, args.prime, host_threads)
    stats = load_workload_stats(args.workload) if args.prime == P31 else None
    if stats:
        assert gb.digest() == stats["digest"], "reduced Groebner basis differs from the reference's"
        log(f"[bench] reduced GB bit-exact vs reference ({stats['nelts']} elements, {stats['nterms']} terms)")
    units_per, units_src = reference_units(recs, stats)
    if units_per is None:
        from oracle import oracle
        log("[bench] no committed reference unit counts for this workload: counting with the oracle (CPU)")
        units_per = [float(oracle.reduce(r.matrix)[1]) for r in recs]
        units_src = "oracle restatement counters"
    units = float(sum(units_per))
    log(f"[bench] {len(recs)} matrices, {units:.4g} reference units; F4 with GPU LA took {f4_s:.1f} s "
        f"(LA {tot['ms_total'] / 1e3:.2f} s: flatten {tot['ms_flatten'] / 1e3:.2f} h2d {tot['ms_h2d'] / 1e3:.2f} "
        f"prep {tot['ms_prep'] / 1e3:.2f} reduce {tot['ms_reduce'] / 1e3:.2f} echelon {tot['ms_echelon'] / 1e3:.2f} "
        f"d2h {tot['ms_d2h'] / 1e3:.2f} materialize {tot['ms_materialize'] / 1e3:.2f})")

    be = Backend(local)
    mats = [r.matrix for r in recs]
    pinned = [be.pin(m) for m in mats]
    resident = [be.upload(m) for m in mats]
    h2d_bytes = sum(int(m.row_ptr.nbytes + m.cols.nbytes + m.row_cf.nbytes + m.cf_ptr.nbytes + m.cf.nbytes) for m in mats)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pass_resident():
        agg = {"ms_prep": 0.0, "ms_reduce": 0.0, "ms_echelon": 0.0, "ms_d2h": 0.0, "kernel_launches": 0, "bytes_d2h": 0,
               "ms_hot_kernel": 0.0, "hot_kernel_launches": 0, "units": 0, "dense_macs": 0}
        for rm in resident:
            _, st = be.reduce_resident(rm, want_result=False)
            for k in agg:
                agg[k] += st[k]
        return agg

    def pass_host():
        agg = {"ms_h2d": 0.0, "kernel_launches": 0, "bytes_d2h": 0, "bytes_h2d": 0}
        digests = []
        for pm in pinned:
            out, st = be.reduce(pm.matrix)
            for k in agg:
                agg[k] += st[k]
            digests.append(out)
        return agg, digests

    if args.per_matrix:
        for r, rm, u in zip(recs, resident, units_per):
            m = r.matrix
            _, st = be.reduce_resident(rm, want_result=False)
            _, st = be.reduce_resident(rm, want_result=False)
            log(f"[bench]   {m.nru:7d}+{m.nrl:5d} x {m.ncl:7d}+{m.ncr:5d} nnz {m.nnz:9d} units {u:10.3g} rank {st['rank']:4d} "
                f"launches {st['kernel_launches']:5d} | prep {st['ms_prep']:8.2f} reduce {st['ms_reduce']:8.2f} "
                f"echelon {st['ms_echelon']:8.2f} d2h {st['ms_d2h']:6.2f} total {st['ms_total']:8.2f} ms | "
                f"{u * 8 / max(st['ms_reduce'], 1e-9) / 1e6:7.1f} GB/s-equiv dense-macs {st['dense_macs']:.3g} ref-units(known pivots) {st['units']:.4g}")

    # parity of what is being timed: the results of the host path against the reference's rows
    _, outs = pass_host()
    for o, r in zip(outs, recs):
        assert o.same_rows(r.result)

    def timed(fn, label):
        for _ in range(args.warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        aggs = [fn() for _ in range(args.steps)]
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        log(f"[bench] rank {rank} {label}: {el / args.steps * 1e3:.1f} ms/step")
        return el, aggs, clocks

    el_v, aggs_v, clocks = timed(pass_resident, "device-resident")
    el_e, aggs_e, _ = timed(lambda: pass_host()[0], "host buffers (e2e)")

    # the reference's -l 44 semantics on the same matrices: random row combinations (F4LA_FLAG_PROBABILISTIC)
    prob = None
    if not args.no_prob:
        from msolve_b200.flat import FLAG_PROBABILISTIC
        for rm in resident:
            rm.release()
        resident = []
        for m in mats:
            m.flags = FLAG_PROBABILISTIC
            resident.append(be.upload(m))
        for rm, r in zip(resident, recs):
            out, _ = be.reduce_resident(rm)
            assert out.same_rows(r.result), "probabilistic mode changed the reduced rows"
        el_p, aggs_p, _ = timed(pass_resident, "device-resident, probabilistic (-l 44 semantics)")
        prob = {"what": "same matrices with F4LA_FLAG_PROBABILISTIC (random combinations of the lower rows, the "
                        "reference's -l 42/43/44 semantics); reduced rows verified identical to the exact ones",
                "la_seconds": el_p / args.steps, "value": 2.0 * units * world / (el_p / args.steps) / 1e9, "unit": "Gop/s",
                "phases_ms": {k: sum(a[k] for a in aggs_p) / args.steps for k in ("ms_prep", "ms_reduce", "ms_echelon", "ms_d2h")}}
        for m in mats:
            m.flags = 0

    # the matrices are not needed any more
    for rm in resident:
        rm.release()
    for pm in pinned:
        pm.free()
    be.close()
    qq = None
    if not args.no_qq:
        try:
            qq = run_qq(args, rank, world, local, dist, torch)
        except Exception as e:          # the headline number must survive a failure of the extra block
            qq = {"error": repr(e)}
            log(f"[bench] rank {rank}: QQ block failed: {e!r}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    steps = args.steps
    sec_v = el_v / steps
    sec_e = el_e / steps
    gops_v = 2.0 * units * world / sec_v / 1e9
    gops_e = 2.0 * units * world / sec_e / 1e9
    ms_reduce = sum(a["ms_reduce"] for a in aggs_v) / steps
    ms_hot = sum(a["ms_hot_kernel"] for a in aggs_v) / steps
    n_hot = sum(a["hot_kernel_launches"] for a in aggs_v) / steps
    dense_macs = sum(a["dense_macs"] for a in aggs_v) / steps
    peak, peak_src = measured_peaks()
    bytes_per_unit = 8
    # algorithmic bytes of all launches of the hot kernel in a step / their summed duration
    # (= the launch-weighted average of bytes-per-launch / launch-duration)
    achieved = units * bytes_per_unit / (ms_hot / 1e3) / 1e9
    launches = sum(a["kernel_launches"] for a in aggs_v) + sum(a["kernel_launches"] for a in aggs_e)

    # CPU baseline on a bounded sample of the same matrices (reference code, all host threads)
    cpu = cpu_baseline(recs, units_per, os.cpu_count() or 1, 1)

    line = {
        "metric": "mod-p Gop/s, F4 linear algebra, katsura-12 mod 31-bit p" if args.workload == "katsura-12"
                  else f"mod-p Gop/s, F4 linear algebra, {args.workload}",
        "value": gops_v, "unit": "Gop/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": sec_v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 residues, u64 accumulators (exact)", "data": "synthetic (formula-generated system)",
        "config": {"workload": f"{args.workload} mod {args.prime}, DRL, all {len(recs)} F4 linear-algebra matrices of the run (-l 2)",
                   "units_per_step": units, "units_source": units_src,
                   "largest_matrix": max(((r.matrix.nru + r.matrix.nrl, r.matrix.ncl + r.matrix.ncr) for r in recs)),
                   "cache": "inputs larger than L2 (%.2f GB per step), no flush" % (h2d_bytes / 1e9),
                   "multi_gpu": "replicas only (one prime field F4 does not shard); each rank runs the full workload"},
        "la_seconds": sec_v,
        "phases_ms": {k: sum(a[k] for a in aggs_v) / steps for k in ("ms_prep", "ms_reduce", "ms_echelon", "ms_d2h")},
        "e2e": {"value": gops_e, "unit": "Gop/s", "la_seconds": sec_e,
                "h2d_bytes_per_step": int(sum(a["bytes_h2d"] for a in aggs_e) / steps),
                "d2h_bytes_per_step": int(sum(a["bytes_d2h"] for a in aggs_e) / steps)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_pull_flow", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch: first group (15 of 20
                     # row chunks) of the largest matrix, profiles/r01_ncu_full_k_pull_flow_v5_raw.csv; the
                     # algorithmic bytes of that launch are 5.34e10 units x 8 B x 15/20 = 3.2e11
                     "traffic": 2.36e10,
                     # what actually bounds the kernel: the multiplier pipe.  One 31 x 31 -> 96-bit multiply-add
                     # is one IMAD.WIDE.U32 (+ half an add-with-carry): 8.1e12/s measured on this GPU
                     # (tools/ubench/imad_wide.cu, profiles/README.md).  "dense_macs" counts every (pull-list
                     # entry, lower row) pair, including the all-zero source vectors the kernel skips.
                     "pipe": {"bound": "IMAD.WIDE issue (fmaheavy pipe)", "peak_mac_per_s": 8.1e12,
                              "dense_macs_per_step": dense_macs,
                              "achieved_mac_per_s": dense_macs / (ms_hot / 1e3),
                              "frac": dense_macs / (ms_hot / 1e3) / 8.1e12},
                     # L2 -> SM traffic of the same capture (every multiply-add needs its own 4-byte vector element):
                     # 7.18e9 sectors x 32 B in 16.1 ms = 14 TB/s, about the chip's L2 ceiling -- but cutting that
                     # traffic (column groups sharing vector loads) did not help: the kernel is bound by
                     # instruction issue on the multiplier-pipe path
                     "l2": {"ncu_bytes_per_launch": 2.30e11, "ncu_TB_per_s": 14.3,
                            "source": "profiles/r01_ncu_full_k_pull_flow_v5_raw.csv (lts__t_sectors_srcunit_tex_op_read)"},
                     "peak_source": peak_src, "bytes_per_unit": bytes_per_unit,
                     "launches_per_step": n_hot, "kernel_ms_per_step": ms_hot,
                     "algorithmic_bytes_per_launch": units * bytes_per_unit / max(n_hot, 1),
                     "ms_per_launch": ms_hot / max(n_hot, 1),
                     "note": "the kernel is restructured (transposed solve, vectors L2-resident): it moves far fewer "
                             "DRAM bytes than the reference algorithm's 8 B per coefficient application; the fraction "
                             "uses the survey's yardstick (reference units x 8 B / kernel time / measured HBM copy peak) "
                             "and can exceed 1"},
        "cpu_baseline": cpu,
        "f4_total_seconds_with_gpu_la": f4_s,
        "f4_drop_in": {"what": "whole reference F4 run with all LA pointers on the GPU backend, second run in the process",
                       "f4_seconds": tot["warm_f4_seconds"], "la_seconds": tot["warm_la_seconds"],
                       "la_breakdown_ms": {k: tot["warm"][k] for k in ("ms_flatten", "ms_h2d", "ms_prep", "ms_reduce",
                                                                       "ms_echelon", "ms_d2h", "ms_materialize")},
                       "host_threads": host_threads},
        "probabilistic": prob,
        "qq": qq,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


SAMPLE_FIRST = 9    # the bounded sample is drawn from the first rounds of the run


def cpu_sample(recs, units_per):
    """Bounded sample: the first rounds of the run while cumulative units stay under the cap."""
    idx, cum = [], 0.0
    for i in range(min(len(recs), SAMPLE_FIRST)):
        if cum + units_per[i] > CPU_SAMPLE_UNITS and idx:
            break
        idx.append(i); cum += units_per[i]
    return idx, cum


def run_qq(args, rank, world, local, dist, torch):
    """QQ multi-modular tracer loop (BASELINE config 5): learn once, then replay the
    trace for independent primes.  Primes are sharded over the ranks (GPU = rank),
    every rank drives its GPU with a fixed number of host threads; the only
    collective is the gather of the per-prime results (here: 64-bit digests of the
    modular bases; msolve gathers residues for CRT, msolve.c:2694-2712)."""
    import gpu_helpers
    from msolve_b200 import systems
    from oracle import refharness as rh
    system = systems.by_name(args.qq_system)
    threads = max(1, (os.cpu_count() or 8) // 8)          # fixed per GPU, whatever N is
    per_rank = args.qq_primes_per_thread * threads
    all_primes = rh.primes_above(1 << 30, 1 + per_rank * world)
    from msolve_b200 import sharding
    p0, mine = all_primes[0], sharding.shard(all_primes[1:], rank, world)[:per_rank]
    # every rank always reaches the collectives below, whatever happens locally
    local_err, res, learn, tot = None, [], {"seconds": 0.0, "nelts": 0}, {"ms_total": 0.0}
    L = gpu_helpers.install_backend()
    try:
        rh.qq_setup(system, threads=threads)
        learn = rh.qq_learn(p0)
        assert learn["rc"] == 0
        # ngpus=0: every host thread of this rank uses the rank's own GPU ($F4LA_B200_DEVICE)
        rh.qq_apply(mine[:threads], threads=threads, ngpus=0)    # warm-up
        gpu_helpers.neogb_totals(L)
    except Exception as e:
        local_err = repr(e)
    torch.cuda.set_device(local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if local_err is None:
        try:
            res, wall = rh.qq_apply(mine, threads=threads, ngpus=0)
            assert all(r["err"] == 0 and r["nelts"] == learn["nelts"] for r in res)
        except Exception as e:
            local_err, res = repr(e), []
    torch.cuda.set_device(local)
    # NCCL all-gather of the per-prime results (msolve: residues for CRT on the host)
    parts = sharding.gather_uint64([r["digest"] for r in res], dist if world > 1 else None,
                                  device=torch.device("cuda", local))
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    try:
        tot = gpu_helpers.neogb_totals(L)
    finally:
        gpu_helpers.uninstall_backend()
    if world > 1:
        t = torch.tensor([el], device=torch.device("cuda", local), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    done = sum(len(x) for x in parts)
    if local_err is not None or done != per_rank * world:
        return {"error": local_err or f"only {done} of {per_rank * world} primes completed"}
    out = {"metric": "QQ primes/s (F4 tracer application phase)", "system": f"{args.qq_system} over QQ",
           "value": per_rank * world / el, "unit": "primes/s", "n_gpus": world, "primes": per_rank * world,
           "host_threads_per_gpu": threads, "seconds": el, "scaling": "weak",
           "learn_seconds": learn["seconds"], "gb_elements": learn["nelts"],
           "la_share_of_host_thread_time": tot["ms_total"] / 1e3 / max(1e-9, sum(r["seconds"] for r in res)),
           "la_ms_per_prime": tot["ms_total"] / max(1, len(res)),
           "collective": "nccl all_gather of per-prime results" if world > 1 else "none (1 GPU)"}
    if rank == 0 and not args.no_qq_cpu:
        rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
        rh.qq_setup(system, threads=os.cpu_count() or 1)
        rh.qq_learn(p0)
        ncpu = os.cpu_count() or 1
        sample = all_primes[1:1 + 2 * ncpu]
        _, wall_cpu = rh.qq_apply(sample, threads=ncpu)
        out["cpu_baseline"] = {"value": len(sample) / wall_cpu, "unit": "primes/s", "cores": ncpu, "kind": "reference",
                               "sample": f"{len(sample)} primes, reference core_gba apply, one prime per thread"}
    return out


def cpu_baseline(recs, units_per, threads, repeats):
    from oracle import refharness as rh
    idx, cum = cpu_sample(recs, units_per)
    best = None
    for _ in range(repeats):
        sec = 0.0
        for i in idx:
            out, s, _, _ = rh.reference_la(recs[i].matrix, laopt=2, threads=threads)
            assert out.same_rows(recs[i].result)
            sec += s
        best = sec if best is None else min(best, sec)
    return {"value": 2.0 * cum / best / 1e9, "unit": "Gop/s", "cores": threads, "kind": "reference",
            "la_seconds": best,
            "sample": f"{len(idx)} of {len(recs)} matrices of the same run ({cum:.3g} of the step's units), "
                      f"reference exact_sparse_linear_algebra_ff_32 via dispatch_linear_algebra, -l 2"}


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU LA on a bounded sample, all host threads."""
    if rank != 0:
        return
    from msolve_b200 import systems
    from oracle import refharness as rh
    threads = os.cpu_count() or 1
    log(f"[bench] reference arm: recording the first {SAMPLE_FIRST} {args.workload} matrices with the reference itself "
        f"({threads} threads; the reference F4 has to run to completion)")
    stats = load_workload_stats(args.workload) if args.prime == P31 else None
    rh.reset(); rh.set_mode(rh.MODE_RECORD, SAMPLE_FIRST)
    t0 = time.time()
    gb = rh.run_f4(systems.by_name(args.workload, args.prime), args.prime, threads=threads, laopt=2)
    recs = [r for r in rh.records() if r.kind == 0]
    rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
    if stats:
        assert gb.digest() == stats["digest"]
    log(f"[bench] reference F4 run took {time.time() - t0:.1f} s, {len(recs)} matrices kept")
    units_per = [r.units for r in recs]
    if stats:
        ref_rounds = [r for r in stats["rounds"] if r["kind"] == 0][:len(recs)]
        units_per = [float(r["units"]) for r in ref_rounds]       # -t 1 counters are exact, threaded ones race
    idx, cum = cpu_sample(recs, units_per)
    for _ in range(args.warmup):
        for i in idx[:2]:
            rh.reference_la(recs[i].matrix, laopt=2, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in idx:
            rh.reference_la(recs[i].matrix, laopt=2, threads=threads)
    el = (time.perf_counter() - t0) / args.steps
    gops = 2.0 * cum / el / 1e9
    line = {
        "impl": "reference",
        "metric": "mod-p Gop/s, F4 linear algebra, katsura-12 mod 31-bit p" if args.workload == "katsura-12"
                  else f"mod-p Gop/s, F4 linear algebra, {args.workload}",
        "value": gops, "unit": "Gop/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64 accumulators (exact)", "data": "synthetic (formula-generated system)",
        "config": {"workload": f"{args.workload} mod {args.prime}, DRL, F4 linear-algebra matrices (-l 2); "
                               f"bounded sample: {len(idx)} matrices, {cum:.3g} units per step"},
        "cpu_baseline": {"value": gops, "unit": "Gop/s", "cores": threads, "kind": "reference",
                         "sample": f"{len(idx)} matrices ({cum:.3g} units) per step"},
        "e2e": {"value": gops, "unit": "Gop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath print there too (NCCL's version banner,
    the reference's verbose lines), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is
    written to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="katsura-12")
    ap.add_argument("--prime", type=int, default=P31)
    ap.add_argument("--per-matrix", action="store_true", help="log per-matrix phase times (stderr)")
    ap.add_argument("--no-qq", action="store_true", help="skip the QQ multi-modular block")
    ap.add_argument("--no-prob", action="store_true", help="skip the probabilistic-mode block")
    ap.add_argument("--no-qq-cpu", action="store_true", help="skip the CPU baseline of the QQ block")
    ap.add_argument("--qq-system", default="katsura-10")
    ap.add_argument("--qq-primes-per-thread", type=int, default=4)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world)


if __name__ == "__main__":
    main()
