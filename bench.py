#!/usr/bin/env python
"""bench.py -- F4 linear algebra over GF(p) on B200: mod-p Gop/s and LA seconds.

Workload (default, BASELINE.json configs[1]): katsura-12 (13 variables, formula
convention) modulo the 31-bit prime 1073741827, degree-reverse-lex, `-l 2`:
every linear-algebra matrix the F4 run builds.  One bench "step" = one pass of
the hot path over all those matrices.  `--workload cyclic-8 --prime 65521|251`,
`--workload eco-14`, `--workload randquad-12` give the lines of configs 3 and 4.

  value : whole-job throughput with the matrices already resident in HBM
          (f4la_b200_reduce_resident), Gop/s = 2 * units / seconds where
          `units` is the REFERENCE algorithm's count of coefficient
          applications for the same matrices (its own counters at -t 1,
          tests/golden/workload_*_t1.json; msolve.c:2666-2673).
  e2e   : the same matrices through the PLUGIN POINTER the reference calls,
          f4la_b200_linear_algebra(mat_t *, bs_t *, bs_t *, md_t *): gathering
          the malloc()ed rows into staging, host->device copies, kernels,
          device->host copy and materialisation of the result rows as neogb
          rows are all inside the timed calls.  `e2e.flat_abi` keeps the number
          of the flat C ABI (f4la_b200_reduce from pinned host arrays).

The matrices are produced by running the unmodified reference F4 driver
(oracle/_ref/libneogb_ref*.so = reference neogb compiled as is) with all five
linear-algebra pointers set to the GPU backend -- the drop-in scenario -- and
capturing what it hands to the backend.  Only this library's calls are timed.

  --impl reference : the reference's own CPU linear algebra (all host threads)
                     on the IDENTICAL matrices, recorded from the reference's
                     own F4 run.

With --gpus N > 1 (torchrun, one rank per GPU) the headline becomes the metric
north_star names for several GPUs: QQ primes/s of the multi-modular tracer
application phase (structure-cached, device-resident replay of a recorded
application run, include/f4la_b200_neogb.h), a FIXED batch of primes sharded
over the ranks, all host cores in use at every N, residues gathered with NCCL
to the rank that would lift them by CRT; the per-GPU replicas of the
single-prime workload are reported under `replicas`.  At N = 1 the same block
is reported under `qq` next to the single-prime headline.
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

ALL_CPUS = os.sched_getaffinity(0)       # before any OpenMP runtime loads and pins this thread (see below)

# the reference's OpenMP row loops (and the adapter's gather of the rows): one thread per core, no migration.
# Not under torchrun: every rank would bind its threads to the same first cores.
if int(os.environ.get("WORLD_SIZE", "1")) == 1:
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
# the multi-modular section drives one stream per host thread: give each its own hardware queue (the default is 8)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

P31 = 1073741827
CPU_SAMPLE_UNITS = 4.5e10     # cpu_baseline leg of the GPU arm: bounded sample of the step (about 10-30 s of CPU work)
BYTES_PER_UNIT = {4: 8, 2: 6, 1: 5}      # 4 B column index + coefficient


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_workload_stats(system: str, prime: int):
    if prime != P31:
        return None
    path = os.path.join(ROOT, "tests", "golden", f"workload_{system}_t1.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def golden_digest(system: str, prime: int):
    import golden_io
    return golden_io.gb_digests().get(f"{system}@{prime}")


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
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

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
        sm, mx, util, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                if len(r) > 7:
                    util.append(float(r[7]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "gpu_util_pct_mean": float(np.mean(util)) if util else None}


def generate_workload(system_name: str, p: int, threads: int):
    """Reference F4 driver + GPU backend (all five LA pointers); returns the captured LA inputs,
    the cold (first run of the process) and warm (second run) drop-in timings."""
    import gpu_helpers
    from msolve_b200 import systems
    from oracle import refharness as rh
    # run 1: first use of the backend in this process, nothing recorded -- what a one-shot msolve process pays
    L = gpu_helpers.install_backend(mode=rh.MODE_BACKEND)
    assert rh.hooked_pointers() == 0b11111
    t0 = time.time()
    gb = rh.run_f4(systems.by_name(system_name, p), p, threads=threads, laopt=2)
    cold = {"f4_seconds": time.time() - t0, "la_seconds": gb.la_seconds, "totals": gpu_helpers.neogb_totals(L)}
    # run 2: steady state (second run of the process): the reference's own "F4 LA time" metric (st->la_rtime)
    rh.reset(); rh.set_mode(rh.MODE_BACKEND)
    t0 = time.time()
    gb2 = rh.run_f4(systems.by_name(system_name, p), p, threads=threads, laopt=2)
    warm = {"f4_seconds": time.time() - t0, "la_seconds": gb2.la_seconds, "totals": gpu_helpers.neogb_totals(L)}
    # run 3: the harness records what the driver hands to the backend (its copies perturb the timing: not reported)
    rh.reset(); rh.set_mode(rh.MODE_RECORD_BACKEND)
    gb3 = rh.run_f4(systems.by_name(system_name, p), p, threads=threads, laopt=2)
    recs = [r for r in rh.records() if r.kind == 0]
    gpu_helpers.neogb_totals(L)
    assert gb2.digest() == gb.digest() and gb3.digest() == gb.digest()
    gpu_helpers.uninstall_backend()
    return gb, recs, cold, warm


def reference_units(recs, stats):
    """Per-matrix reference unit counts (from the committed -t 1 run) matched by shape."""
    if stats:
        ref_rounds = [r for r in stats["rounds"] if r["kind"] == 0]
        if len(ref_rounds) == len(recs) and all(
                (a["nru"], a["nrl"], a["ncr"], a["nnz"]) == (b.matrix.nru, b.matrix.nrl, b.matrix.ncr, b.matrix.nnz)
                for a, b in zip(ref_rounds, recs)):
            return [float(r["units"]) for r in ref_rounds], "reference counters, -t 1 run (tests/golden)"
    return None, None


def count_units(recs, stats, be):
    units_per, src = reference_units(recs, stats)
    if units_per is not None:
        return units_per, src
    total_nnz = sum(r.matrix.nnz for r in recs)
    if total_nnz < 2.5e7:
        from oracle import oracle
        log("[bench] no committed reference unit counts for this workload: counting with the oracle restatement (CPU)")
        return [float(oracle.reduce(r.matrix)[1]) for r in recs], "oracle restatement counters (identical to the reference's at -t 1)"
    log("[bench] no committed reference unit counts for this workload: using the device's counters "
        "(exact for the known pivots, the reference adds the applications of the new pivots)")
    out = []
    for r in recs:
        rm = be.upload(r.matrix)
        _, st = be.reduce_resident(rm, want_result=False)
        rm.release()
        out.append(float(st["units"]))
    return out, "device counters: reference units of the reduction by the known pivots (lower bound of the reference's total)"


def workload_label(args, n):
    return f"{args.workload} mod {args.prime}, DRL, all {n} F4 linear-algebra matrices of the run (-l 2)"


def metric_name(args):
    if args.workload == "katsura-12" and args.prime == P31:
        return "mod-p Gop/s, F4 linear algebra, katsura-12 mod 31-bit p"
    return f"mod-p Gop/s, F4 linear algebra, {args.workload} mod {args.prime}"


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
                                timeout=datetime.timedelta(seconds=600))
    os.environ["F4LA_B200_DEVICE"] = str(local)
    cpu = rh.host_cpu()
    host_threads = max(1, (os.cpu_count() or 8) // world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    log(f"[bench] rank {rank}: generating {args.workload} mod {args.prime} matrices (reference F4 driver + GPU backend, "
        f"{host_threads} host threads)")
    gb, recs, cold, warm = generate_workload(args.workload, args.prime, host_threads)
    want = golden_digest(args.workload, args.prime)
    if want:
        assert gb.digest() == want["digest"], "reduced Groebner basis differs from the reference's"
        log(f"[bench] reduced GB bit-exact vs the unmodified reference ({want['nelts']} elements, {want['nterms']} terms)")
    be = Backend(local)
    stats = load_workload_stats(args.workload, args.prime)
    units_per, units_src = count_units(recs, stats, be)
    units = float(sum(units_per))
    ct, wt = cold["totals"], warm["totals"]
    log(f"[bench] {len(recs)} matrices, {units:.4g} reference units; F4 with GPU LA: first run {cold['f4_seconds']:.1f} s "
        f"(LA {cold['la_seconds']:.3f} s), second run {warm['f4_seconds']:.1f} s (LA {warm['la_seconds']:.3f} s: "
        f"flatten {wt['ms_flatten'] / 1e3:.3f} h2d {wt['ms_h2d'] / 1e3:.3f} prep {wt['ms_prep'] / 1e3:.3f} "
        f"reduce {wt['ms_reduce'] / 1e3:.3f} echelon {wt['ms_echelon'] / 1e3:.3f} d2h {wt['ms_d2h'] / 1e3:.3f} "
        f"materialize {wt['ms_materialize'] / 1e3:.3f})")

    mats = [r.matrix for r in recs]
    cf_bytes = mats[0].cf_bytes
    pinned = [be.pin(m) for m in mats]
    resident = [be.upload(m) for m in mats]
    in_bytes = sum(int(m.row_ptr.nbytes + m.cols.nbytes + m.row_cf.nbytes + m.cf_ptr.nbytes + m.cf.nbytes) for m in mats)

    KEYS_V = ("ms_prep", "ms_reduce", "ms_echelon", "ms_d2h", "kernel_launches", "bytes_d2h", "ms_hot_kernel",
              "hot_kernel_launches", "units", "dense_macs", "ms_schur", "schur_macs")

    def pass_resident():
        agg = dict.fromkeys(KEYS_V, 0.0)
        for rm in resident:
            _, st = be.reduce_resident(rm, want_result=False)
            for k in agg:
                agg[k] += st[k]
        return agg

    def pass_flat():
        agg = {"ms_h2d": 0.0, "kernel_launches": 0, "bytes_d2h": 0, "bytes_h2d": 0}
        outs = []
        for pm in pinned:
            out, st = be.reduce(pm.matrix)
            for k in agg:
                agg[k] += st[k]
            outs.append(out)
        return agg, outs

    if args.per_matrix:
        for r, rm, u in zip(recs, resident, units_per):
            m = r.matrix
            _, st = be.reduce_resident(rm, want_result=False)
            _, st = be.reduce_resident(rm, want_result=False)
            log(f"[bench]   {m.nru:7d}+{m.nrl:5d} x {m.ncl:7d}+{m.ncr:5d} nnz {m.nnz:9d} units {u:10.3g} rank {st['rank']:4d} "
                f"launches {st['kernel_launches']:5d} | prep {st['ms_prep']:8.2f} reduce {st['ms_reduce']:8.2f} "
                f"echelon {st['ms_echelon']:8.2f} d2h {st['ms_d2h']:6.2f} total {st['ms_total']:8.2f} ms | "
                f"dense-macs {st['dense_macs']:.3g} ref-units(known pivots) {st['units']:.4g}")

    # what is timed returns the rows the drop-in run produced (whose final basis equals the reference's);
    # per-matrix parity against the reference's own rows is tests/test_gpu_configs.py
    _, outs = pass_flat()
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
        el = max_over_ranks(el)
        log(f"[bench] rank {rank} {label}: {el / args.steps * 1e3:.1f} ms/step")
        return el, aggs, clocks

    el_v, aggs_v, clocks = timed(pass_resident, "device-resident")
    el_f, aggs_f, _ = timed(lambda: pass_flat()[0], "flat C ABI from pinned host buffers")

    # ---- e2e: the plugin pointer on mat_t (what msolve calls) ----
    import ctypes as C
    import gpu_helpers
    Lib = gpu_helpers.install_backend()
    la_fn = C.cast(Lib.f4la_b200_linear_algebra, C.c_void_p).value

    def pass_plugin(check=False):
        sec = 0.0
        for r in recs:
            out, _, _ = rh.call_entry(r.matrix, 0, la_fn, threads=host_threads)   # builds mat_t/bs_t/md_t, calls, reads back
            sec += rh.last_side()["call_seconds"]                               # wall time of the pointer call alone
            if check:
                assert out.same_rows(r.result)
        return sec

    pass_plugin(check=True)
    for _ in range(max(0, args.warmup - 1)):
        pass_plugin()
    gpu_helpers.neogb_totals(Lib)
    barrier()
    plugin_secs = [pass_plugin() for _ in range(args.steps)]
    barrier()
    pt = gpu_helpers.neogb_totals(Lib)
    gpu_helpers.uninstall_backend()
    sec_e = max_over_ranks(float(np.mean(plugin_secs)))
    log(f"[bench] rank {rank} plugin pointer on mat_t (e2e): {sec_e * 1e3:.1f} ms/step")

    # the reference's -l 44 semantics on the same matrices: random row combinations (F4LA_FLAG_PROBABILISTIC)
    prob = None
    if not args.no_prob and args.prime > 32768:
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

    mac_peak = be.measure_mac_peak(200.0)
    for rm in resident:
        rm.release()
    for pm in pinned:
        pm.free()
    be.close()

    qq = None
    if not args.no_qq:
        try:
            qq = run_qq(args, rank, world, local, dist, torch, cpu)
        except Exception as e:          # the single-prime numbers must survive a failure of the QQ block
            qq = {"error": repr(e)}
            log(f"[bench] rank {rank}: QQ block failed: {e!r}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    steps = args.steps
    sec_v, sec_f = el_v / steps, el_f / steps
    gops = lambda s: 2.0 * units * world / s / 1e9                                         # noqa: E731
    ms_hot = sum(a["ms_hot_kernel"] for a in aggs_v) / steps
    n_hot = sum(a["hot_kernel_launches"] for a in aggs_v) / steps
    dense_macs = sum(a["dense_macs"] for a in aggs_v) / steps
    ms_schur = sum(a["ms_schur"] for a in aggs_v) / steps
    schur_macs = sum(a["schur_macs"] for a in aggs_v) / steps
    ms_flow = ms_hot - ms_schur
    flow_macs = dense_macs - schur_macs
    known_units = sum(a["units"] for a in aggs_v) / steps
    peak, peak_src = measured_peaks()
    bpu = BYTES_PER_UNIT[cf_bytes]
    achieved = units * bpu / (ms_hot / 1e3) / 1e9 if ms_hot > 0 else 0.0
    launches = (sum(a["kernel_launches"] for a in aggs_v) + sum(a["kernel_launches"] for a in aggs_f)
                + pt["kernel_launches"])

    # one-shot process: a fresh interpreter without torch runs the reference F4 with the backend once
    # (tools/run_f4_gpu.py) -- what the msolve binary would pay, CUDA context and module load included
    one_shot = None
    if world == 1 and not args.no_cold:
        try:
            runs = []
            for _ in range(2):       # two fresh processes: the first one on a box also pays for paging the libraries in
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_f4_gpu.py"), args.workload, "--prime",
                                      str(args.prime), "--threads", str(host_threads)], capture_output=True, text=True,
                                     timeout=900, env=dict(os.environ, F4LA_B200_DEVICE=str(local)))
                runs.append(json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1]))
            row = min(runs, key=lambda r: r["la_seconds_gpu"])
            one_shot = {"what": "fresh process (no torch), one F4 run with all LA pointers on the GPU: tools/run_f4_gpu.py; "
                                "two such processes, the faster one reported, both listed",
                        "f4_seconds": row["f4_seconds_gpu_la"], "la_seconds": row["la_seconds_gpu"],
                        "la_seconds_all_runs": [r["la_seconds_gpu"] for r in runs],
                        "la_breakdown_ms": row["device_ms"], "bit_exact_vs_reference": row["bit_exact_vs_reference"]}
            log(f"[bench] one-shot process: F4 {row['f4_seconds_gpu_la']} s, LA {[r['la_seconds_gpu'] for r in runs]} s")
        except Exception as e:
            one_shot = {"error": repr(e)}

    cpu_base = None
    if world == 1 and not args.no_cpu:
        cpu_base = cpu_baseline(recs, units_per, os.cpu_count() or 1, 3, cpu)

    roofline = {
        # what binds the kernel (ncu, profiles/): L2 -> SM vector traffic and the issue rate of the multiplier
        # pipe, not HBM; `achieved / peak / frac` keep the contract's definition on the survey's yardstick
        # (reference units x algorithmic bytes / kernel time vs the measured HBM copy peak) and can exceed 1
        # because the restructured kernel keeps the vectors in L2 and moves far fewer DRAM bytes
        "bound": "l2+imad", "kernel": "k_pull_flow (+ k_schur_gemm for the columns without pivot)",
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "frac_survey_yardstick": achieved / peak,
        "traffic": 1.16e10 if args.workload == "katsura-12" else None,
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the dominant k_pull_flow launch (largest matrix, "
                          "16 of its 20 row chunks), profiles/r02_ncu_full_k_pull_flow_k_schur_gemm_raw.csv: 10.6 GB + 0.97 GB "
                          "(the k_schur_gemm launch that follows it: 13.8 GB + 0.40 GB); algorithmic bytes of that launch 3.2e11",
        "binding": {
            "what": "the reduction by the known pivots runs on two pipes: the dataflow solve of the pivot columns on the "
                    "multiplier pipe (IMAD.WIDE), the product for the columns without pivot on the tensor pipe "
                    "(tcgen05.mma.kind::i8, 16 int8 products per residue product, non-zero 64x64 tiles only)",
            "pivot_columns": {
                "kernel": "k_pull_flow", "ms_per_step": ms_flow, "nominal_macs_per_step": flow_macs,
                "peak_mac_per_s": mac_peak,
                "peak_source": "f4la_b200_measure_mac_peak(): the kernel's 96-bit multiply-accumulate with register "
                               "operands, measured in this run on this GPU",
                "achieved_mac_per_s": flow_macs / (ms_flow / 1e3) if ms_flow > 0 else None,
                "frac_nominal": flow_macs / (ms_flow / 1e3) / mac_peak if ms_flow > 0 and mac_peak else None,
                "note": "nominal = every (pull-list entry, lower row) pair, the all-zero source vectors the kernel skips "
                        "included: the executed fraction is lower; ncu: issue slots 49 %, L2 40 %, warps active 50 %",
            },
            "other_columns": {
                "kernel": "k_schur_a_split + k_schur_gemm + k_schur_combine", "ms_per_step": ms_schur,
                "nominal_macs_per_step": schur_macs,
                "achieved_residue_mac_per_s": schur_macs / (ms_schur / 1e3) if ms_schur > 0 else None,
                "tensor_pipe_active_ncu": [0.659, 0.718],
                "source": "profiles/r02_ncu_full_k_pull_flow_k_schur_gemm_raw.csv "
                          "(sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active of the two launches of the "
                          "largest matrix)",
            } if ms_schur > 0 else None,
            "l2": {"ncu_bytes_per_launch": 5.69e10, "ncu_TB_per_s": 9.1,
                   "source": "profiles/r02_ncu_full_k_pull_flow_k_schur_gemm_raw.csv "
                             "(lts__t_sectors_srcunit_tex_op_read of the dominant k_pull_flow launch, 6.27 ms)"}},
        "peak_source": peak_src, "bytes_per_unit": bpu, "launches_per_step": n_hot, "kernel_ms_per_step": ms_hot,
        "algorithmic_bytes_per_launch": units * bpu / max(n_hot, 1), "ms_per_launch": ms_hot / max(n_hot, 1),
        "reference_units_known_pivots_counted_on_device": known_units,
    }
    single = {
        "metric": metric_name(args), "value": gops(sec_v), "unit": "Gop/s", "ms_per_step": sec_v * 1e3,
        "la_seconds": sec_v,
        "phases_ms": {k: sum(a[k] for a in aggs_v) / steps for k in ("ms_prep", "ms_reduce", "ms_echelon", "ms_d2h")},
    }
    e2e = {"value": gops(sec_e), "unit": "Gop/s", "la_seconds": sec_e,
           "what": "f4la_b200_linear_algebra(mat_t*, bs_t*, bs_t*, md_t*) on reference-built matrices: row gather, "
                   "H2D, kernels, D2H and materialisation of neogb rows inside the timed calls (sum of the wall clock "
                   "of the pointer calls of a step; building mat_t between calls is the F4 driver's job)",
           "h2d_bytes_per_step": int(pt["bytes_h2d"] / steps), "d2h_bytes_per_step": int(pt["bytes_d2h"] / steps),
           "breakdown_ms_per_step": {k: pt[k] / steps for k in ("ms_flatten", "ms_h2d", "ms_prep", "ms_reduce",
                                                                 "ms_echelon", "ms_d2h", "ms_materialize")},
           "host_threads": host_threads,
           "flat_abi": {"value": gops(sec_f), "unit": "Gop/s", "la_seconds": sec_f,
                        "what": "f4la_b200_reduce() from pinned flat arrays",
                        "h2d_bytes_per_step": int(sum(a["bytes_h2d"] for a in aggs_f) / steps),
                        "d2h_bytes_per_step": int(sum(a["bytes_d2h"] for a in aggs_f) / steps)},
           "one_shot_process": one_shot,
           "first_f4_run_of_this_process": {"what": "first F4 run inside the bench process (torch and NCCL loaded): first "
                                                    "device / staging allocations and module load included",
                                 "f4_seconds": cold["f4_seconds"], "la_seconds": cold["la_seconds"],
                                 "la_breakdown_ms": {k: ct[k] for k in ("ms_flatten", "ms_h2d", "ms_prep", "ms_reduce",
                                                                        "ms_echelon", "ms_d2h", "ms_materialize")}},
           "warm_f4_run": {"f4_seconds": warm["f4_seconds"], "la_seconds": warm["la_seconds"]}}
    config = {"workload": workload_label(args, len(recs)), "units_per_step": units, "units_source": units_src,
              "largest_matrix": max(((r.matrix.nru + r.matrix.nrl, r.matrix.ncl + r.matrix.ncr) for r in recs)),
              "cache": "inputs larger than L2 (%.2f GB per step), no flush" % (in_bytes / 1e9)
                       if in_bytes > 2.6e8 else "L2 flushed by the working set of the step: %.2f GB of inputs plus the "
                       "vector blocks written by every call" % (in_bytes / 1e9),
              "host_cpu": cpu}
    line = {"n_gpus": world, "steps": steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
            "dtype": "u32 residues, 96-bit accumulators (exact)" if cf_bytes == 4 else "u32 residues, u64 accumulators (exact)",
            "data": "synthetic (formula-generated system)", "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu_base, "probabilistic": prob}
    if world == 1 or qq is None or "error" in qq:
        config["multi_gpu"] = "replicas only for one prime field (an F4 run does not shard); QQ primes shard, see `qq`"
        line.update(single)
        line.update({"scaling": "weak", "config": config, "e2e": e2e, "qq": qq})
    else:
        # several GPUs: the axis north_star names is the QQ multi-modular loop
        config["workload"] = qq["workload"]
        config["multi_gpu"] = "primes of the QQ tracer application phase sharded over the ranks (one GPU each), residues " \
                              "all-gathered with NCCL; `replicas` = the single-prime workload run independently on every GPU"
        line.update({"metric": qq["metric"], "value": qq["value"], "unit": "primes/s", "ms_per_step": qq["seconds"] * 1e3,
                     "scaling": "strong", "config": config,
                     "e2e": {"value": qq["value"], "unit": "primes/s", "what": qq["what"],
                             "h2d_bytes_per_step": qq["h2d_bytes"], "d2h_bytes_per_step": qq["d2h_bytes"]},
                     "qq": qq, "replicas": dict(single, e2e=e2e, scaling="weak")})
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_sample(recs, units_per):
    """Bounded sample: the first rounds of the run while cumulative units stay under the cap (a small workload
    is taken whole)."""
    idx, cum = [], 0.0
    for i in range(len(recs)):
        if cum + units_per[i] > CPU_SAMPLE_UNITS and idx:
            break
        idx.append(i); cum += units_per[i]
    return idx, cum


def run_qq(args, rank, world, local, dist, torch, cpu):
    """QQ multi-modular tracer loop (BASELINE config 5; msolve.c:1628-1632 learn, :1815-1824 apply): learn once,
    then the application phase for a FIXED batch of independent primes, sharded over the ranks (GPU = rank), every
    rank using its share of ALL host cores (cpu_count / world threads).  Two ways through the same batch:

      replay   : the structure-cached tape (f4la_b200_neogb.h): one application run is recorded through the drop-in
                 pointers, every other prime is replayed with no symbolic host work -- coefficient pools in, reduced
                 basis out, all linear algebra on structures resident in HBM.  This is the headline.
      drop_in  : the reference's own per-prime F4 replay on the host with every linear-algebra call on the GPU
                 through the plugin pointer (bound by the host's symbolic work).

    The collective is the all-gather of the residues (coefficients of every modular basis), the payload msolve
    lifts by CRT on the host (msolve.c:2694-2712)."""
    import gpu_helpers
    from concurrent.futures import ThreadPoolExecutor
    from msolve_b200 import sharding, systems
    from msolve_b200.backend import Backend, Tape
    from oracle import refharness as rh
    system = systems.by_name(args.qq_system)
    ncpu = os.cpu_count() or 8
    threads = max(1, ncpu // world)
    batch = args.qq_primes
    all_primes = rh.primes_above(1 << 30, 2 + batch + world * 64)
    p0, p_tape = all_primes[0], all_primes[1]
    warm_primes = sharding.shard(all_primes[2 + batch:], rank, world)
    mine = sharding.shard(all_primes[2:2 + batch], rank, world)
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.set_device(local)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    L = gpu_helpers.install_backend()
    err = None
    out = {}
    try:
        rh.qq_setup(system, threads=threads)
        learn = rh.qq_learn(p0)
        assert learn["rc"] == 0
        words = learn["nterms"]
        # ---- the tape: one application run through the drop-in pointers ----
        Tape.begin(L)
        rec, _, rec_resid = rh.qq_apply([p_tape], threads=1, ngpus=0, residue_words=words)
        tape = Tape.end(L)
        assert tape is not None and rec[0]["err"] == 0 and tape.residue_words == words
        # host threads per GPU and how they wait: tried on warm-up primes, the fastest runs the timed batch.
        # (spinning waits are fastest while every thread has a core; with fewer cores than it takes to keep the GPU
        # busy -- 8 ranks on a 16-core host -- more threads that sleep in their waits do better)
        modes = sorted({(max(1, threads // 2), False), (min(threads, 16), False),
                        (min(2 * threads, 32), True), (min(4 * threads, 32), True)})
        backends = [Backend(local) for _ in range(max(m[0] for m in modes))]

        init_lock = threading.Lock()

        all_cpus = ALL_CPUS

        def replay_some(k, nthr, primes, res, res_ptr=None):
            """host thread k of nthr: primes k, k + nthr, ... of the batch, residues written straight into their rows
            (res: host array; res_ptr: address of the [len(primes), words] uint32 batch buffer in HBM instead)"""
            # OMP_PROC_BIND pinned the main thread to the first core and new threads inherit that mask
            os.sched_setaffinity(0, all_cpus)
            failed, t0 = [], time.perf_counter()
            mine_k = list(range(k, len(primes), nthr))
            # device-resident replay: up to 16 primes are queued on the context's stream, one wait per group
            for a0 in range(0, len(mine_k), Tape.MAX_INFLIGHT):
                part = mine_k[a0:a0 + Tape.MAX_INFLIGHT]
                for i in part:
                    with init_lock:                  # the harness (reference code + GMP) is driven from one thread at a time
                        cf, off = rh.qq_initial_basis(primes[i], cap=1 << 16)
                    tape.enqueue(backends[k], primes[i], cf, off, res[i] if res_ptr is None else res_ptr + 4 * words * i)
                rcs = tape.collect(backends[k])
                assert len(rcs) == len(part)
                failed += [i for i, rc in zip(part, rcs) if rc != 0]
            return failed, (time.perf_counter() - t0) * 1e3

        def replay_batch(primes, mode, res=None, res_dev=None):
            """res_dev: int32 torch tensor [len(primes), words] in HBM that receives the residues (they then never touch
            the host before the gather); else a host array"""
            nthr, blocking = mode
            for b in backends[:nthr]:
                b.set_host_mode(1, blocking)
                b.set_grid_cap(64)                   # many contexts share the GPU: none may occupy every SM
            if res is None and res_dev is None:
                res = np.empty((len(primes), words), dtype=np.uint32)
            ptr = None if res_dev is None else int(res_dev.data_ptr())
            with ThreadPoolExecutor(max_workers=nthr) as ex:
                outs = list(ex.map(lambda k: replay_some(k, nthr, primes, res, ptr), range(nthr)))
            redo = [i for failed, _ in outs for i in failed]
            if redo:                                 # bad prime / other support than recorded: the ordinary per-prime path
                _, _, rr = rh.qq_apply([primes[i] for i in redo], threads=min(threads, len(redo)), ngpus=0, residue_words=words)
                for i, row in zip(redo, rr):
                    if res_dev is None:
                        res[i] = row
                    else:
                        res_dev[i].copy_(torch.from_numpy(np.ascontiguousarray(row).view(np.int32)))
            return (res if res_dev is None else res_dev), len(redo), sum(o[1] for o in outs)

        # the tape reproduces the recorded prime, and the drop-in path's result for other primes
        nmax = len(backends)
        chk, _, _ = replay_batch([p_tape] + (warm_primes * 2)[:2 * nmax], (nmax, False))     # every context warm
        trials = []
        for mode in modes:
            sample = (warm_primes * 8)[:max(32, 6 * mode[0])]
            t0 = time.perf_counter()
            replay_batch(sample, mode)
            trials.append({"host_threads": mode[0], "blocking_waits": mode[1],
                           "primes_per_s": len(sample) / (time.perf_counter() - t0)})
        best = max(range(len(modes)), key=lambda i: trials[i]["primes_per_s"])
        mode = modes[best]
        nthr = mode[0]
        log(f"rank {rank}: QQ replay host modes {trials} -> {mode}")
        assert np.array_equal(chk[0], rec_resid[0]), "tape replay differs from the recorded application run"
        _, _, want = rh.qq_apply(warm_primes[:2], threads=min(2, threads), ngpus=0, residue_words=words)
        assert np.array_equal(chk[1:3], want), "tape replay differs from the drop-in application run"
        # page-locked once, outside the timed region (msolve keeps its modular images for the whole run)
        init_bytes = int(rh.qq_initial_basis(p_tape, cap=1 << 16)[1][-1]) * 4            # what a replay sends to the device
        resid_dev = torch.empty((len(mine), words), dtype=torch.int32, device=dev)       # this rank's residues, in HBM
        gathered_pinned = sharding.pinned_rows(batch, words) if rank == 0 else None
        # warm-up of the collective on a few rows (NCCL connects its channels on the first large call)
        sharding.gather_residues(resid_dev[:min(4, len(mine))].contiguous(), dist if world > 1 else None, device=dev, root=0)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        # The batch goes through in chunks: while the host threads replay chunk c+1, a background thread gathers chunk c --
        # the residues of every rank meet on the device (one all-gather over NVLink per chunk) and rank 0, where msolve
        # lifts by CRT (msolve.c:2694-2712), copies them to its host.  With equal shards the local rows [a, b) of every
        # rank are the rows [a * world, b * world) of the batch.
        L_loc = len(mine)
        n_chunks = 4 if batch % world == 0 and L_loc >= 1024 else 1
        bounds = [(c * L_loc // n_chunks, (c + 1) * L_loc // n_chunks) for c in range(n_chunks)]
        gather_err = []

        def gather_chunks(q):
            try:
                torch.cuda.set_device(local)
                while True:
                    item = q.get()
                    if item is None:
                        return
                    a, b_ = item
                    sharding.gather_residues(resid_dev[a:b_], dist if world > 1 else None, device=dev,
                                             out=gathered_pinned[a * world:b_ * world] if rank == 0 else None, root=0)
            except Exception as e:                   # noqa: BLE001 -- reported by the main thread
                gather_err.append(e)

        import queue
        gq = queue.Queue()
        gthread = threading.Thread(target=gather_chunks, args=(gq,))
        gthread.start()
        fallbacks, replay_ms = 0, 0.0
        for a, b_ in bounds:
            _, fb, ms = replay_batch(mine[a:b_], mode, res_dev=resid_dev[a:b_])
            fallbacks += fb
            replay_ms += ms
            gq.put((a, b_))
        t_replay = time.perf_counter() - t0
        gq.put(None)
        gthread.join()
        if gather_err:
            raise gather_err[0]
        gathered = gathered_pinned
        torch.cuda.synchronize()
        el_local = time.perf_counter() - t0
        el = max_over_ranks(el_local)
        util = sampler.stop() if rank == 0 else None
        resid = resid_dev.cpu().numpy().view(np.uint32)                  # (untimed) every rank's own rows, for the checks below
        del resid_dev
        if rank == 0:
            assert gathered.shape == (batch, words)
            for j in range(len(mine)):
                assert rh.residue_digest(gathered[rank + j * world][:64]) == rh.residue_digest(resid[j][:64])
        for b in backends:
            b.close()
        tape.free()

        # ---- the per-prime drop-in path on the head of the same batch (it is 10x slower: a bounded sample) ----
        n_drop = min(batch, args.qq_dropin_primes)
        mine2 = sharding.shard(all_primes[2:2 + n_drop], rank, world)
        rh.qq_apply(warm_primes[:threads], threads=threads, ngpus=0)
        gpu_helpers.neogb_totals(L)
        barrier()
        t0 = time.perf_counter()
        res2, _, resid2 = rh.qq_apply(mine2, threads=threads, ngpus=0, residue_words=words)
        gathered2 = sharding.gather_residues(resid2, dist if world > 1 else None, device=dev, root=0)
        torch.cuda.synchronize()
        el2 = max_over_ranks(time.perf_counter() - t0)
        tot = gpu_helpers.neogb_totals(L)
        assert all(r["err"] == 0 for r in res2)
        if rank == 0:
            assert np.array_equal(gathered2, gathered[:n_drop]), "replayed residues differ from the drop-in path's"
        assert np.array_equal(resid2, resid[:len(mine2)]), "replayed residues differ from the drop-in path's"
        host_sec = max(1e-9, sum(r["seconds"] for r in res2))
        out = {"metric": "QQ primes/s (F4 tracer application phase)",
               "workload": f"{args.qq_system} over QQ, tracer application phase, fixed batch of {batch} primes above 2^30",
               "what": "structure-cached, device-resident replay: one application run recorded through the drop-in pointers, "
                       "every prime of the batch replayed from the tape (initial basis mod p in, reduced basis out; the basis "
                       "elements of the prime stay in HBM between its LA calls, no symbolic host work, one wait per 16 primes), "
                       "residues of all primes all-gathered; verified equal to the per-prime drop-in path",
               "value": batch / el, "unit": "primes/s", "n_gpus": world, "primes": batch, "seconds": el,
               "host_threads_per_gpu": nthr, "host_threads_total": nthr * world, "blocking_waits": mode[1],
               "host_mode_trials": trials, "host_cores_per_gpu": threads, "scaling": "strong",
               "tape_calls": tape.calls, "replay_fallbacks": int(fallbacks),
               "replay_ms_per_prime": replay_ms / max(1, len(mine)),
               "rank0_replay_seconds": t_replay, "rank0_gather_seconds": el_local - t_replay,
               "learn_seconds": learn["seconds"], "gb_elements": learn["nelts"], "residue_words_per_prime": words,
               "gather_chunks": n_chunks, "gathered_bytes": int(batch * words * 4), "h2d_bytes": int(len(mine) * init_bytes), "d2h_bytes": int(batch * words * 4) if rank == 0 else 0,
               "gpu0_util_pct_mean": util.get("gpu_util_pct_mean") if util else None,
               "collective": "nccl all_gather of the residues on the devices (uint32 [primes, words]) per chunk of the batch, "
                             "overlapped with the replay of the next chunk; batch copied to the host of rank 0"
                             if world > 1 else "none (1 GPU; device -> host copy per chunk, overlapped)",
               "drop_in": {"what": "the reference's per-prime F4 replay on the host, every LA call on the GPU through the plugin "
                                   "pointer (bound by the host's symbolic work, msolve.c:1815-1824)",
                           "value": n_drop / el2, "unit": "primes/s", "seconds": el2, "primes": n_drop,
                           "host_threads_per_gpu": threads,
                           "la_share_of_host_thread_time": tot["ms_total"] / 1e3 / host_sec,
                           "la_ms_per_prime": tot["ms_total"] / max(1, len(res2))}}
    except Exception as e:
        err = repr(e)
        import traceback
        traceback.print_exc()
    finally:
        gpu_helpers.uninstall_backend()
    # every rank reaches this collective, whatever happened locally
    flag = max_over_ranks(1.0 if err else 0.0)
    if flag:
        return {"error": err or "another rank failed"}
    if rank == 0 and not args.no_qq_cpu:
        rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
        rh.qq_setup(system, threads=ncpu)
        rh.qq_learn(p0)
        sample = all_primes[2:2 + min(batch, 4 * ncpu)]
        rh.qq_apply(sample[:ncpu], threads=ncpu)
        _, wall_cpu = rh.qq_apply(sample, threads=ncpu)
        out["cpu_baseline"] = {"value": len(sample) / wall_cpu, "unit": "primes/s", "cores": ncpu, "kind": "reference",
                               "host_cpu": cpu,
                               "sample": f"{len(sample)} primes of the same batch, reference core_gba apply, one prime per thread"}
    return out


def cpu_baseline(recs, units_per, threads, repeats, cpu):
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
            "la_seconds": best, "repeats": repeats, "host_cpu": cpu,
            "simd": "AVX-512" if rh.selected_lib_path().endswith("avx512.so") else "AVX2",
            "sample": f"{len(idx)} of {len(recs)} matrices of the same run ({cum:.3g} of the step's units), best of {repeats}, "
                      f"reference exact_sparse_linear_algebra_ff_* via dispatch_linear_algebra, -l 2, "
                      f"OMP_PROC_BIND={os.environ.get('OMP_PROC_BIND')} OMP_PLACES={os.environ.get('OMP_PLACES')}"}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU linear algebra on the IDENTICAL workload -- every LA matrix of its own
    F4 run -- with all host threads.  A step is one pass over all matrices; when K passes do not fit the time budget
    fewer are timed (the line says how many), the workload is never cut."""
    if rank != 0:
        return
    from msolve_b200 import systems
    from oracle import refharness as rh
    threads = os.cpu_count() or 1
    cpu = rh.host_cpu()
    log(f"[bench] reference arm: recording the {args.workload} mod {args.prime} matrices with the reference itself "
        f"({threads} threads, {cpu['model']}, {'AVX-512' if rh.selected_lib_path().endswith('avx512.so') else 'AVX2'} build)")
    t_begin = time.time()
    stats = load_workload_stats(args.workload, args.prime)
    rh.reset(); rh.set_mode(rh.MODE_RECORD)
    gb = rh.run_f4(systems.by_name(args.workload, args.prime), args.prime, threads=threads, laopt=2)
    recs = [r for r in rh.records() if r.kind == 0]
    rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
    want = golden_digest(args.workload, args.prime)
    if want:
        assert gb.digest() == want["digest"]
    f4_s = time.time() - t_begin
    log(f"[bench] reference F4 run took {f4_s:.1f} s (LA {gb.la_seconds:.1f} s inside F4), {len(recs)} matrices")
    units_per, units_src = reference_units(recs, stats)
    if units_per is None:
        if args.prime >= 65536:
            units_per = [r.units for r in recs]        # the reference's counters of the recording run (racy with threads)
            units_src = f"reference counters of the recording run at -t {threads} (the counters race between threads)"
        else:
            from oracle import oracle
            units_per = [float(oracle.reduce(r.matrix)[1]) for r in recs]
            units_src = "oracle restatement counters (identical to the reference's at -t 1)"
    units = float(sum(units_per))

    def one_pass():
        sec = 0.0
        for r in recs:
            _, s, _, _ = rh.reference_la(r.matrix, laopt=2, threads=threads)
            sec += s
        return sec

    t0 = time.time()
    first = one_pass()                                  # full warm-up pass over the whole workload
    per_pass_wall = time.time() - t0
    budget = args.ref_budget_s - (time.time() - t_begin)
    warm_extra = max(0, min(args.warmup - 1, int(budget / max(per_pass_wall, 1e-6)) - 1))
    for _ in range(warm_extra):
        one_pass()
    budget = args.ref_budget_s - (time.time() - t_begin)
    steps = max(1, min(args.steps, int(budget / max(per_pass_wall, 1e-6))))
    secs = [one_pass() for _ in range(steps)]
    el = float(np.mean(secs))
    gops = 2.0 * units / el / 1e9
    line = {
        "impl": "reference", "metric": metric_name(args), "value": gops, "unit": "Gop/s", "n_gpus": world,
        "steps": steps, "steps_requested": args.steps, "warmup": 1 + warm_extra, "ms_per_step": el * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64 accumulators (exact)", "data": "synthetic (formula-generated system)",
        "config": {"workload": workload_label(args, len(recs)), "units_per_step": units, "units_source": units_src,
                   "host_cpu": cpu, "step_seconds_min_max": [float(min(secs)), float(max(secs))],
                   "first_pass_seconds": first, "la_seconds_inside_its_own_f4_run": gb.la_seconds,
                   "time_budget_s": args.ref_budget_s},
        "cpu_baseline": {"value": gops, "unit": "Gop/s", "cores": threads, "kind": "reference", "host_cpu": cpu,
                         "simd": "AVX-512" if rh.selected_lib_path().endswith("avx512.so") else "AVX2",
                         "sample": f"all {len(recs)} matrices ({units:.4g} units) per step, {steps} steps, "
                                   f"OMP_PROC_BIND={os.environ.get('OMP_PROC_BIND')} OMP_PLACES={os.environ.get('OMP_PLACES')}"},
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
    ap.add_argument("--workload", default="katsura-12", help="katsura-N, cyclic-N, eco-N, randquad-N")
    ap.add_argument("--prime", type=int, default=P31, help="1073741827 (31-bit), 65521 (16-bit), 251 (8-bit)")
    ap.add_argument("--per-matrix", action="store_true", help="log per-matrix phase times (stderr)")
    ap.add_argument("--no-qq", action="store_true", help="skip the QQ multi-modular block")
    ap.add_argument("--no-prob", action="store_true", help="skip the probabilistic-mode block")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cold", action="store_true", help="skip the one-shot (fresh process) run")
    ap.add_argument("--no-qq-cpu", action="store_true", help="skip the CPU baseline of the QQ block")
    ap.add_argument("--qq-system", default="katsura-10")
    ap.add_argument("--qq-dropin-primes", type=int, default=512,
                    help="primes of the batch the per-prime drop-in path is timed (and compared) on")
    ap.add_argument("--qq-primes", type=int, default=8192, help="size of the fixed prime batch of the QQ block")
    ap.add_argument("--ref-budget-s", type=float, default=300.0,
                    help="--impl reference: wall-clock budget; fewer steps are timed when K passes do not fit")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world)


if __name__ == "__main__":
    main()
