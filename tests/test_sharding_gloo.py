"""CPU tests of the multi-rank path (world_size 2, gloo): primes of the QQ
multi-modular loop are sharded over the ranks, every rank replays the trace
for its primes (here with the reference's CPU linear algebra, there is no GPU
in this container), the per-prime results are gathered, and the gathered set
equals the single-process run."""
import os
import subprocess
import sys

import pytest

from msolve_b200 import sharding
from oracle import refharness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_primes_above_is_a_sieve_of_successive_primes():
    def slow(start, n):
        out, c = [], start + 1
        while len(out) < n:
            if c > 1 and all(c % d for d in range(2, int(c ** 0.5) + 1)):
                out.append(c)
            c += 1
        return out
    for start, n in ((1, 30), (100, 25), (65000, 12), (1 << 20, 40)):
        assert rh.primes_above(start, n) == slow(start, n)
    big = rh.primes_above(1 << 30, 3000)
    assert big[:3] == [1073741827, 1073741831, 1073741833] and len(set(big)) == 3000 and big == sorted(big)


def test_gather_residues_single_process_device_tensor():
    import numpy as np
    import torch
    a = np.arange(12, dtype=np.uint32).reshape(3, 4) + (1 << 31)
    assert sharding.gather_residues(a) is not None and np.array_equal(sharding.gather_residues(a), a)
    t = torch.from_numpy(a.view(np.int32))
    out = np.zeros((3, 4), dtype=np.uint32)
    assert sharding.gather_residues(t, out=out, root=0) is out and np.array_equal(out, a)


def test_shard_roundtrip():
    items = list(range(100, 117))
    for world in (1, 2, 3, 8):
        parts = [sharding.shard(items, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == items
        assert sharding.unshard(parts) == items


WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
from msolve_b200 import sharding, systems
from oracle import refharness as rh
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
primes = rh.primes_above(1 << 30, 7)
rh.set_mode(rh.MODE_REFERENCE)
rh.qq_setup(systems.by_name("katsura-6"), threads=1)
learn = rh.qq_learn(primes[0])
assert learn["rc"] == 0
mine = sharding.shard(primes[1:], rank, world)
res, _, resid = rh.qq_apply(mine, threads=1, residue_words=learn["nterms"])
assert all(r["err"] == 0 and r["nterms"] == learn["nterms"] for r in res)
parts = sharding.gather_uint64([r["digest"] for r in res], dist)
# the payload msolve lifts by CRT: every prime's residues, in batch order, on every rank
allres = sharding.gather_residues(resid, dist)
assert allres.shape == (len(primes) - 1, learn["nterms"])
for i, p in enumerate(primes[1:]):
    assert int(allres[i].max()) < p
# gather to the rank that lifts by CRT only; the batch may also come as an int32 tensor (the replay leaves it in HBM),
# in chunks of equal local size, each chunk = a contiguous range of the batch
import numpy as np, torch
root_only = sharding.gather_residues(resid, dist, root=0)
assert (root_only is None) == (rank != 0)
if rank == 0:
    assert np.array_equal(root_only, allres)
k = min(len(mine), (len(primes) - 1) // world)            # equal shards: the head of every rank's rows
t = torch.from_numpy(np.ascontiguousarray(resid[:k]).view(np.int32))
out = np.zeros((k * world, learn["nterms"]), dtype=np.uint32) if rank == 0 else None
got = sharding.gather_residues(t, dist, out=out, root=0)
if rank == 0:
    assert got is out and np.array_equal(out, allres[:k * world])
if rank == 0:
    print("RESULT " + json.dumps(sharding.unshard(parts)))
    print("RESID " + json.dumps([rh.residue_digest(row) for row in allres]))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(not rh.available(), reason="compiled reference not built")
def test_qq_prime_sharding_world2_gloo(tmp_path):
    from msolve_b200 import systems
    primes = rh.primes_above(1 << 30, 7)
    rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
    rh.qq_setup(systems.by_name("katsura-6"), threads=1)
    assert rh.qq_learn(primes[0])["rc"] == 0
    learn = rh.qq_learn(primes[0])
    single, _, resid = rh.qq_apply(primes[1:], threads=1, residue_words=learn["nterms"])
    want = [r["digest"] for r in single]
    want_resid = [rh.residue_digest(row) for row in resid]

    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
        capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    import json
    got = json.loads(line[len("RESULT "):])
    assert got == want
    line = [l for l in out.stdout.splitlines() if l.startswith("RESID ")][-1]
    assert json.loads(line[len("RESID "):]) == want_resid
