"""Digests of the reduced Groebner bases computed by the UNMODIFIED reference (8 threads, -l 2)
for the larger systems; merged into tests/golden/gb_digests.json.
Usage: python tests/golden/make_digests.py katsura-11 eco-12 randquad-10 eco-13 randquad-11 randquad-12 eco-14"""
import sys, json, time
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refharness as rh
from msolve_b200 import systems
P=1073741827
out={}
for name in sys.argv[1:]:
    t=time.time()
    rh.set_mode(rh.MODE_REFERENCE)
    gb=rh.run_f4(systems.by_name(name,P),P,threads=8,laopt=2)
    out[f"{name}@{P}"]=dict(digest=gb.digest(),nelts=int(len(gb.lens)),nterms=int(gb.lens.sum()))
    print(name,out[f"{name}@{P}"],round(time.time()-t,1),'s LA',round(gb.la_seconds,1),flush=True)
    path = os.path.join(HERE, 'gb_digests.json')
    d = json.load(open(path)) if os.path.exists(path) else {}
    d.update(out)
    json.dump(d, open(path, 'w'), indent=1, sort_keys=True)
