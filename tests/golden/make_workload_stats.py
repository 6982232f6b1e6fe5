"""Per-round statistics of a reference F4 run (shapes, the reference's own operation counters,
LA seconds, digest of the reduced Groebner basis) -> tests/golden/workload_<system>_t<threads>.json.
bench.py takes `units` (coefficient applications of the REFERENCE algorithm, exact at -t 1)
from here.  Usage: python tests/golden/make_workload_stats.py katsura-12 1   (4-5 minutes)"""
import sys, time, json
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import refharness as rh
from msolve_b200 import systems
import numpy as np
P=1073741827
name=sys.argv[1]; thr=int(sys.argv[2])
s=systems.by_name(name,P)
t=time.time()
rh.reset(); rh.set_mode(rh.MODE_RECORD)
gb=rh.run_f4(s,P,threads=thr,laopt=2)
L=rh.lib()
rounds=[]
for i in range(L.refh_num_records()):
    r=L.refh_get_record(i).contents
    m=r.inp
    n=m.nru+m.nrl
    import ctypes as C
    nnz=int(C.cast(m.row_ptr,C.POINTER(C.c_uint64))[n])
    nnz_rr=int(C.cast(m.row_ptr,C.POINTER(C.c_uint64))[m.nru])
    rounds.append(dict(kind=r.kind,nru=m.nru,nrl=m.nrl,ncl=m.ncl,ncr=m.ncr,nnz=nnz,nnz_rr=nnz_rr,ncf=m.ncf,new=r.out.nrows,units=r.units,pivot_apps=r.pivot_apps,la_seconds=r.la_seconds,flags=m.flags))
    L.refh_drop_record(i)
out=dict(system=name,p=P,threads=thr,f4_seconds=gb.f4_seconds,la_seconds=gb.la_seconds,nelts=int(len(gb.lens)),nterms=int(gb.lens.sum()),digest=gb.digest(),rounds=rounds,
         units=sum(r['units'] for r in rounds))
json.dump(out,open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f'workload_{name}_t{thr}.json'),'w'),indent=1)
print(name,'done',time.time()-t,out['units'],out['la_seconds'],out['digest'])
