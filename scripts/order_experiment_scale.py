#!/usr/bin/env python
"""ORDER EXPERIMENT at scale (CPU, ~1 min on 8 threads; needs /root/reference for the real test1 station geometry):
acceptance statistics of the "serial prefix -> fast-iterative ranks -> wavefront replay -> local verification" plan over
1 440 solves of the reference's test1 survey and 400 solves of a 300-station Yunnan-shaped survey, coarse march and
refined source box separately.  See scripts/order_experiment.py for what the fields mean."""
import sys, os, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from dazimsurftomo_b200 import formats as fm, synthetic
from oracle import pyoracle as po
REF="/root/reference/example/test1_syn_foward"
p = fm.read_para_forward(REF+"/para.in")
depz, vs = fm.read_model(REF+"/MODVs.true", p.nx, p.ny, p.nz)
sv = fm.read_surfdata(REF+"/"+p.datafile, p.kmaxRc)
pv, L = po.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=8)
def agg(name, rs, t):
    fl = lambda r: r["verify_order_flags"]+r["verify_key_increase_flags"]>0
    print(json.dumps(dict(case=name, solves=len(rs), rule_mismatch_total=sum(r["rule_mismatch"] for r in rs),
        replay_exact=sum(r["sorted_fim_mismatch"]==0 for r in rs), verified_exact=sum((r["sorted_fim_mismatch"]==0 and not fl(r)) for r in rs),
        flagged=sum(fl(r) for r in rs), wrong_but_not_flagged=sum((r["sorted_fim_mismatch"]>0 and not fl(r)) for r in rs),
        solves_with_interacting_ties=sum(r["pair_ties"]>0 for r in rs), dag_levels=[min(r["dag_levels"] for r in rs), max(r["dag_levels"] for r in rs)],
        cpu_s=round(t,1))), flush=True)
jobs=[(k,s) for k in range(0,36,3) for s in range(int(sv.nsrcsurf1[k]))]     # 12 periods x 120 sources = 1440 solves
def coarse(j): k,s=j; return po.fmm_order_stats(p.nx,p.ny,p.goxd,p.gozd,p.dvxd,p.dvzd,pv[:,k],float(sv.scxf[s,k]),float(sv.sczf[s,k]),prefix=64)
def refined(j): k,s=j; return po.fmm_order_stats(p.nx,p.ny,p.goxd,p.gozd,p.dvxd,p.dvzd,pv[:,k],float(sv.scxf[s,k]),float(sv.sczf[s,k]),refined=True,prefix=4)
with ThreadPoolExecutor(8) as ex:
    t0=time.time(); rc=list(ex.map(coarse, jobs)); agg("reference test1 survey, every 3rd period x all 120 sources: COARSE march, serial prefix 64", rc, time.time()-t0)
    t0=time.time(); rr=list(ex.map(refined, jobs)); agg("same solves: REFINED source box, serial prefix 4", rr, time.time()-t0)
    fl = lambda r: r["verify_order_flags"]+r["verify_key_increase_flags"]>0
    both=sum((not fl(a)) and (not fl(b)) and a["sorted_fim_mismatch"]==0 and b["sorted_fim_mismatch"]==0 for a,b in zip(rc,rr))
    print(json.dumps({"case":"same solves: both stages verified-exact", "solves": len(jobs), "both": both}), flush=True)
    w = synthetic.yunnan_shaped(nsta=300); tb = synthetic.proxy_tables(w)
    jobs2=[(k,s) for k in (0,12,24,35) for s in range(0,299,3)]      # 4 periods x 100 sources
    def c2(j): k,s=j; return po.fmm_order_stats(w.nx,w.ny,w.goxd,w.gozd,w.dvxd,w.dvzd,tb["pvRc"][:,k],float(w.sv.scxf[s,k]),float(w.sv.sczf[s,k]),prefix=64)
    def r2(j): k,s=j; return po.fmm_order_stats(w.nx,w.ny,w.goxd,w.gozd,w.dvxd,w.dvzd,tb["pvRc"][:,k],float(w.sv.scxf[s,k]),float(w.sv.sczf[s,k]),refined=True,prefix=4)
    t0=time.time(); a=list(ex.map(c2, jobs2)); agg("Yunnan-shaped survey (300 stations), 4 periods x 100 sources: COARSE march, serial prefix 64", a, time.time()-t0)
    t0=time.time(); b=list(ex.map(r2, jobs2)); agg("same solves: REFINED source box, serial prefix 4", b, time.time()-t0)
    both=sum((not fl(x)) and (not fl(y)) and x["sorted_fim_mismatch"]==0 and y["sorted_fim_mismatch"]==0 for x,y in zip(a,b))
    print(json.dumps({"case":"same solves: both stages verified-exact", "solves": len(jobs2), "both": both}), flush=True)
