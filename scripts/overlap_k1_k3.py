"""Feasibility: does the FD depth-kernel stage (K1, FP64-bound, 0.87 s) hide behind the eikonal stage (K3, latency-bound,
5.2 s) when both run at once on one B200?  Two handles (two streams), K3 launched first so that its 1 000 CTAs are
resident, K1 squeezed into the registers that are left (k_disp<8>: 64 registers)."""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dazimsurftomo_b200 import api, synthetic  # noqa: E402

w = synthetic.s200()
h = api.Handle(0); h2 = api.Handle(0)
os.environ["DAZIM_KDISP_MINB"] = "8"     # run the 64-register build once alone: a first launch of a kernel with larger stack
                                          # frames resizes the local-memory pool, which waits for everything running
pv2, L = api.depthkernelTI(w.vs, w.depz, w.tRc, w.sublayers, handle=h)
pv, svs, svp, srho = api.depthkernel(w.vs, w.depz, w.tRc, w.sublayers, handle=h2)
k1_alone = h2.times["kernels_ms"]
tb = dict(pvRc=pv, sen_vs=svs, sen_vp=svp, sen_rho=srho, Lsen_Gsc=L)
plan = api.Plan(2, w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv, tb, handle=h)
tm = plan.run(); tm = plan.run()
alone = dict(k1_ms=k1_alone, run_ms=tm["total_ms"], fmm_ms=tm["fmm_ms"])
out = dict(alone=alone)
for minb in ("8", "8"):
    os.environ["DAZIM_KDISP_MINB"] = minb
    res = {}

    def k1():
        time.sleep(0.15)            # after the eikonal kernel has been launched
        t0 = time.perf_counter()
        api.depthkernel(w.vs, w.depz, w.tRc, w.sublayers, handle=h2)
        res["k1_wall_ms"] = 1e3 * (time.perf_counter() - t0); res["k1_ms"] = h2.times["kernels_ms"]

    t = threading.Thread(target=k1)
    t0 = time.perf_counter()
    t.start()
    tm = plan.run()
    res["run_ms"] = tm["total_ms"]; res["fmm_ms"] = tm["fmm_ms"]
    t.join()
    res["both_wall_ms"] = 1e3 * (time.perf_counter() - t0)
    out["overlapped_minb%s_%d" % (minb, len(out))] = res
print(json.dumps(out))
plan.close()
