#!/usr/bin/env python
"""Next-stage measurement: LSMR on the joint G of a synthetic workload.
Reports device ms per iteration and the achieved HBM bandwidth of the two sparse products
(algorithmic bytes: 8 B per non-zero per product = value + index; vectors live in L2)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dazimsurftomo_b200 import api, synthetic
    from oracle import pyoracle as po
    wl = sys.argv[1] if len(sys.argv) > 1 else "S200-125"
    itn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    w = synthetic.s200(src_per_period=int(wl.split("-")[1])) if wl.startswith("S200-") else synthetic.s200()
    h = api.Handle(0)
    pv, svs, svp, srho = api.depthkernel(w.vs, w.depz, w.tRc, w.sublayers, handle=h)
    pv2, L = api.depthkernelTI(w.vs, w.depz, w.tRc, w.sublayers, handle=h)
    tb = dict(pvRc=pv, sen_vs=svs, sen_vp=svp, sen_rho=srho, Lsen_Gsc=L)
    plan = api.Plan(2, w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv, tb, handle=h)
    plan.run()
    out = plan.fetch()
    nnz = plan.nnz
    rows = np.repeat(np.arange(1, plan.rows + 1, dtype=np.int32), np.diff(out["rowptr"]).astype(np.int64))
    m = plan.rows; n = 3 * (w.nx - 2) * (w.ny - 2) * (w.nz - 1)
    plan.close()
    rng = np.random.default_rng(1)
    b = rng.standard_normal(m).astype(np.float32)
    x, info = api.LSMR(m, n, rows, out["col"], out["val"], b, damp=1.0, atol=0.0, btol=0.0, conlim=0.0, itnlim=itn, localSize=10, handle=h)
    t_iter = info["solve_ms"] / info["itn"]
    peak = 6551.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = 2 * nnz * 8 / (t_iter * 1e-3) / 1e9
    # CPU port on a bounded number of iterations
    t0 = time.time()
    ox, oi = po.lsmr(m, n, rows, out["col"], out["val"], b, damp=1.0, atol=0.0, btol=0.0, conlim=0.0, itnlim=3, localSize=10)
    cpu_iter = (time.time() - t0) / oi["itn"]
    print(json.dumps({"stage": "LSMR (lsmrModule.f90:36) on joint G", "workload": w.name, "m": m, "n": n, "nnz": int(nnz),
                      "iterations": info["itn"], "istop": info["istop"], "ms_per_iteration": t_iter,
                      "setup_ms": info["setup_ms"], "spmv_algorithmic_GBps": gbs, "hbm_peak_GBps": peak, "frac": gbs / peak,
                      "cpu_port_s_per_iteration_1core": cpu_iter, "speedup_vs_1core": cpu_iter / (t_iter * 1e-3)}))


if __name__ == "__main__":
    main()
