#!/usr/bin/env python
"""EXPERIMENT (CPU only; oracle/fim_experiment.cpp): how far is an order-free fixed point of the reference's local
eikonal solver -- what a fast-iterative-method kernel converges to -- from the reference's heap fast-marching result?

  1. coarse travel-time fields, source by source, on the reference's test1 model (71 x 71) and on the S200 model
     (996 x 996): relative differences, fraction of nodes that differ at all;
  2. the whole forward path on the test1 subset (1 240 rays): travel times, ray footprints, G sparsity pattern and
     values with the fixed-point fields in place of the heap-march fields.

    python scripts/fim_vs_fmm.py [--s200]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dazimsurftomo_b200 import formats as fm, synthetic   # noqa: E402
from oracle import pyoracle as po                          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "test1")


def fim_field(*a):
    return po.fmm_source_fim(*a)


def compare_fields(name, nx, ny, goxd, gozd, dvxd, dvzd, pv, sources):
    rel_max, frac, sweeps, t_fim, t_fmm, rel_99 = [], [], [], 0.0, 0.0, []
    for scx, scz in sources:
        t0 = time.time(); a = po.fmm_source(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz)["ttn"]; t_fmm += time.time() - t0
        t0 = time.time(); b, sw = fim_field(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz); t_fim += time.time() - t0
        rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-6)
        rel_max.append(float(rel.max())); frac.append(float((a != b).mean())); sweeps.append(sw)
        rel_99.append(float(np.quantile(rel, 0.99)))
    return {"case": name, "sources": len(sources), "nodes_per_field": int(a.size),
            "fraction_of_nodes_that_differ": [min(frac), float(np.mean(frac)), max(frac)],
            "max_relative_difference": [min(rel_max), float(np.mean(rel_max)), max(rel_max)],
            "p99_relative_difference_mean": float(np.mean(rel_99)),
            "gauss_seidel_passes_x4_orderings": [min(sweeps), max(sweeps)],
            "cpu_s_heap_march": t_fmm, "cpu_s_fixed_point": t_fim}


def main():
    out = []
    p = fm.read_para_forward(os.path.join(GOLD, "para.in"))
    depz, vs = fm.read_model(os.path.join(GOLD, "MODVs.true"), p.nx, p.ny, p.nz)
    gc = fm.read_gcgs(os.path.join(GOLD, "MODGc.true"), p.nx, p.ny, p.nz)
    gs = fm.read_gcgs(os.path.join(GOLD, "MODGs.true"), p.nx, p.ny, p.nz)
    sv = fm.read_surfdata(os.path.join(GOLD, "surfdata_subset.dat"), p.kmaxRc)
    pv, L = po.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=8)
    srcs = [(float(sv.scxf[s, k]), float(sv.sczf[s, k])) for k in range(sv.kmax) for s in range(int(sv.nsrcsurf1[k]))]
    for k in (0, 3):
        out.append(compare_fields("test1 model, period index %d" % k, p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd,
                                  pv[:, k], srcs[:10]))
    if "--s200" in sys.argv:
        w = synthetic.s200(src_per_period=2)
        tb = synthetic.proxy_tables(w)
        s2 = [(float(w.sv.scxf[s, 0]), float(w.sv.sczf[s, 0])) for s in range(2)]
        out.append(compare_fields("S200 model (996 x 996 nodes), period index 0", w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd,
                                  tb["pvRc"][:, 0], s2))
    # whole forward path with the fixed-point fields
    pvf, svs, svp, srho, _ = po.depthkernel(vs, depz, p.tRc, p.sublayers, nthreads=8)
    tb = dict(pvRc=pv, Lsen_Gsc=L, sen_vs=svs, sen_vp=svp, sen_rho=srho)
    args = (2, vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv)
    os.environ.pop("ORC_FIM_EXPERIMENT", None)
    ref = po.gbuild(*args, tables=tb, nthreads=8)
    os.environ["ORC_FIM_EXPERIMENT"] = "1"
    alt = po.gbuild(*args, tables=tb, nthreads=8, experiments=True)
    os.environ.pop("ORC_FIM_EXPERIMENT", None)
    ka = set(zip(ref["row"].tolist(), ref["col"].tolist())); kb = set(zip(alt["row"].tolist(), alt["col"].tolist()))
    common = ka & kb
    da = dict(zip(zip(ref["row"].tolist(), ref["col"].tolist()), ref["rw"].tolist()))
    db = dict(zip(zip(alt["row"].tolist(), alt["col"].tolist()), alt["rw"].tolist()))
    relv = np.array([abs(da[k] - db[k]) / max(abs(da[k]), 1e-12) for k in common])
    rel_t = np.abs(ref["dsurf"] - alt["dsurf"]) / ref["dsurf"]
    rows_changed = len({r for r, _ in (ka ^ kb)})
    out.append({"case": "forward path on the test1 subset (joint G), fixed-point fields instead of heap-march fields",
                "rays": int(sv.dall), "travel_time_relative_difference": {"max": float(rel_t.max()), "mean": float(rel_t.mean()),
                                                                          "rays_that_differ": int((ref["dsurf"] != alt["dsurf"]).sum())},
                "G_entries_reference": len(ka), "G_entries_fixed_point": len(kb),
                "pattern_entries_only_in_one": len(ka ^ kb), "rows_with_a_pattern_change": rows_changed,
                "common_entries_relative_value_difference": {"max": float(relv.max()), "p99": float(np.quantile(relv, 0.99)),
                                                             "median": float(np.median(relv)),
                                                             "fraction_above_1e-5": float((relv > 1e-5).mean())},
                "ray_steps_reference": int(ref["n_steps"]), "ray_steps_fixed_point": int(alt["n_steps"])})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
