#!/usr/bin/env python
"""EXPERIMENT (CPU only; oracle/fim_experiment.cpp, ORDER EXPERIMENT): can the eikonal stage be made parallel AND stay
bit-identical to the reference's heap march?

In the reference the value a node ends with depends on the acceptance ORDER only: it is the quadrant solver evaluated
when the last of its direct neighbours was accepted before its own pop, with exactly the nodes accepted up to then alive.
Given a rank per node, every value is therefore a LOCAL function of already-final neighbours: a solve can be replayed in
dependency wavefronts.  This script measures, per solve,
  * rule_mismatch          nodes where that rule, fed with the reference's own order, differs from the reference (0 = the
                           rule is the reference's semantics, bit for bit);
  * replay exact           whether the replay fed with PREDICTED ranks -- the sorted order of the order-free fixed-point
                           (fast-iterative) values -- reproduces the reference's field bit for bit;
  * flagged                whether local, order-free checks on that replay (strict order of interacting pairs; key-increase
                           windows) raise a hazard; "wrong but not flagged" must be 0 for the checks to be sound;
  * dag_levels             length of the longest dependency chain = number of wavefronts a parallel replay needs.

    python scripts/order_experiment.py [--s200]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dazimsurftomo_b200 import formats as fm, synthetic   # noqa: E402
from oracle import pyoracle as po                          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "test1")


def summarize(name, rs):
    flagged = lambda r: r["verify_order_flags"] + r["verify_key_increase_flags"] > 0
    rng = lambda k: [min(r[k] for r in rs), max(r[k] for r in rs)]
    print(json.dumps(dict(
        case=name, solves=len(rs), nodes_per_solve=int(np.mean([r["popped"] for r in rs])),
        rule_mismatch_total=sum(r["rule_mismatch"] for r in rs),
        replay_from_fixed_point_ranks_exact=sum(1 for r in rs if r["sorted_fim_mismatch"] == 0),
        flagged_by_local_checks=sum(1 for r in rs if flagged(r)),
        wrong_but_not_flagged=sum(1 for r in rs if r["sorted_fim_mismatch"] > 0 and not flagged(r)),
        exact_and_verified=sum(1 for r in rs if r["sorted_fim_mismatch"] == 0 and not flagged(r)),
        dag_levels=rng("dag_levels"), interacting_pair_ties=rng("pair_ties"),
        interacting_pair_inversions=rng("pair_inversions"), key_increase_events=rng("key_increase_events"),
        mismatching_nodes_when_wrong=[r["sorted_fim_mismatch"] for r in rs if r["sorted_fim_mismatch"] > 0][:8])), flush=True)


def main():
    p = fm.read_para_forward(os.path.join(GOLD, "para.in"))
    depz, vs = fm.read_model(os.path.join(GOLD, "MODVs.true"), p.nx, p.ny, p.nz)
    sv = fm.read_surfdata(os.path.join(GOLD, "surfdata_subset.dat"), p.kmaxRc)
    pv, _ = po.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=8)
    rs = [po.fmm_order_stats(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv[:, k], float(sv.scxf[s, 0]), float(sv.sczf[s, 0]))
          for k in (0, 7, 20, 35) for s in range(int(sv.nsrcsurf1[0]))]
    summarize("test1 model (71 x 71 nodes), 4 periods x 5 sources", rs)
    # the key-increase hazards of the coarse march sit where the refined box hands over (band nodes whose injected keys
    # get overwritten): taking the first 64 accepts from the serial march removes them
    rs = [po.fmm_order_stats(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv[:, k], float(sv.scxf[s, 0]), float(sv.sczf[s, 0]),
                             prefix=64) for k in (0, 7, 20, 35) for s in range(int(sv.nsrcsurf1[0]))]
    summarize("test1 model (71 x 71 nodes), serial prefix of 64 accepts", rs)
    # the refined source box (129 x 129, stopping rule and close-node hand-off included); the first 4 accepts -- the
    # corners of the source cell, whose analytic keys get overwritten -- are taken from the serial march
    for pre in (0, 4):
        rs = [po.fmm_order_stats(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv[:, k], float(sv.scxf[s, 0]), float(sv.sczf[s, 0]),
                                 refined=True, prefix=pre) for k in (0, 7, 20, 35) for s in range(int(sv.nsrcsurf1[0]))]
        summarize("test1 model, REFINED source box (<= 129 x 129 nodes), serial prefix of %d accepts" % pre, rs)
    w = synthetic.yunnan_shaped(nsta=40); tb = synthetic.proxy_tables(w)
    rs = [po.fmm_order_stats(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, tb["pvRc"][:, (3 * s) % 36], float(w.sv.scxf[s, 0]),
                             float(w.sv.sczf[s, 0])) for s in range(24)]
    summarize("Yunnan-shaped model (176 x 196 nodes), 24 solves", rs)
    rs = [po.fmm_order_stats(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, tb["pvRc"][:, (3 * s) % 36], float(w.sv.scxf[s, 0]),
                             float(w.sv.sczf[s, 0]), prefix=64) for s in range(24)]
    summarize("Yunnan-shaped model (176 x 196 nodes), serial prefix of 64 accepts", rs)
    rs = [po.fmm_order_stats(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, tb["pvRc"][:, (3 * s) % 36], float(w.sv.scxf[s, 0]),
                             float(w.sv.sczf[s, 0]), refined=True, prefix=4) for s in range(24)]
    summarize("Yunnan-shaped model, REFINED source box, serial prefix of 4 accepts", rs)
    if "--s200" in sys.argv:
        w = synthetic.s200(src_per_period=4); tb = synthetic.proxy_tables(w)
        rs = [po.fmm_order_stats(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, tb["pvRc"][:, s], float(w.sv.scxf[s, 0]),
                                 float(w.sv.sczf[s, 0])) for s in range(4)]
        summarize("S200 model (996 x 996 nodes), 4 solves", rs)
        rs = [po.fmm_order_stats(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, tb["pvRc"][:, s], float(w.sv.scxf[s, 0]),
                                 float(w.sv.sczf[s, 0]), prefix=64) for s in range(4)]
        summarize("S200 model (996 x 996 nodes), serial prefix of 64 accepts", rs)


if __name__ == "__main__":
    main()
