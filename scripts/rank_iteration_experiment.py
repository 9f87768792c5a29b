#!/usr/bin/env python
"""RANK-ITERATION EXPERIMENT (CPU; oracle/fim_experiment.cpp::orc_fmm_rank_iteration): the cheapest way found to make the
eikonal stage parallel and bit-identical to the reference's heap march.

Start from the order of a trivial guess (geometric distance from the source), then iterate
        ranks -> wavefront replay with the reference's arithmetic -> values -> sort -> ranks
until the ranks stop changing.  Each round is one quadrant-solver evaluation per node (parallel over 64-350 dependency
wavefronts) plus a sort; the heap march itself makes about 4 evaluations per node, serially.  Reports the rounds needed,
whether the fixed point is the reference's field bit for bit, and whether the local hazard checks pass it.

    python scripts/rank_iteration_experiment.py [--s200]      (test1 survey needs /root/reference; falls back to the fixture)
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dazimsurftomo_b200 import formats as fm, synthetic   # noqa: E402
from oracle import pyoracle as po                          # noqa: E402


def distance_guess(nx, ny, goxd, gozd, dvxd, dvzd, scx, scz):
    nnx = (nx - 3) * 5 + 1; nnz = (ny - 3) * 5 + 1
    gox = (90 - goxd) * np.pi / 180; goz = gozd * np.pi / 180
    x = gox + np.arange(nnx) * (dvxd * np.pi / 180 / 5); z = goz + np.arange(nnz) * (dvzd * np.pi / 180 / 5)
    X, Z = np.meshgrid(x, z)                     # (nnz, nnx)
    return np.hypot(X - scx, (Z - scz) * np.sin(X)).astype(np.float32)


def run(name, geo, jobs, pvs):
    def one(j):
        k, scx, scz = j
        return po.fmm_rank_iteration(*geo, pvs[:, k], scx, scz, distance_guess(*geo, scx, scz), prefix=64)
    t0 = time.time()
    with ThreadPoolExecutor(8) as ex:
        rs = list(ex.map(one, jobs))
    rounds = np.array([r["rounds"] for r in rs])
    print(json.dumps(dict(case=name, solves=len(rs), nodes_per_solve=int(np.mean([r["popped"] for r in rs])),
                          exact=int(sum(r["mismatch"] == 0 for r in rs)),
                          exact_and_verified=int(sum(r["mismatch"] == 0 and r["flags"] == 0 for r in rs)),
                          wrong_but_not_flagged=int(sum(r["mismatch"] > 0 and r["flags"] == 0 for r in rs)),
                          rounds=dict(min=int(rounds.min()), median=float(np.median(rounds)), mean=float(rounds.mean()),
                                      p95=float(np.quantile(rounds, 0.95)), max=int(rounds.max())),
                          cpu_s=round(time.time() - t0, 1))), flush=True)


def run_refined(name, geo, jobs, pvs):
    """Same iteration on the refined source box (serial prefix 4; prefix code 1004 selects the distance guess)."""
    def one(j):
        k, scx, scz = j
        return po.fmm_order_stats(*geo, pvs[:, k], scx, scz, refined=True, prefix=1004)
    t0 = time.time()
    with ThreadPoolExecutor(8) as ex:
        rs = list(ex.map(one, jobs))
    fl = lambda r: r["verify_order_flags"] + r["verify_key_increase_flags"] > 0
    rounds = np.array([r["fim_passes"] for r in rs])
    print(json.dumps(dict(case=name, solves=len(rs), nodes_per_solve=int(np.mean([r["popped"] for r in rs])),
                          rule_mismatch_total=int(sum(r["rule_mismatch"] for r in rs)),
                          exact=int(sum(r["sorted_fim_mismatch"] == 0 for r in rs)),
                          exact_and_verified=int(sum(r["sorted_fim_mismatch"] == 0 and not fl(r) for r in rs)),
                          wrong_but_not_flagged=int(sum(r["sorted_fim_mismatch"] > 0 and not fl(r) for r in rs)),
                          rounds=dict(min=int(rounds.min()), median=float(np.median(rounds)), mean=float(rounds.mean()),
                                      p95=float(np.quantile(rounds, 0.95)), max=int(rounds.max())),
                          cpu_s=round(time.time() - t0, 1))), flush=True)


def main():
    ref = "/root/reference/example/test1_syn_foward"
    gold = os.path.join(ROOT, "tests", "golden", "test1")
    base = ref if os.path.exists(ref) else gold
    p = fm.read_para_forward(os.path.join(base, "para.in"))
    depz, vs = fm.read_model(os.path.join(base, "MODVs.true"), p.nx, p.ny, p.nz)
    sv = fm.read_surfdata(os.path.join(base, p.datafile) if base == ref else os.path.join(gold, "surfdata_subset.dat"), p.kmaxRc)
    pv, _ = po.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=8)
    geo = (p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd)
    jobs = [(k, float(sv.scxf[s, k]), float(sv.sczf[s, k])) for k in range(0, 36, 3) for s in range(int(sv.nsrcsurf1[k]))]
    run("reference test1 survey (71 x 71 nodes), coarse march, serial prefix 64, distance guess", geo, jobs, pv)
    run_refined("same solves: refined source box (<= 129 x 129 nodes), serial prefix 4, distance guess", geo, jobs, pv)
    w = synthetic.yunnan_shaped(nsta=300); tb = synthetic.proxy_tables(w)
    geo = (w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd)
    jobs = [(k, float(w.sv.scxf[s, k]), float(w.sv.sczf[s, k])) for k in (0, 12, 24, 35) for s in range(0, 299, 3)]
    run("Yunnan-shaped survey (176 x 196 nodes), coarse march, serial prefix 64, distance guess", geo, jobs, tb["pvRc"])
    run_refined("same solves: refined source box, serial prefix 4, distance guess", geo, jobs, tb["pvRc"])
    if "--s200" in sys.argv:
        w = synthetic.s200(src_per_period=4); tb = synthetic.proxy_tables(w)
        geo = (w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd)
        jobs = [(s, float(w.sv.scxf[s, 0]), float(w.sv.sczf[s, 0])) for s in range(4)]
        run("S200 model (996 x 996 nodes), coarse march, serial prefix 64, distance guess", geo, jobs, tb["pvRc"])


if __name__ == "__main__":
    main()
