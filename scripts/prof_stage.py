"""Stage timings of the eikonal / ray / assembly kernels on a synthetic workload with proxy
depth-kernel tables (profiling helper; the bench uses the real tables)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

from dazimsurftomo_b200 import api, synthetic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=202)
    ap.add_argument("--nz", type=int, default=9)
    ap.add_argument("--src", type=int, default=125)
    ap.add_argument("--kmax", type=int, default=8)
    ap.add_argument("--mode", type=int, default=2)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    t0 = time.time()
    w = synthetic.s200(src_per_period=a.src, kmax=a.kmax, n=a.n, nz=a.nz)
    tb = synthetic.proxy_tables(w)
    print("workload", w.name, "solves", w.n_solves, "rays", w.n_rays, "coarse nodes", w.nodes_coarse, "gen %.1fs" % (time.time() - t0), flush=True)
    h = api.Handle(0)
    t0 = time.time()
    plan = api.Plan(a.mode, w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv, tb, w.gc, w.gs, handle=h)
    print("plan create %.2fs h2d %.1f MB" % (time.time() - t0, plan.h2d_bytes / 1e6), flush=True)
    for r in range(a.reps):
        t0 = time.time()
        tm = plan.run()
        wall = time.time() - t0
        print(json.dumps(dict(rep=r, wall_s=round(wall, 3), nnz=plan.nnz, **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()})), flush=True)
    nodes = tm["n_accept"]
    print("accepts/s %.3e  steps/s %.3e  rays/s(fmm+trace) %.3e  rows/s(all) %.3e" % (
        nodes / (tm["fmm_ms"] * 1e-3), tm["n_steps"] / (tm["trace_ms"] * 1e-3),
        w.n_rays / ((tm["dice_ms"] + tm["fmm_ms"] + tm["trace_ms"]) * 1e-3), w.n_rays / (tm["total_ms"] * 1e-3)))


if __name__ == "__main__":
    main()
