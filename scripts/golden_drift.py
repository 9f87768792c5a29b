"""Quantify, period by period, how far the reference's shipped forward output
(example/test1_syn_foward/output/surfphase_forward_RV3th.dat) is from what the shipped sources +
inputs produce (the oracle = line-by-line restatement; pinned on period_Azm_tomo.real at all 36
periods).  Writes profiles/r2_golden_drift.json; DESIGN.md section 2.1 reads from it.

Run in the build container (needs /root/reference):  python scripts/golden_drift.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dazimsurftomo_b200 import formats as fm      # noqa: E402
from oracle import pyoracle as po                 # noqa: E402

REF = "/root/reference/example/test1_syn_foward"


def per_period_table(Tg, Ti, Ta, n):
    rows = []
    for k in range(len(Tg) // n):
        s = slice(k * n, (k + 1) * n)
        rel = Tg[s] / (Ti[s] + Ta[s]) - 1.0
        A = np.stack([Ti[s], Ta[s]], 1)
        x = np.linalg.lstsq(A, Tg[s], rcond=None)[0]
        fit = (A @ x - Tg[s]) / Tg[s]
        rows.append(dict(period_index=k + 1, mean_ratio=float(1 + rel.mean()), median_abs_rel=float(np.median(np.abs(rel))),
                         share_beyond_print=float((np.abs(rel) > 4.5e-6).mean()), max_abs_rel=float(np.abs(rel).max()),
                         lsq_scale_Tiso=float(x[0]), lsq_scale_Taa=float(x[1]), rms_rel_after_lsq=float(np.sqrt((fit ** 2).mean()))))
    return rows


def main():
    po.build(); po.lib()
    p = fm.read_para_forward(os.path.join(REF, "para.in"))
    depz, vs = fm.read_model(os.path.join(REF, "MODVs.true"), p.nx, p.ny, p.nz)
    gc = fm.read_gcgs(os.path.join(REF, "MODGc.true"), p.nx, p.ny, p.nz)
    gs = fm.read_gcgs(os.path.join(REF, "MODGs.true"), p.nx, p.ny, p.nz)
    sv = fm.read_surfdata(os.path.join(REF, p.datafile), p.kmaxRc)
    pv, L = po.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=8)
    tab = dict(pvRc=pv, Lsen_Gsc=L)
    r = po.gbuild(0, vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, gc, gs, tables=tab, nthreads=8)
    g = fm.read_surfphase_velocities(os.path.join(REF, "output", "surfphase_forward_RV3th.dat"))
    dist = fm.forward_velocities(sv, np.ones(sv.dall, np.float32)).astype(np.float64)
    Tg = dist / g
    Ti = r["dsurf"].astype(np.float64); Ta = r["obsTaa"].astype(np.float64)
    n = sv.dall // p.kmaxRc
    out = dict(file="example/test1_syn_foward/output/surfphase_forward_RV3th.dat", rays=int(sv.dall), rays_per_period=int(n),
               periods_s=[float(t) for t in p.tRc], table=per_period_table(Tg, Ti, Ta, n))

    # (i) sensitivity of the reference's eikonal scheme to 1 float32 ulp of the phase-velocity map
    rng = np.random.default_rng(1)
    chaos = []
    for nulp in (1, 4, 16):
        pv32 = pv.astype(np.float32)
        step = rng.integers(-nulp, nulp + 1, size=pv32.shape).astype(np.int32)
        pert = (pv32.view(np.int32) + step).view(np.float32).astype(np.float64)
        q = po.gbuild(0, vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, gc * 0, gs * 0,
                      tables=dict(pvRc=np.asfortranarray(pert), Lsen_Gsc=L), nthreads=8)
        rel = np.abs(q["dsurf"].astype(np.float64) / Ti - 1)
        chaos.append(dict(ulps=nulp, share_moved_beyond_1p5e5=float((rel > 1.5e-5).mean()), share_moved_beyond_1e3=float((rel > 1e-3).mean()),
                          max_rel=float(rel.max()), median_rel=float(np.median(rel))))
    out["ulp_sensitivity"] = chaos

    # (ii) azimuth / distance dependence of the drift at a few periods
    scx, scz, rcx, rcz = fm.survey_loop_coords(sv)[:4]
    lat1 = np.pi / 2 - scx; lat2 = np.pi / 2 - rcx; dlon = rcz - scz
    az = np.arctan2(np.sin(dlon) * np.cos(lat2), np.cos(lat1) * np.sin(lat2) - np.sin(lat1) * np.cos(lat2) * np.cos(dlon))
    azt = []
    for k in (0, 10, 15, 20, 25, 30, 35):
        s = slice(k * n, (k + 1) * n)
        rel = Tg[s] / (Ti[s] + Ta[s]) - 1
        A = np.stack([np.ones(n), np.cos(2 * az[s]), np.sin(2 * az[s]), dist[s] / dist[s].mean() - 1], 1)
        x = np.linalg.lstsq(A, rel, rcond=None)[0]
        azt.append(dict(period_s=float(p.tRc[k]), mean=float(x[0]), cos2psi=float(x[1]), sin2psi=float(x[2]), per_unit_distance=float(x[3]),
                        rms_left=float(np.std(rel - A @ x))))
    out["azimuth_distance_regression"] = azt

    # (iii) where in depth a model difference would have to sit: 1-D experiments on the background profile
    T = np.asarray(p.tRc, float)

    def c_of(vsl, dep):
        vp, rho = zip(*[po.brocher(v) for v in vsl])
        m = po.refine_layer_mdl(float(p.sublayers), dep, vp, vsl, rho)
        return po.surfdisp96(m["rthk"], m["rvp"], m["rvs"], m["rrho"], T)[0]
    base = c_of([3.2, 3.4, 3.8, 4.2], [0, 10, 35, 60])
    tgt = np.array([t["mean_ratio"] - 1 for t in out["table"]])
    exps = {}
    for name, (vsl, dep) in {"half-space Vs -1 %": ([3.2, 3.4, 3.8, 4.2 * 0.99], [0, 10, 35, 60]),
                            "third node Vs -1 %": ([3.2, 3.4, 3.8 * 0.99, 4.2], [0, 10, 35, 60]),
                            "extra node 75 km, 4.10 km/s": ([3.2, 3.4, 3.8, 4.2, 4.10], [0, 10, 35, 60, 75]),
                            "extra node 90 km, 4.00 km/s": ([3.2, 3.4, 3.8, 4.2, 4.0], [0, 10, 35, 60, 90])}.items():
        rr = base / c_of(vsl, dep) - 1
        sc = tgt[-1] / rr[-1]
        exps[name] = dict(scale_to_match_40s=float(sc), rms_misfit_of_curve=float(np.sqrt(((rr * sc - tgt) ** 2).mean())),
                          ratio_at_20s_over_40s=float(rr[15] / rr[35]))
    exps["golden"] = dict(ratio_at_20s_over_40s=float(tgt[15] / tgt[35]))
    out["depth_localisation"] = exps
    path = os.path.join(ROOT, "profiles", "r2_golden_drift.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    for t in out["table"]:
        print("%2d  mean %.6f  median|rel| %.1e  beyond print %.3f  max %.1e  a=%.6f b=%.3f  rms after %.1e" % (
            t["period_index"], t["mean_ratio"], t["median_abs_rel"], t["share_beyond_print"], t["max_abs_rel"], t["lsq_scale_Tiso"],
            t["lsq_scale_Taa"], t["rms_rel_after_lsq"]))
    print(json.dumps(out["ulp_sensitivity"])); print(json.dumps(out["depth_localisation"]))


if __name__ == "__main__":
    main()
