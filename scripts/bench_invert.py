#!/usr/bin/env python
"""Next-stage measurement: the whole outer iteration of DAzimSurfTomo through the file-based
driver (python -m dazimsurftomo_b200.invert) on a T1-shaped synthetic survey (BASELINE config 3/4 shape: 17x17x4
model, 36 periods, 120 sources per period, 261 360 rays).

Observations are made by this library's own forward path on a checkerboard model (+ Gc/Gs checkerboards in joint
mode), the start model is its 1-D average; para.in / MOD / the '#'-block data file are written in the reference's
formats and the driver is run on them.  Reports, per outer iteration: device ms of the depth kernels, the G build,
the iteration tail (weights, Tikhonov rows, LSMR, update, norms) and its row-scaling kernel against the HBM roofline
(8 B per non-zero), wall time, the residual history; beside it the oracle's (C++ restatement
of the reference) loop timed on all host cores for a bounded number of iterations."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PARA = """cccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccc
c INPUT PARAMETERS
cccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccc
surfphase_forward.dat                c: traveltime data file
{nx} {ny} {nz}                       c: nx ny nz
{goxd}  {gozd}                       c: goxd gozd
{dvxd} {dvzd}                        c: dvxd dvzd
{sub}                                c: number of sublayers
2.0 5.0                              c: minimum and maximum Vsv
{nsrc}                               c: max(sources, receivers)
0.2                                  c: sparsity fraction
{maxiter}                            c: maximum of iteration
{iso}                                c: iso-mode
cccccccc control parameters
{wvs}                                c: smoothing for dVsv
{wg}                                 c: smoothing for Gc,s
0                                    c: damping
cccccccccc periods
{kmax}                               c: kmaxRc
{periods}
"""


def real_example(tag, h, cpu_iters):
    """BASELINE configs 3 / 4 for real: the reference's example/test2_syn_iso_inv or test3_syn_joint_inv (fixture copies
    of para.in, MOD and the data file), all outer iterations, compared with the model the reference shipped."""
    from dazimsurftomo_b200 import formats as fm, invert
    from oracle import pyoracle as po
    inv = os.path.join(ROOT, "tests", "golden", "inv")
    tmp = tempfile.mkdtemp(prefix="dazim_%s_" % tag)
    fm.stage_reference_example(inv, tag, tmp)
    t0 = time.time()
    out = invert.run(os.path.join(tmp, "para.in"), handle=h, log_stream=open(os.devnull, "w"))
    wall = time.time() - t0
    hist = out["history"]; p = out["para"]; sv = out["survey"]; g = out["gpu_ms"]; n = len(hist)
    fix = np.load(os.path.join(inv, "%s_iter.npz" % tag))
    depz, vs0 = fm.read_model(os.path.join(tmp, "MOD"), p.nx, p.ny, p.nz)
    if tag == "test2":
        ours = np.array([float(l[24:32]) for l in open(os.path.join(tmp, "DSurfTomo.inv")).read().splitlines()])
        vs_ref = {"file": "plot_script/DSurfTomo.inv", "max_abs_dVs_km_s": float(np.abs(ours - fix["shipped"]).max()),
                  "rms_dVs_km_s": float(np.sqrt(((ours - fix["shipped"]) ** 2).mean())),
                  "model_moved_km_s": float(np.abs(fix["shipped"] - vs0.ravel(order="F")).max())}
    else:
        tab = np.loadtxt(os.path.join(tmp, "Gc_Gs_model.inv"))
        vs_ref = {"file": "plot_script/Gc_Gs_model.inv", "max_abs_dVs_mid_km_s": float(np.abs(tab[:, 3] - fix["shipped"][:, 0]).max()),
                  "max_abs_dGc_percent": float(np.abs(tab[:, 6] - fix["shipped"][:, 1]).max()),
                  "max_abs_dGs_percent": float(np.abs(tab[:, 7] - fix["shipped"][:, 2]).max()),
                  "shipped_max_Gc_percent": float(np.abs(fix["shipped"][:, 1]).max())}
    # against the ORACLE's own final model (scripts/pin_inversion.py, same number of outer iterations)
    of = fix["final"] if tag == "test2" else fix["final_vsf"]
    vs_orc = {"max_abs_dVs_km_s": float(np.abs(out["vsf"] - of).max())}
    if tag != "test2":
        vs_orc["max_abs_dGc_percent"] = float(np.abs(out["gcf"] - fix["final_gcf"]).max() * 100)
        vs_orc["max_abs_dGs_percent"] = float(np.abs(out["gsf"] - fix["final_gsf"]).max() * 100)
        vs_orc["rms_dGc_percent"] = float(np.sqrt(((out["gcf"] - fix["final_gcf"]) ** 2).mean()) * 100)
    obst = (sv.dist / sv.obsvel).astype(np.float32)
    cpu_s = None
    if cpu_iters > 0:                  # 0: skip (test4's CPU iteration takes minutes; it is timed off the GPU box)
        t0 = time.time()
        po.invert(vs0, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, obst, p.iso_mod, p.weightVs, p.weightGcs,
                  p.damp, p.minvel, p.maxvel, cpu_iters, spfra=p.spfra, nthreads=os.cpu_count() or 1)
        cpu_s = (time.time() - t0) / cpu_iters
    dev_s = (g["kernels"] + g["gbuild"] + g["iterate"]) * 1e-3
    print(json.dumps({
        "stage": "DAzimSurfTomo para.in on the reference's example/%s (BASELINE config %d), file-based GPU driver" %
                 {"test2": ("test2_syn_iso_inv", 3), "test3": ("test3_syn_joint_inv", 4), "test4": ("test4_Yunnan, real data", 5)}[tag],
        "rays": int(sv.dall), "outer_iterations": n, "nnz_G": int(np.mean([s["nar1"] for s in hist])),
        "vs_reference_shipped_model": vs_ref, "vs_oracle_final_model": vs_orc,
        "device_s_total": dev_s, "wall_s_total_incl_file_io": wall,
        "gpu_ms_per_iteration": {"depth_kernels": g["kernels"] / n, "g_build": g["gbuild"] / n, "iteration_tail": g["iterate"] / n,
                                 "of_which_lsmr": g["lsmr"] / n},
        "per_iteration_ms": {"tail": [round(s["step_ms"], 1) for s in hist], "lsmr_solve": [round(s["lsmr"]["solve_ms"], 1) for s in hist]},
        "lsmr_itn": [s["lsmr"]["itn"] for s in hist],
        "rms_before_first_after_last": [hist[0]["before"]["rms"], hist[-1]["after"]["rms"]],
        "cpu_port_s_per_iteration": cpu_s, "cpu_cores": os.cpu_count(),
        "speedup_device_vs_cpu_port_per_iteration": None if cpu_s is None else cpu_s / (dev_s / n)}))


def main():
    from dazimsurftomo_b200 import api, formats as fm, invert, synthetic
    from oracle import pyoracle as po
    mode = sys.argv[1] if len(sys.argv) > 1 else "iso"
    maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    cpu_iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    h = api.Handle(0)
    if mode in ("test2", "test3", "test4"):
        return real_example(mode, h, cpu_iters)
    iso = mode == "iso"
    w = synthetic.t1_shaped()
    # observations: T_iso (+ T_aa) of the true model through the forward path
    fwd = api.FwdObsTraveltimeCPS(w.vs, w.gc if not iso else np.zeros_like(w.gc), w.gs if not iso else np.zeros_like(w.gs),
                                  w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv, handle=h)
    tsyn = (fwd["dsurf"] + fwd["obsTaa"]).astype(np.float32)
    start = np.asfortranarray(np.broadcast_to(w.vs.mean(axis=(0, 1), keepdims=True), w.vs.shape).astype(np.float32))
    tmp = tempfile.mkdtemp(prefix="dazim_inv_")
    fm.write_surfphase_forward(os.path.join(tmp, "surfphase_forward.dat"), w.sv, tsyn)
    fm.write_model(os.path.join(tmp, "MOD"), w.depz, start)
    open(os.path.join(tmp, "para.in"), "w").write(PARA.format(
        nx=w.nx, ny=w.ny, nz=w.nz, goxd=w.goxd, gozd=w.gozd, dvxd=w.dvxd, dvzd=w.dvzd, sub=int(w.sublayers),
        nsrc=w.sv.nsrc + 1, maxiter=maxiter, iso="T" if iso else "F", wvs=240, wg=35, kmax=len(w.tRc),
        periods=" ".join("%g" % t for t in w.tRc)))
    t0 = time.time()
    out = invert.run(os.path.join(tmp, "para.in"), handle=h, log_stream=open(os.devnull, "w"))
    wall = time.time() - t0
    hist = out["history"]
    peak = 6551.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    nnz = int(np.mean([s["nar1"] for s in hist]))
    scale_ms = float(np.mean([s["scale_ms"] for s in hist]))
    _, start_r = fm.read_model(os.path.join(tmp, "MOD"), w.nx, w.ny, w.nz)       # the 3-decimal start model the driver read
    # the oracle's loop on the same files, all host cores, bounded
    p = out["para"]; sv = out["survey"]
    obst = (sv.dist / sv.obsvel).astype(np.float32)
    t0 = time.time()
    o = po.invert(start_r, out["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, obst, p.iso_mod, p.weightVs,
                  p.weightGcs, p.damp, p.minvel, p.maxvel, cpu_iters, spfra=p.spfra, nthreads=os.cpu_count() or 1)
    cpu_s = (time.time() - t0) / cpu_iters
    g = out["gpu_ms"]
    print(json.dumps({
        "stage": "outer iteration of DAzimSurfTomo (Main_Jt.f90:364-750), file-based driver", "mode": mode,
        "workload": w.name, "rays": int(w.sv.dall), "solves": w.n_solves, "outer_iterations": maxiter, "nnz_G": nnz,
        "gpu_ms_per_iteration": {"depth_kernels": g["kernels"] / maxiter, "g_build": g["gbuild"] / maxiter,
                                 "iteration_tail": g["iterate"] / maxiter, "of_which_lsmr": g["lsmr"] / maxiter,
                                 "of_which_row_scaling": scale_ms},
        "row_scaling_GBps": 8.0 * nnz / (scale_ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
        "row_scaling_frac_of_hbm": 8.0 * nnz / (scale_ms * 1e-3) / 1e9 / peak,
        "wall_s_per_iteration_incl_file_io": wall / maxiter,
        "lsmr_itn": [s["lsmr"]["itn"] for s in hist], "lsmr_istop": [s["lsmr"]["istop"] for s in hist],
        "rms_before": [round(s["before"]["rms"], 4) for s in hist], "rms_after": [round(s["after"]["rms"], 4) for s in hist],
        "per_iteration_ms": {"tail": [round(s["step_ms"], 2) for s in hist],
                             "lsmr_setup": [round(s["lsmr"]["setup_ms"], 2) for s in hist],
                             "lsmr_solve": [round(s["lsmr"]["solve_ms"], 2) for s in hist]},
        "oracle_first_iteration": {"rms_before": o["history"][0]["before"]["rms"], "rms_after": o["history"][0]["after"]["rms"],
                                   "lsmr_itn": o["history"][0]["lsmr"]["itn"]},
        "cpu_port_s_per_iteration": cpu_s, "cpu_cores": os.cpu_count(),
        "speedup_device_vs_cpu_port": cpu_s / ((g["kernels"] + g["gbuild"] + g["iterate"]) / maxiter * 1e-3)}))


if __name__ == "__main__":
    main()
