"""ORACLE (test infrastructure, NOT product code): ctypes binding of oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_LIB_EXP = os.path.join(_HERE, "_build", "liboracle_experiments.so")
_lib = None
_lib_exp = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="F_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="F_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="F_CONTIGUOUS")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h"))]
    stale = force or not (os.path.exists(_LIB) and os.path.exists(_LIB_EXP)) or any(
        os.path.getmtime(s) > min(os.path.getmtime(_LIB), os.path.getmtime(_LIB_EXP)) for s in srcs + [os.path.join(_HERE, "Makefile")])
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB_EXP if os.environ.get("ORC_USE_EXPERIMENTS_LIB") == "1" else _LIB)
        _lib.orc_delsph.restype = C.c_float
        _lib.orc_delsph.argtypes = [C.c_float] * 4
    return _lib


def lib_exp():
    """The analysis experiments (fim_experiment.cpp) live in their own library so that nothing of them is inside the
    library the CPU baseline times."""
    global _lib_exp
    if _lib_exp is None:
        if not os.path.exists(_LIB_EXP):
            build()
        _lib_exp = C.CDLL(_LIB_EXP)
    return _lib_exp


class GBuildArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("vels", C.c_void_p),
        ("goxd", C.c_float), ("gozd", C.c_float), ("dvxd", C.c_float), ("dvzd", C.c_float),
        ("kmaxRc", C.c_int), ("tRc", C.c_void_p), ("depz", C.c_void_p), ("minthk", C.c_float),
        ("kmax", C.c_int), ("nsrc", C.c_int), ("nrcf", C.c_int),
        ("periods", C.c_void_p), ("nrc1", C.c_void_p), ("nsrcsurf1", C.c_void_p),
        ("scxf", C.c_void_p), ("sczf", C.c_void_p), ("rcxf", C.c_void_p), ("rczf", C.c_void_p),
        ("Gctrue", C.c_void_p), ("Gstrue", C.c_void_p),
        ("precomputed", C.c_int),
        ("pvRc", C.c_void_p), ("sen_vs", C.c_void_p), ("sen_vp", C.c_void_p), ("sen_rho", C.c_void_p),
        ("Lsen_Gsc", C.c_void_p),
        ("dsurf", C.c_void_p), ("obsTaa", C.c_void_p), ("tRcV", C.c_void_p),
        ("rw", C.c_void_p), ("iw_row", C.c_void_p), ("col", C.c_void_p),
        ("maxnar", C.c_long), ("nar", C.c_long),
        ("nthreads", C.c_int), ("rbint", C.c_int),
        ("t_kernels", C.c_double), ("t_dice_fmm", C.c_double), ("t_trace", C.c_double), ("t_assemble", C.c_double),
        ("n_accept", C.c_long), ("n_steps", C.c_long),
    ]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def surfdisp96(thk, vp, vs, rho, periods, iflsph=1):
    thk, vp, vs, rho = (np.ascontiguousarray(x, np.float32) for x in (thk, vp, vs, rho))
    t = np.ascontiguousarray(periods, np.float64)
    cg = np.zeros(len(t), np.float64)
    nev = C.c_long(0)
    st = lib().orc_surfdisp96(_p(thk), _p(vp), _p(vs), _p(rho), C.c_int(len(thk)), C.c_int(iflsph), C.c_int(2),
                              C.c_int(1), C.c_int(0), C.c_int(len(t)), _p(t), _p(cg), C.byref(nev))
    if st:
        raise RuntimeError(f"orc_surfdisp96 status {st}")
    return cg, nev.value


def refine_layer_mdl(minthk0, dep, vp, vs, rho):
    dep, vp, vs, rho = (np.ascontiguousarray(x, np.float32) for x in (dep, vp, vs, rho))
    out = [np.zeros(200, np.float32) for _ in range(5)]
    nsub = np.zeros(200, np.int32)
    rmax = C.c_int(0)
    lib().orc_refine_layer_mdl(C.c_float(minthk0), C.c_int(len(dep)), _p(dep), _p(vp), _p(vs), _p(rho),
                               C.byref(rmax), _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), _p(out[4]), _p(nsub))
    n = rmax.value
    return dict(rmax=n, rdep=out[0][:n], rvp=out[1][:n], rvs=out[2][:n], rrho=out[3][:n], rthk=out[4][:n],
                nsublay=nsub[:len(dep) - 1])


def brocher(vs):
    vp = C.c_float(0); rho = C.c_float(0)
    lib().orc_brocher(C.c_float(vs), C.byref(vp), C.byref(rho))
    return vp.value, rho.value


def tregn96(thk, TA, TC, TF, TL, TN, TRho, qp, qs, t_in, cp_in):
    n = len(thk)
    arrs = [np.ascontiguousarray(x, np.float32) for x in (thk, TA, TC, TF, TL, TN, TRho, qp, qs)]
    z = np.zeros(n, np.float32); o = np.ones(n, np.float32)
    t_in = np.ascontiguousarray(t_in, np.float32); cp_in = np.ascontiguousarray(cp_in, np.float32)
    d1 = np.zeros((60, 200), np.float32, order="F"); d2 = np.zeros_like(d1); d3 = np.zeros_like(d1)
    st = lib().orc_tregn96(C.c_int(n), *[_p(a) for a in arrs], _p(z), _p(z), _p(o), _p(o), C.c_int(len(t_in)),
                           _p(t_in), _p(cp_in), _p(d1), _p(d2), _p(d3))
    if st:
        raise RuntimeError(f"orc_tregn96 status {st}")
    k = len(t_in)
    return d1[:k, :n], d2[:k, :n], d3[:k, :n]


def depthkernel(vel, depz, tRc, minthk, nthreads=1):
    nx, ny, nz = vel.shape
    vel = np.asfortranarray(vel, np.float32); depz = np.ascontiguousarray(depz, np.float32)
    tRc = np.ascontiguousarray(tRc, np.float64); k = len(tRc)
    pv = np.zeros((nx * ny, k), np.float64, order="F")
    s = [np.zeros((nx * ny, k, nz), np.float64, order="F") for _ in range(3)]
    nev = C.c_long(0)
    st = lib().orc_depthkernel(C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(vel), _p(pv), _p(s[0]), _p(s[1]), _p(s[2]),
                               C.c_int(k), _p(tRc), _p(depz), C.c_float(minthk), C.c_int(nthreads), C.byref(nev))
    if st:
        raise RuntimeError(f"orc_depthkernel status {st}")
    return pv, s[0], s[1], s[2], nev.value   # pv, sen_vs, sen_vp, sen_rho


def depthkernel_ti(vel, depz, tRc, minthk, nthreads=1):
    nx, ny, nz = vel.shape
    vel = np.asfortranarray(vel, np.float32); depz = np.ascontiguousarray(depz, np.float32)
    tRc = np.ascontiguousarray(tRc, np.float64); k = len(tRc)
    pv = np.zeros((nx * ny, k), np.float64, order="F")
    L = np.zeros((nx * ny, k, nz - 1), np.float32, order="F")
    st = lib().orc_depthkernel_ti(C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(vel), _p(pv), C.c_int(k), _p(tRc),
                                  _p(depz), C.c_float(minthk), _p(L), C.c_int(nthreads))
    if st:
        raise RuntimeError(f"orc_depthkernel_ti status {st}")
    return pv, L


def fmm_source(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz):
    nnx = (nx - 3) * 5 + 1; nnz = (ny - 3) * 5 + 1
    pv = np.ascontiguousarray(pv, np.float64)
    veln = np.zeros((nnz, nnx), np.float32, order="F"); ttn = np.zeros_like(veln)
    nsts = np.zeros((nnz, nnx), np.int32, order="F")
    ttnr = np.zeros((129, 129), np.float32, order="F"); nstsr = np.full((129, 129), -1, np.int32, order="F")
    geom = np.zeros(8, np.int32); fgeom = np.zeros(8, np.float32)
    st = lib().orc_fmm_source(C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd),
                              C.c_float(dvzd), _p(pv), C.c_float(scx), C.c_float(scz), _p(veln), _p(ttn), _p(nsts),
                              _p(ttnr), _p(nstsr), _p(geom), _p(fgeom))
    if st:
        raise RuntimeError(f"orc_fmm_source status {st}")
    return dict(veln=veln, ttn=ttn, nsts=nsts, ttnr=ttnr, nstsr=nstsr, geom=geom, fgeom=fgeom)


def ray(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, rcx, rcz, azim=True):
    pv = np.ascontiguousarray(pv, np.float64)
    shp = (ny - 2 + 2, nx - 2 + 2)
    fdm = np.zeros(shp, np.float32, order="F"); fdmc = np.zeros_like(fdm); fdms = np.zeros_like(fdm)
    tt = C.c_float(0); ns = C.c_long(0)
    st = lib().orc_ray(C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd), C.c_float(dvzd),
                       _p(pv), C.c_float(scx), C.c_float(scz), C.c_float(rcx), C.c_float(rcz), C.c_int(int(azim)),
                       C.byref(tt), _p(fdm), _p(fdmc), _p(fdms), C.byref(ns))
    if st:
        raise RuntimeError(f"orc_ray status {st}")
    return tt.value, fdm, fdmc, fdms, ns.value


def gbuild(mode, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, gc=None, gs=None, tables=None,
           nthreads=1, maxnar=None, experiments=False):
    """Run the oracle orchestrator.  mode 0 forward, 1 iso G, 2 joint G.  Returns a dict.
    experiments=True runs the copy inside liboracle_experiments.so (the only build in which ORC_FIM_EXPERIMENT acts)."""
    nx, ny, nz = vels.shape
    vels = np.asfortranarray(vels, np.float32); depz = np.ascontiguousarray(depz, np.float32)
    tRc = np.ascontiguousarray(tRc, np.float64); k = len(tRc)
    dall = int(sv.dall)
    a = GBuildArgs()
    a.mode, a.nx, a.ny, a.nz = mode, nx, ny, nz
    a.vels = _p(vels)
    a.goxd, a.gozd, a.dvxd, a.dvzd = goxd, gozd, dvxd, dvzd
    a.kmaxRc = k; a.tRc = _p(tRc); a.depz = _p(depz); a.minthk = minthk
    a.kmax, a.nsrc, a.nrcf = sv.kmax, sv.nsrc, sv.nrcf
    keep = [np.asfortranarray(x) for x in (sv.periods, sv.nrc1, sv.nsrcsurf1, sv.scxf, sv.sczf, sv.rcxf, sv.rczf)]
    a.periods, a.nrc1, a.nsrcsurf1, a.scxf, a.sczf, a.rcxf, a.rczf = (_p(x) for x in keep)
    if gc is not None:
        gc = np.asfortranarray(gc, np.float32); gs = np.asfortranarray(gs, np.float32)
    a.Gctrue = _p(gc); a.Gstrue = _p(gs)
    if tables is None:
        tables = {}
        a.precomputed = 0
    else:
        a.precomputed = 1
    pv = tables.get("pvRc"); pv = np.zeros((nx * ny, k), np.float64, order="F") if pv is None else np.asfortranarray(pv, np.float64)
    sen = []
    for name in ("sen_vs", "sen_vp", "sen_rho"):
        t = tables.get(name)
        sen.append(np.zeros((nx * ny, k, nz), np.float64, order="F") if t is None else np.asfortranarray(t, np.float64))
    L = tables.get("Lsen_Gsc"); L = np.zeros((nx * ny, k, nz - 1), np.float32, order="F") if L is None else np.asfortranarray(L, np.float32)
    a.pvRc = _p(pv); a.sen_vs, a.sen_vp, a.sen_rho = (_p(s) for s in sen); a.Lsen_Gsc = _p(L)
    dsurf = np.zeros(dall, np.float32); taa = np.zeros(dall, np.float32)
    tRcV = np.zeros(((nx - 2) * (ny - 2), k), np.float64, order="F")
    a.dsurf = _p(dsurf); a.obsTaa = _p(taa); a.tRcV = _p(tRcV)
    if maxnar is None:
        maxnar = 0 if mode == 0 else max(1, int(dall) * 3 * 400)
    rw = np.zeros(maxnar, np.float32); iw = np.zeros(maxnar, np.int32); col = np.zeros(maxnar, np.int32)
    a.rw = _p(rw); a.iw_row = _p(iw); a.col = _p(col); a.maxnar = maxnar
    a.nthreads = nthreads
    st = (lib_exp() if experiments else lib()).orc_gbuild(C.byref(a))
    if st:
        raise RuntimeError(f"orc_gbuild status {st}")
    n = a.nar
    return dict(dsurf=dsurf, obsTaa=taa, tRcV=tRcV, pvRc=pv, sen_vs=sen[0], sen_vp=sen[1], sen_rho=sen[2],
                Lsen_Gsc=L, rw=rw[:n], row=iw[:n], col=col[:n], nar=n, rbint=a.rbint,
                times=dict(kernels=a.t_kernels, dice_fmm=a.t_dice_fmm, trace=a.t_trace, assemble=a.t_assemble),
                n_accept=a.n_accept, n_steps=a.n_steps)


class LsmrOut(C.Structure):
    _fields_ = [("istop", C.c_int), ("itn", C.c_int), ("normA", C.c_float), ("condA", C.c_float),
                ("normr", C.c_float), ("normAr", C.c_float), ("normx", C.c_float)]


def lsmr(m, n, row, col, rw, b, damp=0.0, atol=1e-5, btol=1e-4, conlim=200.0, itnlim=500, localSize=10):
    """LSMR (lsmrModule.f90:36) + aprod (aprod.f90:7) in single precision on a COO system (1-based indices)."""
    row = np.ascontiguousarray(row, np.int32); col = np.ascontiguousarray(col, np.int32)
    rw = np.ascontiguousarray(rw, np.float32); b = np.ascontiguousarray(b, np.float32)
    x = np.zeros(n, np.float32)
    out = LsmrOut()
    st = lib().orc_lsmr(C.c_int(m), C.c_int(n), C.c_longlong(len(rw)), _p(row), _p(col), _p(rw), _p(b), C.c_float(damp),
                        C.c_float(atol), C.c_float(btol), C.c_float(conlim), C.c_int(itnlim), C.c_int(localSize), _p(x),
                        C.byref(out))
    if st:
        raise RuntimeError(f"orc_lsmr status {st}")
    return x, {k: getattr(out, k) for k, _ in out._fields_}


# ---- outer inversion iteration: oracle/inversion.cpp ---------------------------------------------

def cal_ddat_sigma(obst, cbst):
    """CalDdatSigma (CalSigamNorm.f90:2-42) -> (sigmaT, meandeltaT)."""
    obst = np.ascontiguousarray(obst, np.float32); cbst = np.ascontiguousarray(cbst, np.float32)
    sig = np.zeros(len(obst), np.float32)
    mean = C.c_float(0)
    lib().orc_cal_ddat_sigma(C.c_int(len(obst)), _p(obst), _p(cbst), _p(sig), C.byref(mean))
    return sig, mean.value


def apply_weights(sigmaT, cbst, row, rw):
    """Main_Jt.f90:461-469 -> (datweight, weighted cbst, weighted rw); inputs are not modified."""
    sigmaT = np.ascontiguousarray(sigmaT, np.float32)
    cb = np.array(cbst, np.float32); w = np.zeros(len(sigmaT), np.float32)
    row = np.ascontiguousarray(row, np.int32); r = np.array(rw, np.float32)
    lib().orc_apply_weights(C.c_int(len(sigmaT)), _p(sigmaT), _p(w), _p(cb), C.c_long(len(r)), _p(row), _p(r))
    return w, cb, r


def tikhonov(nx, ny, nz, dall, iso_inv, weightGcs, weightVs, joint=False):
    """TikhonovRegularization (TikhRegul.f90:2) or, joint=True, TikhRegul_joint (:108): the appended triplets only.
    Returns dict(rw, row, col, count3, narVs) with narVs = entries of the dVs block (joint) or None."""
    maxvp = (nx - 2) * (ny - 2) * (nz - 1)
    cap = 7 * 3 * maxvp
    rw = np.zeros(cap, np.float32); row = np.zeros(cap, np.int32); col = np.zeros(cap, np.int32)
    nar = C.c_long(0); cnt = C.c_int(0); narvs = C.c_long(-1)
    if joint:
        lib().orc_tikh_joint(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(maxvp), C.c_int(dall), C.byref(nar), _p(rw),
                             _p(row), _p(col), C.byref(narvs), C.byref(cnt), C.c_float(weightGcs), C.c_float(weightVs))
    else:
        lib().orc_tikhonov(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(maxvp), C.c_int(dall), C.byref(nar), _p(rw),
                           _p(row), _p(col), C.byref(cnt), C.c_int(int(iso_inv)), C.c_float(weightGcs),
                           C.c_float(weightVs))
    n = nar.value
    return dict(rw=rw[:n], row=row[:n], col=col[:n], count3=cnt.value, narVs=narvs.value if joint else None)


def model_update(dv, vsf, iso_inv, minvel, maxvel):
    """Main_Jt.f90:582-620 -> (clipped dv, new vsf, gcf, gsf)."""
    nx, ny, nz = vsf.shape
    dv = np.array(dv, np.float32); v = np.array(vsf, np.float32, order="F")
    g = (nx - 2, ny - 2, nz - 1)
    gcf = np.zeros(g, np.float32, order="F"); gsf = np.zeros(g, np.float32, order="F")
    lib().orc_model_update(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(int(iso_inv)), _p(dv), _p(v),
                           C.c_float(minvel), C.c_float(maxvel), _p(gcf), _p(gsf))
    return dv, v, gcf, gsf


def model_norms(nar1, narVs, rw, col, dv, lameGcs, lameVs):
    rw = np.ascontiguousarray(rw, np.float32); col = np.ascontiguousarray(col, np.int32)
    dv = np.ascontiguousarray(dv, np.float32)
    out = np.zeros(6, np.float32)
    lib().orc_model_norms(C.c_long(nar1), C.c_long(len(rw)), C.c_long(-1 if narVs is None else narVs), _p(rw), _p(col),
                          _p(dv), C.c_float(lameGcs), C.c_float(lameVs), _p(out))
    return dict(zip(("VsNorm2", "VswNorm2", "GcsNorm2", "GcswNorm2", "Mnorm2", "MwNorm2"), map(float, out)))


def residuals(maxvp, nblk, rw, row, col, dv, datweight, Tdata):
    dall = len(Tdata)
    rw = np.ascontiguousarray(rw, np.float32); row = np.ascontiguousarray(row, np.int32)
    col = np.ascontiguousarray(col, np.int32); dv = np.ascontiguousarray(dv, np.float32)
    datweight = np.ascontiguousarray(datweight, np.float32); Tdata = np.ascontiguousarray(Tdata, np.float32)
    tvs = np.zeros(dall, np.float32); taa = np.zeros(dall, np.float32); res = np.zeros(dall, np.float32)
    nrm = np.zeros(2, np.float32)
    lib().orc_residuals(C.c_int(dall), C.c_long(maxvp), C.c_int(nblk), C.c_long(len(rw)), _p(rw), _p(row), _p(col),
                        _p(dv), _p(datweight), _p(Tdata), _p(tvs), _p(taa), _p(res), _p(nrm))
    return dict(fwdTvs=tvs, fwdTaa=taa, resbst=res, res2Nm=float(nrm[0]), resW2Nm=float(nrm[1]))


def res_stats(r):
    r = np.ascontiguousarray(r, np.float32)
    out = np.zeros(4, np.float32)
    lib().orc_res_stats(C.c_int(len(r)), _p(r), _p(out))
    return dict(meanabs=float(out[0]), std=float(out[1]), rms=float(out[2]), mean=float(out[3]))


def lsmr_controls(iso_inv, n):
    """Main_Jt.f90:542-554."""
    if iso_inv:
        return dict(atol=1e-3, btol=1e-3, conlim=1200.0, itnlim=1000, localSize=n // 4)
    return dict(atol=1e-5, btol=1e-4, conlim=200.0, itnlim=500, localSize=10)


def invert(vsf, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, obst, iso_mod, weightVs, weightGcs, damp, minvel,
           maxvel, maxiter, spfra=None, nthreads=1, log=None, on_iter=None):
    """The outer loop of DAzimSurfTomo (Main_Jt.f90:364-750) on the oracle's pieces.  Returns the final vsf, gcf, gsf,
    the per-iteration statistics and the last iteration's tables (for period_Azm_tomo.inv)."""
    nx, ny, nz = vsf.shape
    maxvp = (nx - 2) * (ny - 2) * (nz - 1)
    dall = int(sv.dall)
    vsf = np.array(vsf, np.float32, order="F")
    obst = np.ascontiguousarray(obst, np.float32)
    gcf = np.zeros((nx - 2, ny - 2, nz - 1), np.float32, order="F"); gsf = np.zeros_like(gcf)
    hist = []
    last = None
    for it in range(1, maxiter + 1):
        iso_inv = bool(iso_mod)
        n = maxvp if iso_inv else 3 * maxvp
        maxnar = None if spfra is None else int(np.float32(spfra) * dall * nx * ny * nz * 3)
        g = gbuild(1 if iso_inv else 2, vsf, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, nthreads=nthreads,
                   maxnar=maxnar)
        dsyn = g["dsurf"]
        cbst = (obst - dsyn).astype(np.float32)
        Tdata = cbst.copy()
        before = res_stats(cbst)
        sigmaT, meandeltaT = cal_ddat_sigma(obst, cbst)
        datweight, cbw, rww = apply_weights(sigmaT, cbst, g["row"], g["rw"])
        nar1 = int(g["nar"])
        tk = tikhonov(nx, ny, nz, dall, iso_inv, weightGcs, weightVs, joint=not iso_inv)
        rw_all = np.concatenate([rww, tk["rw"]]); row_all = np.concatenate([g["row"], tk["row"]])
        col_all = np.concatenate([g["col"], tk["col"]])
        m = dall + tk["count3"]
        b = np.concatenate([cbw, np.zeros(tk["count3"], np.float32)])
        dv, info = lsmr(m, n, row_all, col_all, rw_all, b, damp=damp, **lsmr_controls(iso_inv, n))
        dv, vsf, gc_new, gs_new = model_update(dv, vsf, iso_inv, minvel, maxvel)
        if not iso_inv:
            gcf, gsf = gc_new, gs_new
        mn = model_norms(nar1, None if iso_inv else nar1 + tk["narVs"], rw_all, col_all, dv, weightGcs, weightVs)
        rs = residuals(maxvp, 1 if iso_inv else 3, g["rw"], g["row"], g["col"], dv, datweight, Tdata)
        after = res_stats(rs["resbst"])
        rec = dict(iter=it, before=before, after=after, meandeltaT=meandeltaT, mean_weight=float(datweight.mean()),
                   lsmr=info, norms=mn, res2Nm=rs["res2Nm"], resW2Nm=rs["resW2Nm"], nar1=nar1, nar=len(rw_all),
                   dv_absmean=float(np.abs(dv[:maxvp]).mean()))
        hist.append(rec)
        last = dict(tRcV=g["tRcV"], Lsen_Gsc=g["Lsen_Gsc"], dsyn=dsyn, dv=dv, sigmaT=sigmaT, datweight=datweight,
                    resbst=rs["resbst"], fwdTvs=rs["fwdTvs"], fwdTaa=rs["fwdTaa"], Tdata=Tdata)
        if log:
            log("iter %d: before rms %.4f after rms %.4f itn %d istop %d |dv| %.5f" %
                (it, before["rms"], after["rms"], info["itn"], info["istop"], rec["dv_absmean"]))
        if on_iter:
            on_iter(it, vsf, gcf, gsf, rec)
    return dict(vsf=vsf, gcf=gcf, gsf=gsf, history=hist, last=last)


# ---- EXPERIMENT (oracle/fim_experiment.cpp): order-free fixed point of the reference's local eikonal solver ---------

def fmm_source_fim(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz):
    """Coarse travel-time field of one source with the coarse-grid continuation solved as a fast-iterative fixed point
    instead of the reference's heap march (refined source box and hand-off unchanged).  Returns (ttn (nnz,nnx), passes);
    passes < 0 means the overwrite phase hit its pass limit."""
    nnx = (nx - 3) * 5 + 1; nnz = (ny - 3) * 5 + 1
    pv = np.ascontiguousarray(pv, np.float64)
    ttn = np.zeros((nnz, nnx), np.float32, order="F")
    sw = C.c_long(0)
    st = lib_exp().orc_fmm_source_fim(C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd),
                                  C.c_float(dvzd), _p(pv), C.c_float(scx), C.c_float(scz), _p(ttn), C.byref(sw))
    if st:
        raise RuntimeError(f"orc_fmm_source_fim status {st}")
    return ttn, sw.value


class OrderStats(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("popped", "rule_mismatch", "pairs", "pair_ties", "pair_inversions",
                                        "sorted_exact_mismatch", "sorted_fim_mismatch", "sorted_fim_rank_errors",
                                        "dag_levels", "fim_passes", "fim_evals", "verify_order_flags", "verify_key_increase_flags", "key_increase_events", "harmless_tie_groups")]


def fmm_order_stats(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, refined=False, prefix=0):
    """EXPERIMENT (oracle/fim_experiment.cpp, ORDER EXPERIMENT): statistics of the coarse march of one source --
    is the final value a local function of the acceptance order (rule_mismatch == 0), do sorted arrival times predict
    that order (pair_ties / pair_inversions, sorted_*_mismatch), how deep is the dependency graph (dag_levels)."""
    pv = np.ascontiguousarray(pv, np.float64)
    s = OrderStats()
    head = (C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd), C.c_float(dvzd), _p(pv),
            C.c_float(scx), C.c_float(scz))
    if refined:
        st = lib_exp().orc_fmm_order_stats_refined(*head, C.c_int(prefix), C.byref(s))
    else:
        st = lib_exp().orc_fmm_order_stats(*head, C.c_int(prefix), C.byref(s))
    if st:
        raise RuntimeError(f"orc_fmm_order_stats status {st}")
    return {k: getattr(s, k) for k, _ in s._fields_}


def fmm_rank_iteration(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, guess, prefix=64, max_rounds=200):
    """EXPERIMENT: ranks -> replay -> sort iterated from the sorted order of `guess` (a (nnz,nnx) field) on the coarse
    march of one source; returns dict(rounds, mismatch vs the reference, flags of the local checks, popped)."""
    pv = np.ascontiguousarray(pv, np.float64)
    g = np.asfortranarray(guess, np.float32)
    out = [C.c_long(0) for _ in range(4)]
    st = lib_exp().orc_fmm_rank_iteration(C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd),
                                      C.c_float(dvzd), _p(pv), C.c_float(scx), C.c_float(scz), C.c_int(prefix), _p(g),
                                      C.c_int(max_rounds), *[C.byref(o) for o in out])
    if st:
        raise RuntimeError(f"orc_fmm_rank_iteration status {st}")
    return dict(zip(("rounds", "mismatch", "flags", "popped"), [o.value for o in out]))
