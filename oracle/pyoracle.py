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
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="F_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="F_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="F_CONTIGUOUS")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h"))]
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.orc_delsph.restype = C.c_float
        _lib.orc_delsph.argtypes = [C.c_float] * 4
    return _lib


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
           nthreads=1, maxnar=None):
    """Run the oracle orchestrator.  mode 0 forward, 1 iso G, 2 joint G.  Returns a dict."""
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
    st = lib().orc_gbuild(C.byref(a))
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
