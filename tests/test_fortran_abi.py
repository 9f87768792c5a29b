"""The gfortran-ABI drop-in symbols (dazim_fortran.cu) called the way gfortran-compiled Main_Jt.f90 /
MainForward.f90 would call them: every argument by reference, Fortran column-major arrays, LOGICAL as a
4-byte integer.  Results must equal the C-ABI / Python-mirror path bit for bit."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(x, ct):
    return C.byref(ct(x))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _fortran_tables(test1):
    p = test1["para"]; sv = test1["sv"]
    vs = np.asfortranarray(test1["vs"], np.float32)
    keep = dict(vs=vs, depz=np.ascontiguousarray(test1["depz"], np.float32), tRc=np.ascontiguousarray(p.tRc, np.float64),
                periods=np.asfortranarray(sv.periods, np.int32), nrc1=np.asfortranarray(sv.nrc1, np.int32),
                nsrc1=np.ascontiguousarray(sv.nsrcsurf1, np.int32), scxf=np.asfortranarray(sv.scxf, np.float32),
                sczf=np.asfortranarray(sv.sczf, np.float32), rcxf=np.asfortranarray(sv.rcxf, np.float32),
                rczf=np.asfortranarray(sv.rczf, np.float32))
    return p, sv, keep


def call_calsurfganisojoint(test1):
    """calsurfganisojoint_ exactly as gfortran-compiled Main_Jt.f90:403-406 calls it; returns the caller-owned arrays."""
    from dazimsurftomo_b200 import api
    lib = api.load()
    p, sv, k = _fortran_tables(test1)
    nx, ny, nz = p.nx, p.ny, p.nz
    nparpi = (nx - 2) * (ny - 2) * (nz - 1)
    dall = int(sv.dall)
    maxnar = dall * 3 * 400
    iw = np.zeros(2 * maxnar + 1, np.int32); rw = np.zeros(maxnar, np.float32); col = np.zeros(maxnar, np.int32)
    dsurf = np.zeros(dall, np.float32)
    GVs = np.zeros((dall, nparpi), np.float32, order="F"); GGc = np.zeros_like(GVs); GGs = np.zeros_like(GVs)
    L = np.zeros((nx * ny, p.kmaxRc, nz - 1), np.float32, order="F")
    tRcV = np.zeros(((nx - 2) * (ny - 2), p.kmaxRc), np.float64, order="F")
    nar = C.c_int(0)
    lib.calsurfganisojoint_(_ref(nx, C.c_int), _ref(ny, C.c_int), _ref(nz, C.c_int), _ref(nparpi, C.c_int), _ptr(k["vs"]),
                            _ptr(iw), _ptr(rw), _ptr(col), _ptr(dsurf), _ptr(GVs), _ptr(GGc), _ptr(GGs), _ptr(L),
                            _ref(dall, C.c_int), _ref(10, C.c_int), _ptr(tRcV), _ref(p.goxd, C.c_float),
                            _ref(p.gozd, C.c_float), _ref(p.dvxd, C.c_float), _ref(p.dvzd, C.c_float),
                            _ref(p.kmaxRc, C.c_int), _ptr(k["tRc"]), _ptr(k["periods"]), _ptr(k["depz"]),
                            _ref(p.sublayers, C.c_float), _ptr(k["scxf"]), _ptr(k["sczf"]), _ptr(k["rcxf"]), _ptr(k["rczf"]),
                            _ptr(k["nrc1"]), _ptr(k["nsrc1"]), _ref(sv.kmax, C.c_int), _ref(sv.nsrc, C.c_int),
                            _ref(sv.nrcf, C.c_int), C.byref(nar), _ref(0, C.c_int))
    n = nar.value
    return dict(nar=np.array([n]), rw=rw[:n].copy(), col=col[:n].copy(), row=iw[1:n + 1].copy(), dsurf=dsurf, L=L, tRcV=tRcV,
                GVs=GVs, GGc=GGc, GGs=GGs)


def test_fwdobstraveltimecps_symbol(gpu, test1):
    lib = gpu.load()
    p, sv, k = _fortran_tables(test1)
    nx, ny, nz = p.nx, p.ny, p.nz
    nparpi = (nx - 2) * (ny - 2) * (nz - 1)
    gc = np.asfortranarray(test1["gc"], np.float32); gs = np.asfortranarray(test1["gs"], np.float32)
    dall = int(sv.dall)
    dsurf = np.zeros(dall, np.float32); taa = np.zeros(dall, np.float32)
    tRcV = np.zeros(((nx - 2) * (ny - 2), p.kmaxRc), np.float64, order="F")
    L = np.zeros((nx * ny, p.kmaxRc, nz - 1), np.float32, order="F")
    lib.fwdobstraveltimecps_(_ref(nx, C.c_int), _ref(ny, C.c_int), _ref(nz, C.c_int), _ref(nparpi, C.c_int), _ptr(k["vs"]),
                             _ptr(gc), _ptr(gs), _ptr(dsurf), _ptr(taa), _ref(dall, C.c_int), _ref(10, C.c_int), _ptr(tRcV),
                             _ptr(L), _ref(p.goxd, C.c_float), _ref(p.gozd, C.c_float), _ref(p.dvxd, C.c_float),
                             _ref(p.dvzd, C.c_float), _ref(p.kmaxRc, C.c_int), _ptr(k["tRc"]), _ptr(k["periods"]),
                             _ptr(k["depz"]), _ref(p.sublayers, C.c_float), _ptr(k["scxf"]), _ptr(k["sczf"]),
                             _ptr(k["rcxf"]), _ptr(k["rczf"]), _ptr(k["nrc1"]), _ptr(k["nsrc1"]), _ref(sv.kmax, C.c_int),
                             _ref(sv.nsrc, C.c_int), _ref(sv.nrcf, C.c_int), _ref(0, C.c_int))
    r = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, sv)
    assert np.array_equal(dsurf, r["dsurf"]) and np.array_equal(taa, r["obsTaa"])
    assert np.array_equal(L, r["Lsen_Gsc"]) and np.array_equal(tRcV, r["tRcV"])


def test_calsurfganisojoint_and_lsmr_symbols(gpu, test1):
    lib = gpu.load()
    p, sv, k = _fortran_tables(test1)
    nx, ny, nz = p.nx, p.ny, p.nz
    nparpi = (nx - 2) * (ny - 2) * (nz - 1)
    dall = int(sv.dall)
    ref = gpu.CalSurfGAnisoJoint(test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv)
    nar_ref = int(ref["nar"]); rw_ref = ref["rw"].copy(); row_ref = ref["row"].copy(); col_ref = ref["col"].copy()
    dsurf_ref = ref["dsurf"].copy()
    maxnar = nar_ref + 1000
    # the driver's layout (Main_Jt.f90:325-330,529-532): iw(1) reserved, rows at iw(2:nar+1), columns in col(:)
    iw = np.zeros(2 * maxnar + 1, np.int32); rw = np.zeros(maxnar, np.float32); col = np.zeros(maxnar, np.int32)
    dsurf = np.zeros(dall, np.float32)
    GVs = np.zeros((dall, nparpi), np.float32, order="F"); GGc = np.zeros_like(GVs); GGs = np.zeros_like(GVs)
    L = np.zeros((nx * ny, p.kmaxRc, nz - 1), np.float32, order="F")
    tRcV = np.zeros(((nx - 2) * (ny - 2), p.kmaxRc), np.float64, order="F")
    nar = C.c_int(0)
    lib.calsurfganisojoint_(_ref(nx, C.c_int), _ref(ny, C.c_int), _ref(nz, C.c_int), _ref(nparpi, C.c_int), _ptr(k["vs"]),
                            _ptr(iw), _ptr(rw), _ptr(col), _ptr(dsurf), _ptr(GVs), _ptr(GGc), _ptr(GGs), _ptr(L),
                            _ref(dall, C.c_int), _ref(10, C.c_int), _ptr(tRcV), _ref(p.goxd, C.c_float),
                            _ref(p.gozd, C.c_float), _ref(p.dvxd, C.c_float), _ref(p.dvzd, C.c_float),
                            _ref(p.kmaxRc, C.c_int), _ptr(k["tRc"]), _ptr(k["periods"]), _ptr(k["depz"]),
                            _ref(p.sublayers, C.c_float), _ptr(k["scxf"]), _ptr(k["sczf"]), _ptr(k["rcxf"]), _ptr(k["rczf"]),
                            _ptr(k["nrc1"]), _ptr(k["nsrc1"]), _ref(sv.kmax, C.c_int), _ref(sv.nsrc, C.c_int),
                            _ref(sv.nrcf, C.c_int), C.byref(nar), _ref(0, C.c_int))
    n = nar.value
    assert n == nar_ref
    assert np.array_equal(rw[:n], rw_ref) and np.array_equal(col[:n], col_ref) and np.array_equal(iw[1:n + 1], row_ref)
    assert np.array_equal(dsurf, dsurf_ref)
    # dense copies are rebuilt from the triplets
    blk = (col_ref - 1) // nparpi; cc = (col_ref - 1) % nparpi
    for b, G in enumerate((GVs, GGc, GGs)):
        sel = blk == b
        assert np.array_equal(G[row_ref[sel] - 1, cc[sel]], rw_ref[sel]) and np.count_nonzero(G) == np.count_nonzero(rw_ref[sel])
    # the driver then packs iw(1) = nar, iw(nar+2:2nar+1) = col and calls LSMR (Main_Jt.f90:529-532,562)
    iw[0] = n
    iw[1 + n:1 + 2 * n] = col[:n]
    m, ncol = dall, 3 * nparpi
    rng = np.random.default_rng(4)
    b = rng.standard_normal(m).astype(np.float32)
    bb = b.copy()
    x = np.zeros(ncol, np.float32)
    istop, itn = C.c_int(0), C.c_int(0)
    fl = [C.c_float(0) for _ in range(5)]
    lib.__lsmrmodule_MOD_lsmr(_ref(m, C.c_int), _ref(ncol, C.c_int), _ref(len(iw), C.c_int), _ref(len(rw), C.c_int), _ptr(iw),
                              _ptr(rw), _ptr(bb), _ref(0.5, C.c_float), _ref(1e-5, C.c_float), _ref(1e-4, C.c_float),
                              _ref(200.0, C.c_float), _ref(40, C.c_int), _ref(10, C.c_int), _ref(0, C.c_int), _ptr(x),
                              C.byref(istop), C.byref(itn), *[C.byref(f) for f in fl])
    x2, info = gpu.LSMR(m, ncol, row_ref, col_ref, rw_ref, b, damp=0.5, atol=1e-5, btol=1e-4, conlim=200.0, itnlim=40, localSize=10)
    assert istop.value == info["istop"] and itn.value == info["itn"]
    assert np.array_equal(x, x2) and fl[2].value == info["normr"]


def test_calddatsigma_and_tikhonov_symbols(gpu, oracle):
    """CalDdatSigma, TikhonovRegularization and TikhRegul_joint the way Main_Jt.f90:460,513,515 calls them: rows behind
    iw(1), nar advanced in place, LOGICAL iso_inv as a 4-byte integer."""
    lib = gpu.load()
    rng = np.random.default_rng(8)
    dall = 3001
    obst = rng.uniform(5, 90, dall).astype(np.float32); cbst = (rng.standard_normal(dall) * 0.5).astype(np.float32)
    sig = np.zeros(dall, np.float32); mean = C.c_float(0)
    lib.calddatsigma_(_ref(dall, C.c_int), _ptr(obst), _ptr(cbst), _ptr(sig), C.byref(mean))
    s2, m2 = gpu.CalDdatSigma(obst, cbst)
    assert np.array_equal(sig, s2) and mean.value == np.float32(m2)
    osig, omean = oracle.cal_ddat_sigma(obst, cbst)
    assert np.float32(omean) == np.float32(mean.value) and np.allclose(sig, osig, rtol=2.5e-7, atol=0)
    nx, ny, nz = 9, 8, 5
    maxvp = (nx - 2) * (ny - 2) * (nz - 1)
    nar0 = 57                                                  # entries of "G" already there
    cap = nar0 + 21 * maxvp
    for joint in (False, True):
        iw = np.zeros(2 * cap + 1, np.int32); rw = np.zeros(cap, np.float32); col = np.zeros(cap, np.int32)
        rw[:nar0] = 1.5; col[:nar0] = 3; iw[1:nar0 + 1] = 7
        nar = C.c_int(nar0); count3 = C.c_int(0); narvs = C.c_int(0)
        if joint:
            lib.tikhregul_joint_(_ref(nx, C.c_int), _ref(ny, C.c_int), _ref(nz, C.c_int), _ref(maxvp, C.c_int),
                                 _ref(dall, C.c_int), C.byref(nar), _ptr(rw), _ptr(iw), _ptr(col), C.byref(narvs),
                                 C.byref(count3), _ref(35.0, C.c_float), _ref(240.0, C.c_float))
        else:
            lib.tikhonovregularization_(_ref(nx, C.c_int), _ref(ny, C.c_int), _ref(nz, C.c_int), _ref(maxvp, C.c_int),
                                        _ref(dall, C.c_int), C.byref(nar), _ptr(rw), _ptr(iw), _ptr(col), C.byref(count3),
                                        _ref(1, C.c_int), _ref(35.0, C.c_float), _ref(240.0, C.c_float))
        o = oracle.tikhonov(nx, ny, nz, dall, True, 35.0, 240.0, joint=joint)
        n = nar.value
        assert n == nar0 + len(o["rw"]) and count3.value == o["count3"]
        assert np.all(rw[:nar0] == 1.5) and np.all(col[:nar0] == 3) and np.all(iw[1:nar0 + 1] == 7) and iw[0] == 0
        assert np.array_equal(rw[nar0:n], o["rw"]) and np.array_equal(col[nar0:n], o["col"])
        assert np.array_equal(iw[1 + nar0:1 + n], o["row"])
        if joint:
            assert narvs.value == nar0 + o["narVs"]
