"""CUDA path (through the C ABI) vs the oracle on the same inputs.  Bit-exact for travel-time
fields, ray footprints, sparsity pattern and G values (float32 arithmetic is reproduced
operation by operation); tolerances are written where floating-point libraries differ."""
import numpy as np
import os
import pytest

pytestmark = pytest.mark.gpu


def _sources(test1, n=6):
    sv = test1["sv"]
    out = []
    for k in range(sv.kmax):
        for s in range(int(sv.nsrcsurf1[k])):
            out.append((float(sv.scxf[s, k]), float(sv.sczf[s, k])))
    # add a source hugging the model corner (refined box clipped on two sides)
    return out[:n]


def test_fmm_fields_bit_exact(gpu, oracle, test1, test1_tables):
    p = test1["para"]
    pv = np.ascontiguousarray(test1_tables["pvRc"][:, 7])
    src = _sources(test1, 8)
    # corner / edge sources exercise the clipped source box and the literal exit rule
    g0x = np.float32((90.0 - p.goxd) * np.pi / 180); g0z = np.float32(p.gozd * np.pi / 180)
    dv = np.float32(p.dvxd * np.pi / 180)
    src += [(float(g0x + dv * 0.3), float(g0z + dv * 0.4)), (float(g0x + dv * 13.9), float(g0z + dv * 7.2)),
            (float(g0x + dv * 6.0), float(g0z + dv * 13.95))]
    r = gpu.fmm_solve(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, [s[0] for s in src], [s[1] for s in src])
    for i, (x, z) in enumerate(src):
        o = oracle.fmm_source(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, x, z)
        assert np.array_equal(r["veln"], o["veln"])
        nzr, nxr = o["geom"][0], o["geom"][1]
        assert tuple(r["geom"][:6, i]) == tuple(o["geom"][:6])
        assert np.array_equal(r["nstsr"][:nzr, :nxr, i] == 0, o["nstsr"][:nzr, :nxr] == 0)
        alive = o["nstsr"][:nzr, :nxr] >= 0
        assert np.array_equal(r["ttnr"][:nzr, :nxr, i][alive], o["ttnr"][:nzr, :nxr][alive])
        assert np.array_equal(r["nstsr"][:nzr, :nxr, i], o["nstsr"][:nzr, :nxr])     # heap slots too
        assert np.array_equal(r["ttn"][:, :, i], o["ttn"]), i
        assert np.all(r["nsts"][:, :, i] == 0)


def test_ray_footprints_bit_exact(gpu, oracle, test1, test1_tables):
    p = test1["para"]; sv = test1["sv"]
    pv = np.ascontiguousarray(test1_tables["pvRc"][:, 0])
    scx, scz, rcx, rcz = [], [], [], []
    for s in range(3):
        for r in range(0, int(sv.nrc1[s, 0]), 5):
            scx.append(sv.scxf[s, 0]); scz.append(sv.sczf[s, 0]); rcx.append(sv.rcxf[r, s, 0]); rcz.append(sv.rczf[r, s, 0])
    # a receiver within two steps of the source (no path) and one inside the source box
    scx += [scx[0], scx[0]]; scz += [scz[0], scz[0]]
    rcx += [np.float32(scx[0] + 1e-5), np.float32(scx[0] + 3e-3)]; rcz += [np.float32(scz[0] + 1e-5), np.float32(scz[0] - 2e-3)]
    for azim in (True, False):
        tt, fdm, fdmc, fdms = gpu.raytrace(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, scx, scz, rcx, rcz, azim=azim)
        for i in range(len(scx)):
            ot, of, oc, os_, ns = oracle.ray(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, scx[i], scz[i], rcx[i], rcz[i], azim)
            assert tt[i] == np.float32(ot), (i, tt[i], ot)
            assert np.array_equal(fdm[:, :, i], of), i
            if azim:
                # cos/sin(2 psi) come from double-precision libm on both sides, rounded to float32
                assert np.array_equal(fdmc[:, :, i], oc), i
                assert np.array_equal(fdms[:, :, i], os_), i


from conftest import note as _note


def _cmp_coo(r, o):
    assert r["nar"] == o["nar"] > 0
    assert np.array_equal(r["row"], o["row"])          # bit-exact sparsity pattern
    assert np.array_equal(r["col"], o["col"])
    assert np.array_equal(r["rw"], o["rw"])            # and values (same float32/float64 op order)
    assert np.array_equal(r["dsurf"], o["dsurf"])


def test_forward_subset(gpu, oracle, test1, test1_tables):
    from dazimsurftomo_b200 import formats as fm
    p = test1["para"]
    r = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, test1["sv"], tables=test1_tables)
    o = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"],
                      test1["gc"], test1["gs"], tables=test1_tables)
    assert np.array_equal(r["dsurf"], o["dsurf"])
    assert np.array_equal(r["obsTaa"], o["obsTaa"])
    c = fm.forward_velocities(test1["sv"], r["dsurf"] + r["obsTaa"])
    assert np.abs(c - test1["gold_c"]).max() < 1.5e-5        # the reference's own output, f9.5
    assert r["times"]["n_launch"] >= 3 and r["times"]["n_accept"] == o["n_accept"]
    assert r["times"]["n_steps"] == o["n_steps"]


def test_gmatrix_iso_and_joint(gpu, oracle, test1, test1_tables):
    p = test1["para"]
    pv, svs, svp, srho, _ = oracle.depthkernel(test1["vs"], test1["depz"], p.tRc, p.sublayers, nthreads=8)
    tb = dict(test1_tables, sen_vs=svs, sen_vp=svp, sen_rho=srho)
    args = (test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"])
    _cmp_coo(gpu.CalSurfG(*args, tables=tb), oracle.gbuild(1, *args, tables=tb))
    _cmp_coo(gpu.CalSurfGAnisoJoint(*args, tables=tb), oracle.gbuild(2, *args, tables=tb))


def test_batched_equals_single(gpu, test1, test1_tables, monkeypatch):
    """Source batching (workspace reuse) must not change a bit."""
    p = test1["para"]
    args = (test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"])
    a = gpu.FwdObsTraveltimeCPS(*args, tables=test1_tables)
    monkeypatch.setenv("DAZIM_BATCH", "3")
    b = gpu.FwdObsTraveltimeCPS(*args, tables=test1_tables)
    assert np.array_equal(a["dsurf"], b["dsurf"]) and np.array_equal(a["obsTaa"], b["obsTaa"])
    assert b["times"]["n_fmm_launch"] == 7


def test_heap_spill_path(gpu, test1, test1_tables, monkeypatch):
    """Tiny shared-memory heap: most of the narrow band lives in the global spill array."""
    p = test1["para"]
    args = (test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"])
    a = gpu.FwdObsTraveltimeCPS(*args, tables=test1_tables)
    monkeypatch.setenv("DAZIM_HCAP", "64")
    b = gpu.FwdObsTraveltimeCPS(*args, tables=test1_tables)
    assert np.array_equal(a["dsurf"], b["dsurf"]) and np.array_equal(a["obsTaa"], b["obsTaa"])


def test_errors_mirror_reference_stops(gpu, test1, test1_tables):
    import copy
    p = test1["para"]
    sv = copy.deepcopy(test1["sv"])
    sv.scxf = sv.scxf.copy(); sv.scxf[0, 0] = np.float32(0.2)           # far outside the model
    with pytest.raises(gpu.DazimError) as e:
        gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, sv, tables=test1_tables)
    assert e.value.code == 1
    sv = copy.deepcopy(test1["sv"])
    sv.rcxf = sv.rcxf.copy(); sv.rcxf[0, 0, 0] = np.float32(0.2)
    with pytest.raises(gpu.DazimError) as e:
        gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, sv, tables=test1_tables)
    assert e.value.code == 2


# ---------------------------------------------------------------------------------------------
# Thomson-Haskell stage (K1 root search, K2 eigenfunction partials)
def test_surfdisp96_kat_and_random_profiles(gpu, oracle):
    rng = np.random.default_rng(3)
    T = np.arange(5, 41, dtype=float)
    profs = []
    vs = [3.2, 3.4, 3.8, 4.2]; dep = [0, 10, 35, 60]
    vp, rho = zip(*[oracle.brocher(v) for v in vs])
    profs.append(oracle.refine_layer_mdl(2.0, dep, vp, vs, rho))
    for _ in range(63):
        v = np.sort(rng.uniform(2.6, 4.6, 4)).astype(np.float32)
        if rng.random() < 0.3:
            v[1], v[2] = v[2], v[1]          # low-velocity zone
        vp, rho = zip(*[oracle.brocher(float(x)) for x in v])
        profs.append(oracle.refine_layer_mdl(2.0, dep, vp, v, rho))
    nl = profs[0]["rmax"]
    arr = {k: np.asfortranarray(np.stack([p[k] for p in profs], 1).astype(np.float32)) for k in ("rthk", "rvp", "rvs", "rrho")}
    cg = gpu.surfdisp96(arr["rthk"], arr["rvp"], arr["rvs"], arr["rrho"], T)
    ref = np.stack([oracle.surfdisp96(p["rthk"], p["rvp"], p["rvs"], p["rrho"], T)[0] for p in profs], 1)
    assert cg.shape == ref.shape == (36, 64)
    kat = [3.04704, 3.06911, 3.09024, 3.11063, 3.13054, 3.15021, 3.16979, 3.18940]
    assert np.abs(cg[:8, 0] - kat).max() < 5.1e-6
    # float32-rounded roots: identical except where CUDA/glibc exp,sin,cos differ in the last ulp
    # exactly at a rounding boundary (tolerance north_star: 1e-5 relative)
    assert np.abs(cg - ref).max() <= 1e-5 * ref.max()
    assert (cg != ref).mean() < 0.01


def test_depthkernel_ti_vs_oracle(gpu, oracle, test1):
    """Lsen_Gsc (depthkernelTI.f90:96-106) = a float32 sum over sub-layers of double-precision tregn96 partials.  With
    identical roots the only difference is libm's last-ulp behaviour inside the complex propagator chain; entries are
    compared RELATIVE to their own size (north star: 1e-5), and the ones outside are counted, printed and shown to be
    cancellation cases: |entry| small against the layer's largest kernel."""
    p = test1["para"]
    pv, L = gpu.depthkernelTI(test1["vs"], test1["depz"], p.tRc, p.sublayers)
    opv, oL = oracle.depthkernel_ti(test1["vs"], test1["depz"], p.tRc, p.sublayers, nthreads=8)
    assert np.array_equal(pv, opv), int((pv != opv).sum())       # 289 x 36 float32-rounded roots
    scale = np.abs(oL).max()
    d = np.abs(L.astype(np.float64) - oL.astype(np.float64))
    assert d.max() <= 1e-5 * scale                                   # absolute, against the largest kernel
    rel = d / np.maximum(np.abs(oL), 1e-30)
    out = rel > 1e-5
    ulps = d / np.spacing(np.abs(oL).astype(np.float32)).astype(np.float64)
    _note("depthkernelTI Lsen_Gsc: %d of %d entries differ, %d beyond 1e-5 relative (largest %.2e, at |entry| <= %.2e x max); "
          "max difference %.1f float32 ulps of the entry" % (int((d > 0).sum()), d.size, int(out.sum()), rel.max(),
                                                             (np.abs(oL)[out].max() / scale) if out.any() else 0.0, ulps.max()))
    assert out.sum() <= 0.002 * d.size
    if out.any():
        assert np.abs(oL)[out].max() <= 0.05 * scale                 # only where the sub-layer terms cancel
    assert rel.max() < 2e-4 and np.median(rel) < 1e-6
    # and against the reference's own period_Azm_tomo.real (col 4)
    g = test1["azm"]
    mine = [pv[jj * p.nx + ii, tt] for tt in range(p.kmaxRc) for jj in range(1, p.ny - 1) for ii in range(1, p.nx - 1)]
    assert np.abs(np.array(mine) - g[:, 3]).max() < 6e-6


def test_depthkernel_fd_vs_oracle(gpu, oracle, test1):
    """sen_* = (cg2 - cg1) / (0.01 * par) with float32-rounded roots cg (CalSurfG.f90:60-120).  The roots are
    identical except where CUDA's and glibc's exp/sin/cos differ in the last ulp of a double exactly at a float32
    rounding boundary: such a 1-ulp flip of ONE root moves the entry by ulp32(c) / (0.01 par) ~ 7e-6 km/s per unit,
    which is > 1e-5 RELATIVE for the small kernels.  Every entry outside 1e-5 relative is counted, printed, and tied
    to an integer number (1 or 2) of float32 ulps of c; everything else must be inside 1e-5 relative."""
    p = test1["para"]
    vs = np.asfortranarray(test1["vs"][3:9, 4:8, :])          # 24 nodes x 25 variants x 36 periods
    pv, s1, s2, s3 = gpu.depthkernel(vs, test1["depz"], p.tRc, p.sublayers)
    opv, o1, o2, o3, _ = oracle.depthkernel(vs, test1["depz"], p.tRc, p.sublayers, nthreads=8)
    assert np.array_equal(pv, opv)
    nxy, kmax, nz = o1.shape
    vsn = vs.reshape(nxy, nz, order="F").astype(np.float64)
    vpn = np.array([[oracle.brocher(float(v))[0] for v in row] for row in vsn])
    rhon = np.array([[oracle.brocher(float(v))[1] for v in row] for row in vsn])
    assert 2.0 <= opv.min() and opv.max() < 8.0
    ulp = np.full(opv.shape, 2.0 ** -22)            # float32 ulp of c in [2, 4); roots in [4, 8) move by 2 of these
    total = bad_total = 0
    for name, a, b, par in (("sen_vs", s1, o1, vsn), ("sen_vp", s2, o2, vpn), ("sen_rho", s3, o3, rhon)):
        d = a - b
        rel = np.abs(d) / np.maximum(np.abs(b), 1e-300)
        bad = (rel > 1e-5) & (d != 0)
        # every such entry = k ulps of a root, k in {1, 2}: d * 0.01 * par / ulp32(c) is an integer
        k = d * (0.01 * par[:, None, :]) / ulp[:, :, None]
        kb = k[bad]
        assert np.all(np.abs(kb - np.round(kb)) < 0.05) and np.all(np.abs(np.round(kb)) <= 4), (name, kb[:8])
        # and entries that differ at all are ulp flips too (no other source of difference)
        kd = k[d != 0]
        assert np.all(np.abs(kd - np.round(kd)) < 0.05)
        _note("depthkernel FD %s: %d of %d entries differ (all by 1-2 float32 ulps of a root), %d of them beyond 1e-5 relative"
              % (name, int((d != 0).sum()), d.size, int(bad.sum())))
        total += d.size; bad_total += int(bad.sum())
    assert bad_total <= 0.005 * total, (bad_total, total)


def test_forward_end_to_end_gpu_tables(gpu, oracle, test1):
    """Whole hot path on the GPU (K1,K2,K0,K3,K4,K5), checked against the reference's own output."""
    from dazimsurftomo_b200 import formats as fm
    p = test1["para"]
    r = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, test1["sv"])
    c = fm.forward_velocities(test1["sv"], r["dsurf"] + r["obsTaa"])
    assert np.abs(c - test1["gold_c"]).max() < 1.5e-5
    assert r["times"]["kernels_ms"] > 0
    o = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"],
                      test1["gc"], test1["gs"], nthreads=4)
    assert np.array_equal(r["dsurf"], o["dsurf"])
    assert np.abs(r["obsTaa"] - o["obsTaa"]).max() <= 1e-5 * np.abs(o["obsTaa"]).max()


def test_partition_matches_single(gpu, oracle, test1, test1_tables):
    """(period x source) ranges run as separate plans (what each rank of a multi-GPU job does) and
    concatenated in rank order give exactly the single-plan system: same rows, columns, values."""
    from dazimsurftomo_b200 import partition as pt
    p = test1["para"]; sv = test1["sv"]
    pv, svs, svp, srho, _ = oracle.depthkernel(test1["vs"], test1["depz"], p.tRc, p.sublayers, nthreads=8)
    tb = dict(test1_tables, sen_vs=svs, sen_vp=svp, sen_rho=srho)
    args = (test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv)
    full = gpu.CalSurfGAnisoJoint(*args, tables=tb)
    for world in (2, 3):
        b = pt.split_units(sv, world)
        ds, rows, cols, vals = [], [], [], []
        for r in range(world):
            plan = gpu.Plan(2, *args, tb, src_begin=b[r], src_end=b[r + 1])
            plan.run()
            out = plan.fetch()
            assert plan.row0 == sv.row_offsets()[b[r]]
            ds.append(out["dsurf"]); cols.append(out["col"]); vals.append(out["val"])
            cnt = np.diff(out["rowptr"])
            rows.append(np.repeat(np.arange(plan.row0 + 1, plan.row0 + plan.rows + 1), cnt))
            plan.close()
        assert np.array_equal(np.concatenate(ds), full["dsurf"])
        assert np.array_equal(np.concatenate(rows), full["row"])
        assert np.array_equal(np.concatenate(cols), full["col"])
        assert np.array_equal(np.concatenate(vals), full["rw"])


# ---------------------------------------------------------------------------------------------
# BASELINE config 5 shape (test4_Yunnan grid: 38 x 42 x 18, 86 layers, non-square propagation grid)
@pytest.fixture(scope="module")
def yunnan():
    from dazimsurftomo_b200 import synthetic
    return synthetic.yunnan_shaped(nsta=24, src_per_period=3, nrec=8, kmax=6)


def test_yunnan_shape_depth_kernels(gpu, oracle, yunnan):
    """K1/K2 on 86-layer profiles: roots identical, TI kernels within the north-star tolerance."""
    w = yunnan
    sub = np.asfortranarray(w.vs[5:11, 7:11, :])                # 24 nodes
    pv, L = gpu.depthkernelTI(sub, w.depz, w.tRc, w.sublayers)
    opv, oL = oracle.depthkernel_ti(sub, w.depz, w.tRc, w.sublayers, nthreads=8)
    assert (pv != opv).mean() < 0.01 and np.abs(pv - opv).max() <= 1e-5 * opv.max()
    scale = np.abs(oL).max()
    assert np.abs(L - oL).max() <= 2e-5 * scale
    pv2, s1, s2, s3 = gpu.depthkernel(sub[:2, :2, :], w.depz, w.tRc[:3], w.sublayers)   # 4 nodes x 109 variants
    opv2, o1, o2, o3, _ = oracle.depthkernel(sub[:2, :2, :], w.depz, w.tRc[:3], w.sublayers, nthreads=8)
    assert (pv2 != opv2).mean() < 0.01
    for a, b in ((s1, o1), (s2, o2), (s3, o3)):
        bad = np.abs(a - b) > 1e-5 * np.abs(b).max()
        assert bad.mean() < 0.01, bad.mean()


def test_yunnan_shape_gmatrix(gpu, oracle, yunnan):
    """Iso and joint G on the non-square 176 x 196 propagation grid, tables from the oracle so that the
    eikonal / ray / assembly stages are compared bit for bit."""
    w = yunnan
    pv, L = oracle.depthkernel_ti(w.vs, w.depz, w.tRc, w.sublayers, nthreads=8)
    rng = np.random.default_rng(11)
    nxy = w.nx * w.ny
    sen = [np.asfortranarray(0.05 + 0.2 * rng.random((nxy, len(w.tRc), w.nz))) for _ in range(3)]
    tb = dict(pvRc=pv, Lsen_Gsc=L, sen_vs=sen[0], sen_vp=sen[1], sen_rho=sen[2])
    args = (w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv)
    mx = int(w.sv.dall) * 3 * (w.nx - 2) * (w.ny - 2) * (w.nz - 1) // 8
    _cmp_coo(gpu.CalSurfG(*args, tables=tb, maxnar=mx), oracle.gbuild(1, *args, tables=tb, maxnar=mx))
    _cmp_coo(gpu.CalSurfGAnisoJoint(*args, tables=tb, maxnar=mx), oracle.gbuild(2, *args, tables=tb, maxnar=mx))
    r = gpu.FwdObsTraveltimeCPS(w.vs, w.gc, w.gs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv, tables=tb)
    o = oracle.gbuild(0, w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv, w.gc, w.gs, tables=tb)
    assert np.array_equal(r["dsurf"], o["dsurf"]) and np.array_equal(r["obsTaa"], o["obsTaa"])


def test_empty_and_ragged_surveys(gpu, oracle, test1, test1_tables):
    """Edge cases: a period without sources, a source without receivers, a single-ray survey."""
    import copy
    from dazimsurftomo_b200 import partition as pt
    p = test1["para"]; sv0 = test1["sv"]
    sv = copy.deepcopy(sv0)
    ks = [k for k in range(sv.kmax) if sv.nsrcsurf1[k] > 0]
    k0 = ks[0]
    sv.nrc1 = sv.nrc1.copy(); sv.nsrcsurf1 = sv.nsrcsurf1.copy()
    sv.nrc1[0, k0] = 0                      # a source with no receivers
    if len(ks) > 1:
        sv.nsrcsurf1[ks[1]] = 0             # a period with no sources
    sv.dall = int(sum(sv.nrc1[s, k] for k in range(sv.kmax) for s in range(int(sv.nsrcsurf1[k]))))
    args = (test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv)
    r = gpu.FwdObsTraveltimeCPS(*args, tables=test1_tables)
    o = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv,
                      test1["gc"], test1["gs"], tables=test1_tables)
    assert len(r["dsurf"]) == sv.dall and np.array_equal(r["dsurf"], o["dsurf"]) and np.array_equal(r["obsTaa"], o["obsTaa"])
    one, row0 = pt.sub_survey(sv0, 1, 2)
    one.nrc1 = one.nrc1.copy(); one.nrc1[one.nrc1 > 0] = 1; one.dall = int(one.nrc1.sum())
    r = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, one, tables=test1_tables)
    o = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, one,
                      test1["gc"], test1["gs"], tables=test1_tables)
    assert one.dall == 1 and np.array_equal(r["dsurf"], o["dsurf"]) and np.array_equal(r["obsTaa"], o["obsTaa"])


def test_limits_mirror_reference_caps(gpu, test1):
    """surfdisp96's hard caps (NL=200 layers, NP=60 periods, surfdisp96.f:57-59) come back as DAZIM_ELAYERS."""
    p = test1["para"]
    vs = np.asfortranarray(test1["vs"][:3, :3, :])
    with pytest.raises(gpu.DazimError) as e:
        gpu.depthkernelTI(vs, test1["depz"], np.arange(1, 62, dtype=float), p.sublayers)      # 61 periods
    assert e.value.code == 5
    deep = np.asfortranarray(np.repeat(vs, 30, axis=2)[:, :, :100])                              # 99 intervals x 3 sub-layers
    depz = np.arange(100, dtype=np.float32) * 2.0
    with pytest.raises(gpu.DazimError) as e:
        gpu.depthkernel(deep, depz, p.tRc[:2], 2.0)
    assert e.value.code == 5


def test_nnz_overflow_is_reported(gpu, oracle, test1, test1_tables):
    """maxnar too small: the reference overruns and stops at Main_Jt.f90:523; we return DAZIM_ENNZ_OVERFLOW."""
    p = test1["para"]
    pv, svs, svp, srho, _ = oracle.depthkernel(test1["vs"], test1["depz"], p.tRc, p.sublayers, nthreads=8)
    tb = dict(test1_tables, sen_vs=svs, sen_vp=svp, sen_rho=srho)
    with pytest.raises(gpu.DazimError) as e:
        gpu.CalSurfG(test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"],
                     tables=tb, maxnar=100)
    assert e.value.code == 3


def test_forward_driver_reproduces_reference_files(gpu, test1, tmp_path):
    """SurfAAForward replacement on the reference's own inputs (subset of example/test1_syn_foward): the
    surfphase_forward.dat it writes matches the reference's shipped output line by line (coordinates
    exactly, velocities to the printed f9.5 +- 1.5e-5) and period_Azm_tomo.real matches the golden table."""
    import shutil
    from conftest import GOLD
    from dazimsurftomo_b200 import forward
    for f in ("para.in", "MODVs.true", "MODGc.true", "MODGs.true"):
        shutil.copy(os.path.join(GOLD, f), tmp_path / f)
    shutil.copy(os.path.join(GOLD, "surfdata_subset.dat"), tmp_path / "Surfphase_RV3_5_40s_1s.dat")
    out = forward.run(str(tmp_path / "para.in"))
    mine = open(tmp_path / "surfphase_forward.dat").read().split("\n")
    gold = open(os.path.join(GOLD, "surfphase_subset.dat")).read().split("\n")
    assert len(mine) == len(gold)
    for a, b in zip(mine, gold):
        if not b.strip():
            continue
        if b.startswith("#"):
            assert a.split() == b.split()
        else:
            ta, tb = a.split(), b.split()
            assert ta[:2] == tb[:2] and abs(float(ta[2]) - float(tb[2])) < 1.5e-5
    g = test1["azm"]                       # period_Azm_tomo.real of the reference (36 periods x 225 nodes x 9 cols)
    t = out["azim"]
    n = min(len(g), len(t))
    assert n == 8100
    assert np.abs(t[:n, 3] - g[:n, 3]).max() < 6e-6                  # isotropic phase velocity (col 4)
    assert np.abs(t[:n, 7:9] - g[:n, 7:9]).max() < 2e-5              # Sum Lsen*Gc, Sum Lsen*Gs (cols 8-9)


def test_more_than_2_31_nonzeros(gpu):
    """BASELINE config 5 (Yunnan-shaped grid, 300 stations, all pairs, 36 periods): the joint G has 5.4e9
    non-zeros.  Row pointers must stay monotone past 2^31 (regression: the count scan accumulated in int)."""
    from dazimsurftomo_b200 import synthetic
    w = synthetic.yunnan_shaped()
    pv, svs, svp, srho = gpu.depthkernel(w.vs, w.depz, w.tRc, w.sublayers)
    pv2, L = gpu.depthkernelTI(w.vs, w.depz, w.tRc, w.sublayers)
    tb = dict(pvRc=pv, sen_vs=svs, sen_vp=svp, sen_rho=srho, Lsen_Gsc=L)
    plan = gpu.Plan(2, w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv, tb)
    plan.run()
    nnz = plan.nnz
    out = plan.fetch(csr="rowptr")
    plan.close()
    rp = out["rowptr"]
    assert nnz > 2 ** 31 and rp[0] == 0 and rp[-1] == nnz
    d = np.diff(rp)
    assert d.min() >= 0 and d.max() < 3 * 36 * 40 * 17 and np.isfinite(out["dsurf"]).all() and out["dsurf"].min() > 0


@pytest.mark.parametrize("env", [dict(DAZIM_DUO="0", DAZIM_SPC="2"), dict(DAZIM_DUO="0", DAZIM_SPC="2", DAZIM_HCAP="64"),
                                 dict(DAZIM_DUO="0", DAZIM_SPC="1"), dict(DAZIM_DUO="1", DAZIM_DUO_MINB="16"),
                                 dict(DAZIM_DUO="1", DAZIM_HCAP="64"), dict(DAZIM_TPS="1", DAZIM_COH="0"),
                                 dict(DAZIM_TPS="1", DAZIM_COH="0", DAZIM_HCAP="16"), dict(DAZIM_TPS="1", DAZIM_COH="1"),
                                 dict(DAZIM_TPS="1", DAZIM_COH="1", DAZIM_HCAP="16"), dict(DAZIM_TPS="1", DAZIM_TPS_PER_SM="4")])
def test_every_eikonal_kernel_variant_is_bit_identical(gpu, oracle, test1, test1_tables, monkeypatch, env):
    """The library picks the eikonal kernel from the number of solves (two-warp latency kernel, half-warp
    throughput kernel, shared / spilled heap).  Force each variant on the same inputs: fields and G must not change."""
    p = test1["para"]
    pv = np.ascontiguousarray(test1_tables["pvRc"][:, 3])
    sv = test1["sv"]
    src = [(float(sv.scxf[s, 0]), float(sv.sczf[s, 0])) for s in range(int(sv.nsrcsurf1[0]))]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    r = gpu.fmm_solve(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, [s[0] for s in src], [s[1] for s in src])
    for i, (x, z) in enumerate(src):
        o = oracle.fmm_source(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, x, z)
        nzr, nxr = o["geom"][0], o["geom"][1]
        assert np.array_equal(r["ttn"][:, :, i], o["ttn"])
        assert np.array_equal(r["nstsr"][:nzr, :nxr, i] == 0, o["nstsr"][:nzr, :nxr] == 0)
        alive = o["nstsr"][:nzr, :nxr] >= 0
        assert np.array_equal(r["ttnr"][:nzr, :nxr, i][alive], o["ttnr"][:nzr, :nxr][alive])
    a = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, sv, tables=test1_tables)
    o = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv,
                      test1["gc"], test1["gs"], tables=test1_tables)
    assert np.array_equal(a["dsurf"], o["dsurf"]) and np.array_equal(a["obsTaa"], o["obsTaa"])


# ---------------------------------------------------------------------------------------------
# BASELINE config 2 = the grid bench.py measures: S200 (202 x 202 x 9 model, 996 x 996 propagation grid).
# 11-12 heap levels, real spill traffic at hcap 512, interleaved slab offsets ~1e6: paths T1 never reaches.
@pytest.fixture(scope="module")
def s200():
    from dazimsurftomo_b200 import synthetic
    w = synthetic.s200(src_per_period=2)            # 8 periods x 2 sources = 16 solves, 512 rays
    return w, synthetic.proxy_tables(w)


_S200_ORACLE = {}


def _s200_oracle_fields(oracle, w, tb, k, srcs):
    key = (k, tuple(srcs))
    if key not in _S200_ORACLE:
        pv = np.ascontiguousarray(tb["pvRc"][:, k])
        _S200_ORACLE[key] = [oracle.fmm_source(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, x, z) for x, z in srcs]
    return _S200_ORACLE[key]


@pytest.mark.parametrize("env", [dict(DAZIM_DUO="0", DAZIM_HCAP="512"), dict(DAZIM_DUO="0", DAZIM_HCAP="4096"),
                                 dict(DAZIM_DUO="1", DAZIM_HCAP="512"), dict(DAZIM_DUO="1", DAZIM_HCAP="4096"), dict(),
                                 dict(DAZIM_TPS="1", DAZIM_COH="0", DAZIM_HCAP="448"), dict(DAZIM_TPS="1", DAZIM_COH="1", DAZIM_HCAP="64")])
def test_s200_eikonal_fields_bit_exact(gpu, oracle, s200, monkeypatch, env):
    """Coarse and refined travel-time fields + status flags on the benchmarked grid, every K3 mode (thread-per-solve
    kernel = the library's own choice, with the shared heap part bench.py runs with and a tiny one; half-warp
    throughput kernel / two-warp latency kernel with shared heaps of 512 (spilling) and 4096 entries): bit for bit
    against the oracle."""
    w, tb = s200
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    k = 5
    srcs = [(float(w.sv.scxf[s, k]), float(w.sv.sczf[s, k])) for s in range(2)]
    # + a source near the model corner: clipped source box on the big grid
    g0x = np.float32((90.0 - w.goxd) * np.pi / 180); g0z = np.float32(w.gozd * np.pi / 180); dv = np.float32(w.dvxd * np.pi / 180)
    srcs.append((float(g0x + dv * 0.6), float(g0z + dv * 198.3)))
    pv = np.ascontiguousarray(tb["pvRc"][:, k])
    r = gpu.fmm_solve(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, [s[0] for s in srcs], [s[1] for s in srcs])
    assert r["ttn"].shape[:2] == (996, 996)
    for i, o in enumerate(_s200_oracle_fields(oracle, w, tb, k, srcs)):
        assert np.array_equal(r["veln"], o["veln"])
        nzr, nxr = o["geom"][0], o["geom"][1]
        assert tuple(r["geom"][:6, i]) == tuple(o["geom"][:6])
        assert np.array_equal(r["ttn"][:, :, i], o["ttn"]), (i, int((r["ttn"][:, :, i] != o["ttn"]).sum()))
        assert np.all(r["nsts"][:, :, i] == 0)
        assert np.array_equal(r["nstsr"][:nzr, :nxr, i] == 0, o["nstsr"][:nzr, :nxr] == 0)      # alive sets
        alive = o["nstsr"][:nzr, :nxr] >= 0
        assert np.array_equal(r["ttnr"][:nzr, :nxr, i][alive], o["ttnr"][:nzr, :nxr][alive])
        assert np.array_equal(r["nstsr"][:nzr, :nxr, i], o["nstsr"][:nzr, :nxr])                # heap slots too


@pytest.mark.parametrize("env", [dict(DAZIM_DUO="0"), dict(DAZIM_DUO="1"), dict()])
def test_s200_joint_system_bit_exact(gpu, oracle, s200, monkeypatch, env):
    """dsurf + joint COO (rows, columns, values) of 16 solves / 512 rays on the S200 grid against the oracle."""
    w, tb = s200
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    args = (w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv)
    mx = int(w.sv.dall) * 60000
    key = "joint"
    if key not in _S200_ORACLE:
        _S200_ORACLE[key] = oracle.gbuild(2, *args, tables=tb, maxnar=mx, nthreads=8)
    o = _S200_ORACLE[key]
    r = gpu.CalSurfGAnisoJoint(*args, tables=tb, maxnar=mx)
    assert r["nar"] == o["nar"] > 0
    assert np.array_equal(r["row"], o["row"]) and np.array_equal(r["col"], o["col"])      # bit-exact sparsity pattern
    assert np.array_equal(r["dsurf"], o["dsurf"])                                          # bit-exact travel times
    assert r["times"]["n_accept"] == o["n_accept"] and r["times"]["n_steps"] == o["n_steps"]
    # values: the dVs block is float32 arithmetic only -> bit for bit.  The Gc / Gs blocks carry cos/sin(2 psi) from
    # azdist's double-precision tan/atan/acos/atan2 chain (rpathsAzim.f90:687): CUDA's and glibc's libm differ in the
    # last ulp of a double now and then, which flips the float32 rounding of a ~1000-step path integral in a few
    # entries.  Count them and hold them to the north star's 1e-5 relative.
    nparpi = (w.nx - 2) * (w.ny - 2) * (w.nz - 1)
    iso = o["col"] <= nparpi
    assert np.array_equal(r["rw"][iso], o["rw"][iso])
    a, b = r["rw"][~iso].astype(np.float64), o["rw"][~iso].astype(np.float64)
    ndiff = int((a != b).sum())
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    _note("s200_joint %s: %d of %d Gc/Gs entries not bit-identical (%.4f %%), max rel %.2e; dVs block %d entries identical"
          % (env, ndiff, a.size, 100.0 * ndiff / a.size, rel.max() if a.size else 0.0, int(iso.sum())))
    assert rel.max() < 1e-5 and ndiff < 0.01 * a.size


def test_footprint_pool_grows_inside_plan_run(gpu, oracle, test1, test1_tables, monkeypatch):
    """a plan reused across outer iterations must survive rays that need more footprint room than the
    straight-line estimate gave it.  A pool 50x too small: plan.run() re-allocates and re-runs by itself."""
    p = test1["para"]
    pv, svs, svp, srho, _ = oracle.depthkernel(test1["vs"], test1["depz"], p.tRc, p.sublayers, nthreads=8)
    tb = dict(test1_tables, sen_vs=svs, sen_vp=svp, sen_rho=srho)
    args = (test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"])
    full = gpu.CalSurfGAnisoJoint(*args, tables=tb)
    monkeypatch.setenv("DAZIM_POOL_SCALE", "0.02")
    plan = gpu.Plan(2, *args, tb)
    plan.run()
    out = plan.fetch()
    plan.close()
    assert np.array_equal(out["dsurf"], full["dsurf"]) and np.array_equal(out["val"], full["rw"])
    assert np.array_equal(out["col"], full["col"])


def test_cohort_kernel_falls_back_when_its_heap_workspace_overflows(gpu, oracle, test1, test1_tables, monkeypatch):
    """The cohort kernel's spill workspace is a bound on the narrow band (8 x grid edge + 1 024); if a solve ever outgrows
    it the plan re-runs on the round-1 kernels instead of failing (plan_run).  Forced here with a 16-entry shared heap and
    a 32-entry spill area: same travel times as the oracle."""
    p = test1["para"]
    monkeypatch.setenv("DAZIM_TPS", "1")
    monkeypatch.setenv("DAZIM_HCAP", "16")
    monkeypatch.setenv("DAZIM_TPS_HSPILL", "32")
    r = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                p.dvxd, p.dvzd, test1["sv"], tables=test1_tables)
    o = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"],
                      test1["gc"], test1["gs"], tables=test1_tables)
    assert np.array_equal(r["dsurf"], o["dsurf"]) and np.array_equal(r["obsTaa"], o["obsTaa"])
