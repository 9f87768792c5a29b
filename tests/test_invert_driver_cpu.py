"""Host logic of the inversion driver (dazimsurftomo_b200/invert.py) on CPU: the loop, the log and every output file,
with the GPU entry points replaced IN THIS TEST ONLY by stand-ins built on the oracle (tests may use it; the product
path has no such fallback -- tests/test_cabi.py::test_no_device_fails_loudly).  The numerical parity of the real
entry points is the business of tests/test_gpu_inversion.py."""
import os

import numpy as np
import pytest

from conftest import ROOT

INV = os.path.join(ROOT, "tests", "golden", "inv")
F32 = np.float32


class _Handle:
    times = {"kernels_ms": 0.0}


def _install_stand_ins(monkeypatch, oracle):
    from dazimsurftomo_b200 import api

    def depthkernel(vel, depz, tRc, minthk, handle=None):
        pv, a, b, c, _ = oracle.depthkernel(vel, depz, tRc, minthk, nthreads=8)
        return pv, a, b, c

    def depthkernelTI(vel, depz, tRc, minthk, handle=None):
        return oracle.depthkernel_ti(vel, depz, tRc, minthk, nthreads=8)

    class Plan:
        def __init__(self, mode, vels, depz, tRc, minthk, goxd, gozd, dvxd, dvzd, sv, tables, src_begin=0, src_end=-1, handle=None):
            self.mode, self.geo, self.sv, self.depz, self.tRc, self.minthk = mode, (goxd, gozd, dvxd, dvzd), sv, depz, tRc, minthk
            self.vels, self.tables = np.array(vels, F32, order="F"), tables
            self.rows, self.row0 = sv.dall, 0

        def update_model(self, vels, tables):
            self.vels, self.tables = np.array(vels, F32, order="F"), tables

        def run(self):
            self.g = oracle.gbuild(self.mode, self.vels, self.depz, self.tRc, self.minthk, *self.geo, self.sv,
                                   tables=self.tables, nthreads=8)
            return {"total_ms": 0.0}

        @property
        def nnz(self):
            return int(self.g["nar"])

        def fetch(self, csr=True):
            return dict(dsurf=self.g["dsurf"])

        def iterate(self, obst, vsf, iso_inv, weightVs, weightGcs, damp, minvel, maxvel, controls=None, want_rows=False):
            g = self.g
            nx, ny, nz = vsf.shape
            maxvp = (nx - 2) * (ny - 2) * (nz - 1)
            n = maxvp if iso_inv else 3 * maxvp
            cbst = (obst - g["dsurf"]).astype(F32)
            before = oracle.res_stats(cbst)
            sig, mdt = oracle.cal_ddat_sigma(obst, cbst)
            w, cbw, rww = oracle.apply_weights(sig, cbst, g["row"], g["rw"])
            tk = oracle.tikhonov(nx, ny, nz, len(obst), iso_inv, weightGcs, weightVs, joint=not iso_inv)
            rw = np.concatenate([rww, tk["rw"]]); row = np.concatenate([g["row"], tk["row"]]); col = np.concatenate([g["col"], tk["col"]])
            dv, info = oracle.lsmr(len(obst) + tk["count3"], n, row, col, rw, np.concatenate([cbw, np.zeros(tk["count3"], F32)]),
                                   damp=damp, **oracle.lsmr_controls(iso_inv, n))
            dv, v, gc, gs = oracle.model_update(dv, vsf, iso_inv, minvel, maxvel)
            rs = oracle.residuals(maxvp, 1 if iso_inv else 3, g["rw"], g["row"], g["col"], dv, w, cbst)
            nm = oracle.model_norms(int(g["nar"]), None if iso_inv else int(g["nar"]) + tk["narVs"], rw, col, dv, weightGcs, weightVs)
            st = dict(before=before, after=oracle.res_stats(rs["resbst"]), norms=nm, lsmr=dict(info, setup_ms=0.0, solve_ms=0.0),
                      meandeltaT=mdt, mean_weight=float(w.mean()), meanabs_weighted=float(np.abs(cbw).mean()),
                      res2Nm=rs["res2Nm"], resW2Nm=rs["resW2Nm"], meanabs_Taa=float(np.abs(rs["fwdTaa"]).mean()),
                      meanabs_Tvs=float(np.abs(rs["fwdTvs"]).mean()), nar1=int(g["nar"]), nar=len(rw), count3=tk["count3"],
                      step_ms=0.0, scale_ms=0.0)
            dws = np.bincount(g["col"] - 1, weights=np.abs(rww), minlength=maxvp).astype(F32) if iso_inv else None
            out = dict(vsf=v, dv=dv, gcf=None if iso_inv else gc, gsf=None if iso_inv else gs, dws=dws, stats=st)
            if want_rows:
                out.update(sigmaT=sig, resbst=rs["resbst"], fwdTvs=rs["fwdTvs"], fwdTaa=rs["fwdTaa"])
            return out

        def close(self):
            pass

    monkeypatch.setattr(api, "depthkernel", depthkernel)
    monkeypatch.setattr(api, "depthkernelTI", depthkernelTI)
    monkeypatch.setattr(api, "Plan", Plan)


def _case(tmp_path, tag, maxiter, weightVs):
    lines = open(os.path.join(INV, "%s_para.in" % tag)).read().splitlines()
    lines[3] = "surfphase_subset.dat                 c: traveltime data file"
    lines[11] = "%d                                   c: maximum of iteration" % maxiter
    lines[14] = "%g                                  c: smoothing for dVsv" % weightVs
    (tmp_path / "para.in").write_text("\n".join(lines) + "\n")
    (tmp_path / "MOD").write_text(open(os.path.join(INV, "%s_MOD" % tag)).read())
    (tmp_path / "surfphase_subset.dat").write_text(open(os.path.join(ROOT, "tests", "golden", "test1", "surfphase_subset.dat")).read())


@pytest.mark.parametrize("tag", ["test2", "test3"])
def test_driver_loop_log_and_files(oracle, monkeypatch, tmp_path, tag):
    from dazimsurftomo_b200 import formats as fm, invert
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    _install_stand_ins(monkeypatch, oracle)
    _case(tmp_path, tag, 2, 8.0)
    log = open(tmp_path / "stdout.txt", "w")
    out = invert.run(str(tmp_path / "para.in"), handle=_Handle(), log_stream=log)
    log.close()
    p = out["para"]
    iso = tag == "test2"
    # the loop is the oracle's loop (pyoracle.invert) when the pieces are the oracle's
    depz, vs = fm.read_model(str(tmp_path / "MOD"), p.nx, p.ny, p.nz)
    sv = out["survey"]
    o = oracle.invert(vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, (sv.dist / sv.obsvel).astype(F32),
                      p.iso_mod, p.weightVs, p.weightGcs, p.damp, p.minvel, p.maxvel, 2, spfra=p.spfra, nthreads=8)
    assert np.array_equal(out["vsf"], o["vsf"]) and np.array_equal(out["gcf"], o["gcf"]) and np.array_equal(out["gsf"], o["gsf"])
    # output files, names and shapes of the reference (Main_Jt.f90:752-787)
    rows = open(tmp_path / "DSurfTomo.inv").read().splitlines()
    assert len(rows) == p.nx * p.ny * p.nz and all(len(r) == 32 for r in rows)
    assert np.abs(np.array([float(r[24:32]) for r in rows]).reshape((p.nx, p.ny, p.nz), order="F") - out["vsf"]).max() <= 5.1e-5
    d2, v2 = fm.read_model(str(tmp_path / "MOD_Ref"), p.nx, p.ny, p.nz)
    assert np.array_equal(d2, depz) and np.abs(v2 - out["vsf"]).max() <= 5.1e-5
    g = np.loadtxt(tmp_path / "Gc_Gs_model.inv")
    assert g.shape == ((p.nx - 2) * (p.ny - 2) * (p.nz - 1), 8)
    assert np.abs(g[:, 6] - out["gcf"].ravel(order="F") * 100).max() <= 5.1e-5 + 1e-4 * np.abs(g[:, 6]).max()
    for name in ("period_phaseVMOD.dat", "phaseV_FWD.dat"):
        t = np.loadtxt(tmp_path / name)
        assert t.shape == (36 * (p.nx - 2) * (p.ny - 2), 4) and t[:, 3].min() > 2.5
    first, lastv = np.loadtxt(tmp_path / "period_phaseVMOD.dat")[:, 3], np.loadtxt(tmp_path / "phaseV_FWD.dat")[:, 3]
    assert np.abs(first - lastv).max() > 1e-3                   # iteration 1's map vs the last iteration's map
    stat = open(tmp_path / "Traveltime_statis_00th.dat").read().splitlines()
    assert len(stat) == sv.dall + 1 and len(stat[1]) == (30 + 36 if iso else 30 + 48)
    assert "E" in stat[1][30:42] and stat[1][30:42].strip()[:2] in ("0.", "-0")   # Fortran E12.3: mantissa 0.ddd
    itervel = open(tmp_path / "IterVel.out").read().splitlines()
    if iso:
        assert len(itervel) == 2 * (2 + p.nz * p.ny + (p.nz - 1) * (p.ny - 2))
        assert "OUTPUT S VELOCITY AT ITERATION" in itervel[0] and len(itervel[1]) == 7 * p.nx
        assert not (tmp_path / "period_Azm_tomo.inv").exists()
    else:
        assert itervel == []
        az = np.loadtxt(tmp_path / "period_Azm_tomo.inv")
        assert az.shape == (36 * (p.nx - 2) * (p.ny - 2), 9)
    text = open(tmp_path / "para.in_inv.log").read()
    assert text == open(tmp_path / "stdout.txt").read()
    assert text.count("Before Inversion: abs mean, std, RMS of Res:") == 2 and "Program finishes successfully" in text
    assert ("invert for isotropic Vs para." in text) == iso and ("Gcs:  ||Lm||^2" in text) == (not iso)
    assert len(open(tmp_path / "lsmr.txt").read().splitlines()) == 2


def test_driver_stops_when_spfra_is_too_small(oracle, monkeypatch, tmp_path):
    """nar > maxnar = spfra*dall*nx*ny*nz*3 -> 'increase sparsity fraction(spfra)' (Main_Jt.f90:325,523)."""
    from dazimsurftomo_b200 import api, invert
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    _install_stand_ins(monkeypatch, oracle)
    _case(tmp_path, "test2", 1, 8.0)
    lines = open(tmp_path / "para.in").read().splitlines()
    lines[10] = "0.00001                              c: sparsity fraction"
    (tmp_path / "para.in").write_text("\n".join(lines) + "\n")
    with pytest.raises(api.DazimError, match="sparsity fraction"):
        invert.run(str(tmp_path / "para.in"), handle=_Handle(), log_stream=open(os.devnull, "w"))
