"""Pin the CPU oracle against the reference's own shipped outputs (example/test1_syn_foward/output)."""
import os

import numpy as np
import pytest

from dazimsurftomo_b200 import formats as fm
from conftest import REF_EX


def test_surfdisp96_kat(oracle):
    """SURVEY s4 KAT: node (2,2) of MODVs.true, rows lon=101.25 lat=26.5 of period_Azm_tomo.real."""
    vs = [3.2, 3.4, 3.8, 4.2]; dep = [0, 10, 35, 60]
    vp, rho = zip(*[oracle.brocher(v) for v in vs])
    m = oracle.refine_layer_mdl(2.0, dep, vp, vs, rho)
    assert m["rmax"] == 10 and list(m["nsublay"]) == [3, 3, 3]
    cg, nev = oracle.surfdisp96(m["rthk"], m["rvp"], m["rvs"], m["rrho"], np.arange(5, 41, dtype=float))
    ref = [3.04704, 3.06911, 3.09024, 3.11063, 3.13054, 3.15021, 3.16979, 3.18940, 3.20911, 3.22894, 3.24892,
           3.26901, 3.28919, 3.30941, 3.32958, 3.34964, 3.36950, 3.38908, 3.40828, 3.42704, 3.44526, 3.46289,
           3.47987, 3.49617, 3.51175, 3.52660, 3.54071, 3.55409, 3.56675, 3.57870, 3.58997, 3.60059, 3.61058,
           3.61999, 3.62883, 3.63715]
    assert np.abs(cg - np.array(ref)).max() < 5.1e-6
    assert np.all(cg == cg.astype(np.float32))      # cg holds float32-rounded values (surfdisp96.f:292,297)
    assert nev > 300


def _azm_rows(p, arr2d):
    nx, ny = p.nx, p.ny
    out = []
    for tt in range(p.kmaxRc):
        for jj in range(1, ny - 1):
            for ii in range(1, nx - 1):
                out.append(arr2d[jj * nx + ii, tt])
    return np.array(out)


def test_phase_velocity_map_vs_golden(test1, test1_tables):
    """col 4 of period_Azm_tomo.real = pvRc at 225 interior nodes x 36 periods, f10.5."""
    g = test1["azm"]
    mine = _azm_rows(test1["para"], test1_tables["pvRc"])
    assert g.shape == (8100, 9)
    assert np.abs(mine - g[:, 3]).max() < 6e-6
    assert abs(g[:, 3].min() - 2.76034) < 1e-5 and abs(g[:, 3].max() - 4.01562) < 1e-5


def test_anisotropy_maps_vs_golden(test1, test1_tables):
    """cols 8-9 = sum_k Lsen_Gsc*Gc/Gs (tregn96 kernels folded by depthkernelTI), f10.5."""
    p = test1["para"]
    tv = np.zeros(((p.nx - 2) * (p.ny - 2), p.kmaxRc), order="F")
    pv = test1_tables["pvRc"]
    for tt in range(p.kmaxRc):
        for jj in range(1, p.ny - 1):
            for ii in range(1, p.nx - 1):
                tv[(jj - 1) * (p.nx - 2) + ii - 1, tt] = pv[jj * p.nx + ii, tt]
    rows = fm.azim_map(p.nx, p.ny, p.nz, p.goxd, p.gozd, p.dvxd, p.dvzd, p.tRc, test1["gc"], test1["gs"],
                       test1_tables["Lsen_Gsc"], tv)
    g = test1["azm"]
    assert np.abs(rows[:, 7] - g[:, 7]).max() < 6e-6      # cosTmp
    assert np.abs(rows[:, 8] - g[:, 8]).max() < 6e-6      # sinTmp
    assert np.abs(rows[:, 6] - g[:, 6]).max() < 6e-6      # amplitude
    # tregn96 KAT: node lon 103.0 lat 24.75, T=5,6,7
    sel = (np.abs(g[:, 0] - 103.0) < 1e-4) & (np.abs(g[:, 1] - 24.75) < 1e-4)
    assert np.abs(rows[sel][:3, 7] - np.array([-0.02867, -0.02141, -0.01347])).max() < 6e-6


def test_forward_subset_vs_golden(oracle, test1, test1_tables):
    """End-to-end hot path: c = delsph/(T_FMM + T_aa) for the fixture rays (periods 5-8 s) of the
    reference's own output, f9.5."""
    p = test1["para"]
    r = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"],
                      test1["gc"], test1["gs"], tables=test1_tables, nthreads=4)
    c = fm.forward_velocities(test1["sv"], r["dsurf"] + r["obsTaa"])
    g = test1["gold_c"]
    assert len(c) == len(g) == test1["sv"].dall > 1000
    assert abs(c[0] - 3.18917) < 1e-5          # ray KAT of SURVEY s4
    assert np.abs(c - g).max() < 1.5e-5        # f9.5 print quantum + float32 delsph
    assert r["rbint"] == 0


@pytest.mark.skipif(not os.path.isdir(REF_EX), reason="reference examples only exist in the build container")
def test_forward_full_vs_golden(oracle, test1, test1_tables):
    """All 261 360 rays of example/test1_syn_foward/output/surfphase_forward_RV3th.dat.

    Periods 1-4 (5-8 s, 29 040 rays) must agree to the print precision.  Beyond that the shipped
    file is not reproducible from the shipped sources + inputs (DESIGN.md "Golden-file findings"):
    (i) the reference's eikonal scheme is chaotic at the float32-ulp level (isolated rays jump by
    ~0.1 s when the phase-velocity map moves by 1 ulp), (ii) from ~13 s on the file drifts smoothly
    away from the phase velocities printed in the reference's own period_Azm_tomo.real."""
    p = test1["para"]
    sv = fm.read_surfdata(os.path.join(REF_EX, p.datafile), p.kmaxRc)
    r = oracle.gbuild(0, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv,
                      test1["gc"], test1["gs"], tables=test1_tables, nthreads=8)
    c = fm.forward_velocities(sv, r["dsurf"] + r["obsTaa"])
    g = fm.read_surfphase_velocities(os.path.join(REF_EX, "output", "surfphase_forward_RV3th.dat"))
    assert len(c) == len(g) == 261360
    d = np.abs(c - g)
    n = 7260
    assert d[:4 * n].max() < 1.5e-5, (d[:4 * n].max(), int((d[:4 * n] > 1.5e-5).sum()))
    for k in (4, 5):      # 9 s, 10 s: only the chaotic outliers differ
        assert np.median(d[k * n:(k + 1) * n]) < 6e-6
        assert (d[k * n:(k + 1) * n] > 1.5e-5).mean() < 0.03


def test_threads_do_not_change_results(oracle, test1, test1_tables):
    p = test1["para"]
    a = oracle.gbuild(2, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"],
                      tables=dict(test1_tables, **_iso_tables(oracle, test1)), nthreads=1)
    b = oracle.gbuild(2, test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"],
                      tables=dict(test1_tables, **_iso_tables(oracle, test1)), nthreads=3)
    assert a["nar"] == b["nar"] > 0
    assert np.array_equal(a["rw"], b["rw"]) and np.array_equal(a["col"], b["col"]) and np.array_equal(a["row"], b["row"])
    assert np.array_equal(a["dsurf"], b["dsurf"])


_ISO = {}


def _iso_tables(oracle, test1):
    if not _ISO:
        p = test1["para"]
        pv, svs, svp, srho, nev = oracle.depthkernel(test1["vs"], test1["depz"], p.tRc, p.sublayers, nthreads=8)
        _ISO.update(sen_vs=svs, sen_vp=svp, sen_rho=srho)
    return _ISO
