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
    """All 261 360 rays of example/test1_syn_foward/output/surfphase_forward_RV3th.dat, every period.

    What is reproducible from the shipped sources + inputs, and is asserted here (DESIGN.md s2.1,
    numbers in profiles/r2_golden_drift.json written by scripts/golden_drift.py):
      * periods 1-4 (5-8 s, 29 040 rays): every ray to the f9.5 print quantum;
      * periods 5-8 (9-12 s): the bulk of the rays still to print precision (median), the rest are the
        eikonal scheme's ulp-chaos outliers (1 float32 ulp of the phase-velocity map moves 0.8 % of
        the rays by up to 5.3e-3, measured) -- i.e. the file's phase-velocity maps start to differ
        from the shipped inputs' in the last bits;
      * periods 9-36: the file drifts smoothly and monotonically away (median |rel| doubling per
        second of period, 0.90 % at 40 s): it was produced from phase-velocity maps / anisotropy
        kernels that differ from what the shipped MODVs/Gc/Gs.true give -- and from the reference's
        OWN period_Azm_tomo.real, which the oracle matches at all 36 periods -- in a way that is
        invisible below 9 s, i.e. below ~60 km depth.  The per-period distance is pinned against the
        committed table so that any change of the oracle (or of the finding) fails loudly."""
    import json
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
    from scripts_golden import per_period_table
    dist = fm.forward_velocities(sv, np.ones(sv.dall, np.float32)).astype(np.float64)
    tab = per_period_table(dist / g, r["dsurf"].astype(np.float64), r["obsTaa"].astype(np.float64), n)
    with open(os.path.join(os.path.dirname(__file__), "..", "profiles", "r2_golden_drift.json")) as f:
        ref = json.load(f)["table"]
    med = np.array([t["median_abs_rel"] for t in tab])
    for k in range(4):                                   # 5-8 s: exact
        assert tab[k]["share_beyond_print"] == 0.0 and abs(tab[k]["lsq_scale_Tiso"] - 1) < 1e-7
    for k in range(4, 8):                                # 9-12 s: bulk exact, outliers = ulp chaos
        assert med[k] < 4.5e-6 and tab[k]["max_abs_rel"] < 6e-3
    assert np.all(np.diff(med[4:]) > 0)                  # smooth, monotone drift from 9 s on
    assert 0.0085 < med[35] < 0.0095 and 1.0075 < tab[35]["lsq_scale_Tiso"] < 1.0082
    for t, q in zip(tab, ref):                           # the finding itself is pinned
        assert abs(t["mean_ratio"] - q["mean_ratio"]) < 2e-6, (t, q)
        assert abs(t["lsq_scale_Tiso"] - q["lsq_scale_Tiso"]) < 1e-5, (t, q)


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
