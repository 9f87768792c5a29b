"""The reference's REAL surveys, every solve: example/test2-3 (4 320 solves, 261 360 rays on the 71 x 71 grid) and
example/test4_Yunnan (1 469 solves, 20 877 rays on the 176 x 196 grid, real station geometry).  Travel times of every
ray from the CUDA path against the oracle, bit for bit, with the library's own kernel choice and with the cohort kernel
forced -- the parity bar the round-1 VERDICT set for any new eikonal kernel ("all 4 320 test1 solves and all 1 469
test4 solves").  Phase-velocity maps come from the oracle so that only dice + eikonal + srtimes + ray tracing are
compared."""
import os

import numpy as np
import pytest

from dazimsurftomo_b200 import formats as fm
from conftest import ROOT

pytestmark = pytest.mark.gpu
INV = os.path.join(ROOT, "tests", "golden", "inv")
_CACHE = {}


def _case(oracle, tag, tmp_path_factory):
    if tag not in _CACHE:
        d = str(tmp_path_factory.mktemp(tag))
        p = fm.read_para_inv(fm.stage_reference_example(INV, tag, d))
        depz, vs = fm.read_model(os.path.join(d, "MOD"), p.nx, p.ny, p.nz)
        sv = fm.read_surfdata(os.path.join(d, p.datafile), p.kmaxRc)
        pv, L = oracle.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=8)
        tb = dict(pvRc=pv, Lsen_Gsc=L)
        zero = np.zeros((p.nx - 2, p.ny - 2, p.nz - 1), np.float32, order="F")
        o = oracle.gbuild(0, vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, zero, zero, tables=tb, nthreads=8)
        _CACHE[tag] = (p, depz, vs, sv, tb, zero, o)
    return _CACHE[tag]


@pytest.mark.parametrize("tag,nsolve,nray", [("test3", 4320, 261360), ("test4", 1469, 20877)])
@pytest.mark.parametrize("env", [dict(), dict(DAZIM_TPS="1"), dict(DAZIM_TPS="1", DAZIM_COH_LANES="32")])
def test_every_solve_of_the_reference_surveys(gpu, oracle, tmp_path_factory, monkeypatch, tag, nsolve, nray, env):
    p, depz, vs, sv, tb, zero, o = _case(oracle, tag, tmp_path_factory)
    assert int(sv.nsrcsurf1.sum()) == nsolve and sv.dall == nray
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    r = gpu.FwdObsTraveltimeCPS(vs, zero, zero, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, tables=tb)
    assert np.array_equal(r["dsurf"], o["dsurf"]), int((r["dsurf"] != o["dsurf"]).sum())
    assert r["times"]["n_accept"] == o["n_accept"] and r["times"]["n_steps"] == o["n_steps"]


def test_joint_system_of_the_whole_yunnan_survey(gpu, oracle, tmp_path_factory):
    """All 20 877 rows of the real test4_Yunnan survey (61.6 M non-zeros, 73 440 columns): pattern, rows and the dVs block
    bit for bit; Gc / Gs values within the north star's 1e-5 (last-ulp differences of double-precision libm inside azdist
    are counted).  Depth kernels for the FD tables are stand-ins (the eikonal / ray / assembly stages are what is compared)."""
    from conftest import note
    p, depz, vs, sv, tb, zero, o0 = _case(oracle, "test4", tmp_path_factory)
    rng = np.random.default_rng(11)
    nxy = p.nx * p.ny
    sen = [np.asfortranarray(0.05 + 0.2 * rng.random((nxy, len(p.tRc), p.nz))) for _ in range(3)]
    tbj = dict(tb, sen_vs=sen[0], sen_vp=sen[1], sen_rho=sen[2])
    args = (vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv)
    mx = int(sv.dall) * 6000
    o = oracle.gbuild(2, *args, tables=tbj, nthreads=8, maxnar=mx)
    r = gpu.CalSurfGAnisoJoint(*args, tables=tbj, maxnar=mx)
    assert r["nar"] == o["nar"] > 6e7
    assert np.array_equal(r["row"], o["row"]) and np.array_equal(r["col"], o["col"]) and np.array_equal(r["dsurf"], o["dsurf"])
    nparpi = (p.nx - 2) * (p.ny - 2) * (p.nz - 1)
    iso = o["col"] <= nparpi
    assert np.array_equal(r["rw"][iso], o["rw"][iso])
    a, b = r["rw"][~iso].astype(np.float64), o["rw"][~iso].astype(np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    note("test4_Yunnan joint G: %d rows, %d non-zeros; %d of %d Gc/Gs entries not bit-identical, max rel %.2e"
         % (sv.dall, r["nar"], int((a != b).sum()), a.size, rel.max()))
    assert rel.max() < 1e-5 and (a != b).sum() < 0.01 * a.size
