"""The thread-per-solve eikonal kernel's LOGIC on the CPU (no device): csrc/dazim_tps.h is __host__ __device__, and
dazim_debug_fmm_host_twin runs exactly those functions for one source.  Compared bit for bit with the oracle -- fields,
alive sets and the heap slots of the close nodes (= the reference's nstsr) -- on the reference's test1 grid and on the
benchmarked S200 grid, with a tiny shared heap so that the spill path is exercised as well."""
import numpy as np
import pytest

from dazimsurftomo_b200 import api


def _cmp(r, o):
    nzr, nxr = o["geom"][0], o["geom"][1]
    assert tuple(r["geom"][:6]) == tuple(o["geom"][:6])
    assert np.array_equal(r["ttn"], o["ttn"]), int((r["ttn"] != o["ttn"]).sum())
    assert np.all(r["nsts"] == 0)
    assert np.array_equal(r["nstsr"][:nzr, :nxr], o["nstsr"][:nzr, :nxr])          # status AND heap slots
    alive = o["nstsr"][:nzr, :nxr] >= 0
    assert np.array_equal(r["ttnr"][:nzr, :nxr][alive], o["ttnr"][:nzr, :nxr][alive])


@pytest.mark.parametrize("hcap", [448, 16, 8])
def test_host_twin_matches_oracle_on_test1(oracle, test1, test1_tables, hcap):
    p = test1["para"]; sv = test1["sv"]
    g0x = np.float32((90.0 - p.goxd) * np.pi / 180); g0z = np.float32(p.gozd * np.pi / 180)
    dv = np.float32(p.dvxd * np.pi / 180)
    for k in (0, 3):
        pv = np.ascontiguousarray(test1_tables["pvRc"][:, k])
        src = [(float(sv.scxf[s, 0]), float(sv.sczf[s, 0])) for s in range(int(sv.nsrcsurf1[0]))]
        # corner / edge sources: clipped source box, literal exit rule
        src += [(float(g0x + dv * 0.3), float(g0z + dv * 0.4)), (float(g0x + dv * 13.9), float(g0z + dv * 7.2)),
                (float(g0x + dv * 6.0), float(g0z + dv * 13.95))]
        for x, z in src:
            o = oracle.fmm_source(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, x, z)
            r = api.fmm_host_twin(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, x, z, hcap=hcap)
            _cmp(r, o)
            assert r["n_accept"] > 5000


def test_host_twin_matches_oracle_on_s200(oracle):
    from dazimsurftomo_b200 import synthetic
    w = synthetic.s200(src_per_period=2)
    tb = synthetic.proxy_tables(w)
    k = 5
    pv = np.ascontiguousarray(tb["pvRc"][:, k])
    x, z = float(w.sv.scxf[1, k]), float(w.sv.sczf[1, k])
    o = oracle.fmm_source(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, x, z)
    r = api.fmm_host_twin(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, x, z, hcap=448)
    _cmp(r, o)
    assert r["n_accept"] > 990000


def test_host_twin_matches_oracle_on_yunnan_shape(oracle):
    """Non-square 176 x 196 propagation grid (test4_Yunnan shape)."""
    from dazimsurftomo_b200 import synthetic
    w = synthetic.yunnan_shaped(nsta=24, src_per_period=3, nrec=8, kmax=6)
    tb = synthetic.proxy_tables(w)
    for k, s_ in ((0, 0), (5, 2)):
        pv = np.ascontiguousarray(tb["pvRc"][:, k])
        x, z = float(w.sv.scxf[s_, k]), float(w.sv.sczf[s_, k])
        o = oracle.fmm_source(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, x, z)
        for hcap in (448, 32):
            _cmp(api.fmm_host_twin(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, x, z, hcap=hcap), o)


@pytest.mark.parametrize("when", [0, 1])
def test_records_computed_ahead_protocol_is_exact(oracle, test1, test1_tables, when):
    """The cohort kernel's stencil threads compute the neighbour records of the node PREDICTED for the next accept while the
    current updates are applied (coh_march_stencil); the heap lane uses them if (node, key) came true and patches the one
    thing they can miss (a neighbour inserted meanwhile).  Replayed serially on the host with the records gathered at either
    end of that window: the march must stay bit-identical to the oracle, with the prediction failing sometimes and the
    patch being needed sometimes (otherwise this test exercises nothing)."""
    p = test1["para"]; sv = test1["sv"]
    tot = np.zeros(3, np.int64)
    for k in (0, 3):
        pv = np.ascontiguousarray(test1_tables["pvRc"][:, k])
        for s in range(0, int(sv.nsrcsurf1[0]), 7):
            x, z = float(sv.scxf[s, 0]), float(sv.sczf[s, 0])
            o = oracle.fmm_source(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, x, z)
            for hcap in (448, 16):
                r = api.fmm_host_twin(p.nx, p.ny, p.goxd, p.gozd, p.dvxd, p.dvzd, pv, x, z, hcap=hcap, ahead=when)
                _cmp(r, o)
                tot += np.array(r["ahead_stats"])
    from conftest import note
    note("records-ahead protocol on the host (when=%d): %d rounds predicted, %d not, %d records patched" % (when, *tot))
    assert tot[0] > 50 * tot[1] and tot[1] > 0          # predicted almost always, but not always
    if when == 0:
        assert tot[2] > 0                                # gathered before the updates: the inserted-meanwhile patch is needed


def test_records_computed_ahead_on_s200(oracle):
    from dazimsurftomo_b200 import synthetic
    w = synthetic.s200(src_per_period=2)
    tb = synthetic.proxy_tables(w)
    k = 5
    pv = np.ascontiguousarray(tb["pvRc"][:, k])
    x, z = float(w.sv.scxf[1, k]), float(w.sv.sczf[1, k])
    o = oracle.fmm_source(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, x, z)
    r = api.fmm_host_twin(w.nx, w.ny, w.goxd, w.gozd, w.dvxd, w.dvzd, pv, x, z, hcap=448, ahead=0)
    _cmp(r, o)
    hit, miss, patched = r["ahead_stats"]
    from conftest import note
    note("records-ahead protocol on the host, S200 solve: %d rounds predicted, %d not, %d records patched" % (hit, miss, patched))
    assert hit + miss == r["n_accept"] and miss < hit // 1000
