"""Multi-GPU behind the C ABI: dazim_gbuild_multi (one call, several devices, one host thread + stream per device)
must give exactly the single-device result -- rows, columns, values, travel times -- and the depth-kernel tables
computed on strips must equal the ones computed in one piece.  Needs >= 2 devices (gpurun --gpus 2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [2, 4])
def test_gbuild_multi_matches_single_device(gpu, oracle, test1, test1_tables, ndev):
    if _ndev() < ndev:
        pytest.skip("needs %d CUDA devices" % ndev)
    p = test1["para"]
    pv, svs, svp, srho, _ = oracle.depthkernel(test1["vs"], test1["depz"], p.tRc, p.sublayers, nthreads=8)
    tb = dict(test1_tables, sen_vs=svs, sen_vp=svp, sen_rho=srho)
    args = (test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"])
    dev = list(range(ndev))
    for fn in (gpu.CalSurfG, gpu.CalSurfGAnisoJoint):
        a = fn(*args, tables=tb)
        a = {k: np.array(v, copy=True) for k, v in a.items() if k in ("dsurf", "rw", "row", "col")}
        b = fn(*args, tables=tb, devices=dev)
        assert b["nar"] == len(a["rw"]) > 0
        for k in ("dsurf", "row", "col", "rw"):
            assert np.array_equal(a[k], b[k]), k
        assert b["times"]["n_accept"] > 0 and b["times"]["fmm_ms"] > 0
    f1 = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                 p.dvxd, p.dvzd, test1["sv"], tables=test1_tables)
    f2 = gpu.FwdObsTraveltimeCPS(test1["vs"], test1["gc"], test1["gs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd,
                                 p.dvxd, p.dvzd, test1["sv"], tables=test1_tables, devices=dev)
    assert np.array_equal(f1["dsurf"], f2["dsurf"]) and np.array_equal(f1["obsTaa"], f2["obsTaa"])


def test_gbuild_multi_depth_kernels_on_strips(gpu, test1):
    """tables_precomputed = 0: every device computes the Thomson-Haskell tables of its strip of grid rows; merged on the
    host they equal the single-device tables bit for bit, and so does the joint system built from them."""
    if _ndev() < 2:
        pytest.skip("needs 2 CUDA devices")
    p = test1["para"]
    args = (test1["vs"], test1["depz"], p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, test1["sv"])
    a = gpu.CalSurfGAnisoJoint(*args)
    a = {k: np.array(v, copy=True) for k, v in a.items() if isinstance(v, np.ndarray)}
    b = gpu.CalSurfGAnisoJoint(*args, devices=[0, 1])
    for k in ("pvRc", "sen_vs", "sen_vp", "sen_rho", "Lsen_Gsc", "dsurf", "row", "col", "rw", "tRcV"):
        assert np.array_equal(a[k], b[k]), k


def test_fortran_symbols_use_every_listed_device(gpu, test1, monkeypatch):
    """The gfortran drop-in calsurfganisojoint_ (what Main_Jt.f90:403-406 links against) with DAZIM_DEVICES=0,1 gives the
    arrays of the single-device call."""
    if _ndev() < 2:
        pytest.skip("needs 2 CUDA devices")
    from test_fortran_abi import call_calsurfganisojoint
    one = call_calsurfganisojoint(test1)
    monkeypatch.setenv("DAZIM_DEVICES", "0,1")
    two = call_calsurfganisojoint(test1)
    for k in one:
        assert np.array_equal(one[k], two[k]), k
