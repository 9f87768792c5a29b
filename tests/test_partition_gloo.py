"""N>1 host logic on CPU (gloo, world_size 2 and 3): the (period x source) partition covers
every unit once with contiguous global rows, and the exchange step (all-gather-v of CSR row
blocks + dsurf) rebuilds exactly the single-rank system.  The per-rank compute is the ORACLE
here (tests may use it); on the GPU box the same partition/gather code runs over NCCL with
the CUDA path (tests/test_gpu_parity.py::test_partition_matches_single)."""
import os
import socket

import numpy as np
import pytest

from conftest import GOLD, ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _load():
    from dazimsurftomo_b200 import formats as fm
    p = fm.read_para_forward(os.path.join(GOLD, "para.in"))
    depz, vs = fm.read_model(os.path.join(GOLD, "MODVs.true"), p.nx, p.ny, p.nz)
    sv = fm.read_surfdata(os.path.join(GOLD, "surfdata_subset.dat"), p.kmaxRc)
    return p, depz, vs, sv


def _tables(p, depz, vs):
    from oracle import pyoracle as po
    pv, L = po.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=2)
    nx, ny, nz = vs.shape
    rng = np.random.default_rng(5)
    # isotropic FD kernels are expensive on CPU; a smooth synthetic table exercises the same assembly code
    sen = [np.asfortranarray(0.02 * rng.standard_normal((nx * ny, len(p.tRc), nz))) for _ in range(3)]
    return dict(pvRc=pv, Lsen_Gsc=L, sen_vs=sen[0], sen_vp=sen[1], sen_rho=sen[2])


def _worker(rank, world, port, align, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from dazimsurftomo_b200 import partition as pt
    from oracle import pyoracle as po
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p, depz, vs, sv = _load()
    tb = _tables(p, depz, vs)
    b = pt.split_units(sv, world, align_periods=align)
    sub, row0 = pt.sub_survey(sv, b[rank], b[rank + 1])
    o = po.gbuild(2, vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sub, tables=tb)
    nnz_row = np.bincount(o["row"] - 1, minlength=sub.dall).astype(np.int64)
    blk = dict(dsurf=torch.from_numpy(o["dsurf"]), nnz_row=torch.from_numpy(nnz_row),
               col=torch.from_numpy(o["col"].copy()), val=torch.from_numpy(o["rw"].copy()))
    full = pt.gather_rows(blk, row0)
    obst = torch.from_numpy(sub.dist / np.maximum(sub.obsvel, 1e-3))
    ms = pt.misfit_sums(obst, blk["dsurf"])
    # what the solver stage (dazim_iterate_device) is handed on every rank
    sysd = pt.assemble_system(full, vs.shape[0], vs.shape[1], vs.shape[2], joint=True)
    if rank == 0:
        q.put(dict(bounds=b, dsurf=full["dsurf"].numpy(), rw=full["rw"].numpy(), col=full["col"].numpy(),
                   row=full["row"].numpy(), nar=full["nar"], misfit=ms.numpy(),
                   sys=dict(nrow=sysd["nrow"], nnz=sysd["nnz"], cap=sysd["cap"], rowptr=sysd["rowptr"].numpy(),
                            col=sysd["col"].numpy(), val=sysd["val"].numpy(), row=sysd["row"].numpy(),
                            dsurf=sysd["dsurf"].numpy())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,align", [(2, False), (3, True)])
def test_partition_and_gather_match_single_rank(oracle, world, align):
    import torch.multiprocessing as mp
    from dazimsurftomo_b200 import partition as pt
    p, depz, vs, sv = _load()
    tb = _tables(p, depz, vs)
    ref = oracle.gbuild(2, vs, depz, p.tRc, p.sublayers, p.goxd, p.gozd, p.dvxd, p.dvzd, sv, tables=tb)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, align, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    b = got["bounds"]
    assert b[0] == 0 and b[-1] == int(sv.nsrcsurf1.sum()) and all(b[i] <= b[i + 1] for i in range(world))
    if align:
        pstart = set(np.concatenate([[0], np.cumsum(sv.nsrcsurf1)]).tolist())
        assert all(x in pstart for x in b)
    # bit-identical to the single-rank result for any world size
    assert got["nar"] == ref["nar"]
    assert np.array_equal(got["dsurf"], ref["dsurf"])
    assert np.array_equal(got["row"], ref["row"])
    assert np.array_equal(got["col"], ref["col"])
    assert np.array_equal(got["rw"], ref["rw"])
    # the assembled solver input: CSR pointers of the global rows, triplets re-housed with room for 3 Tikhonov blocks
    sy = got["sys"]
    n = ref["nar"]
    assert (sy["nrow"], sy["nnz"]) == (sv.dall, n)
    assert sy["cap"] == n + 3 * len(oracle.tikhonov(vs.shape[0], vs.shape[1], vs.shape[2], 1, True, 1.0, 1.0)["rw"])
    assert all(len(sy[k]) == sy["cap"] for k in ("col", "val", "row"))
    assert np.array_equal(sy["rowptr"], np.concatenate([[0], np.cumsum(np.bincount(ref["row"] - 1, minlength=sv.dall))]))
    assert np.array_equal(sy["col"][:n], ref["col"]) and np.array_equal(sy["val"][:n], ref["rw"])
    assert np.array_equal(sy["row"][:n], ref["row"]) and np.array_equal(sy["dsurf"], ref["dsurf"])
    obst = sv.dist / np.maximum(sv.obsvel, 1e-3)
    r = obst.astype(np.float64) - ref["dsurf"].astype(np.float64)
    assert got["misfit"][0] == len(r)
    assert abs(got["misfit"][1] - r.sum()) <= 1e-9 * max(1.0, abs(r.sum()))
    assert abs(got["misfit"][2] - (r * r).sum()) <= 1e-9 * (r * r).sum()


def test_sub_survey_row_ranges():
    from dazimsurftomo_b200 import partition as pt
    p, depz, vs, sv = _load()
    offs = sv.row_offsets()
    for world in (1, 2, 4, 8):
        b = pt.split_units(sv, world)
        tot = 0
        for r in range(world):
            sub, row0 = pt.sub_survey(sv, b[r], b[r + 1])
            assert row0 == offs[b[r]] == tot
            tot += sub.dall
            assert int(sub.nrc1.sum()) == sub.dall
        assert tot == sv.dall
    # empty range
    sub, row0 = pt.sub_survey(sv, 3, 3)
    assert sub.dall == 0 and int(sub.nsrcsurf1.sum()) == 0


def _worker_tables(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dazimsurftomo_b200 import partition as pt
    from oracle import pyoracle as po
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p, depz, vs, sv = _load()
    nx, ny, nz = vs.shape
    strips = pt.node_strips(ny, world)
    sub = pt.strip_model(vs, strips[rank], strips[rank + 1])
    pv, L = po.depthkernel_ti(sub, depz, p.tRc, p.sublayers, nthreads=1)
    full = pt.gather_tables(dict(pvRc=pv, Lsen_Gsc=L), nx, ny, strips, rank)
    if rank == world - 1:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_stage_a_strips_gather(oracle):
    """Depth-kernel tables computed on per-rank strips of grid rows and all-gathered equal the full-model tables."""
    import torch.multiprocessing as mp
    p, depz, vs, sv = _load()
    pv, L = oracle.depthkernel_ti(vs, depz, p.tRc, p.sublayers, nthreads=4)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_tables, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert np.array_equal(got["pvRc"], pv)
    assert np.array_equal(got["Lsen_Gsc"], L)
