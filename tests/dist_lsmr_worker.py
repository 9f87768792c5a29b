"""Worker of tests/test_gpu_dist_lsmr.py and scripts/gpu_*: run under torchrun with one process per GPU.

Builds the same seeded sparse system on every rank (tomography-like: data rows with a few dozen entries + Laplacian-like
regularisation rows), solves it (a) on this rank's GPU alone (dazim_lsmr, pinned to the oracle by tests/test_lsmr.py) and
(b) row-distributed over all ranks (dazim_lsmr_rows), and checks: every rank gets the same x / info bit for bit, the
distributed solution agrees with the single-GPU one to LSMR's parity tolerance, an uneven split and the un-captured path
give the same answer.  Rank 0 prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dazimsurftomo_b200 import api  # noqa: E402


def system(m_data, n, per_row, seed):
    rng = np.random.default_rng(seed)
    # data rows: per_row entries each at random columns (a repeated column is two triplets that add up, as in COO)
    drow = np.repeat(np.arange(1, m_data + 1, dtype=np.int64), per_row)
    dcol = rng.integers(1, n + 1, size=m_data * per_row)
    dval = rng.uniform(0.05, 1.0, m_data * per_row)
    # regularisation-like rows: 0.3 (x_j - 0.5 x_{j-1} - 0.5 x_{j+1})
    j = np.arange(1, n - 1, dtype=np.int64)
    rrow = np.repeat(m_data + j, 3)
    rcol = np.stack([j, j + 1, j + 2], axis=1).ravel()
    rval = np.tile(np.array([-0.15, 0.3, -0.15]), len(j))
    row = np.concatenate([drow, rrow]).astype(np.int32); col = np.concatenate([dcol, rcol]).astype(np.int32)
    rw = np.concatenate([dval, rval]).astype(np.float32)
    m = m_data + n - 2
    xt = rng.normal(0, 1, n).astype(np.float32)
    b = np.bincount(row - 1, weights=rw.astype(np.float64) * xt[col - 1], minlength=m).astype(np.float32)
    b[:m_data] += rng.normal(0, 0.01, m_data).astype(np.float32)
    return m, row, col, rw, b


def main():
    big = "--big" in sys.argv
    huge = "--huge" in sys.argv          # 2.3e8 entries, n = 960 000: the size of the S200-lite joint system
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    h = api.Handle(local)
    comm = api.Comm.from_torch(local)
    out = dict(world=world)
    m_data, n, per_row = (2400000, 960000, 96) if huge else ((400000, 96000, 48) if big else (6000, 900, 24))
    m, row, col, rw, b = system(m_data, n, per_row, 20260101)
    ctl = dict(damp=0.01, atol=0.0, btol=0.0, conlim=1e8, itnlim=120 if (big or huge) else 80, localSize=10)      # runs to itnlim on every path
    t0 = time.time(); x1, i1 = api.LSMR(m, n, row, col, rw, b, handle=h, **ctl); t_single = time.time() - t0

    cache = {}

    def solve(blocks, **over):
        first, cnt = blocks[rank]
        if (first, cnt) not in cache:
            cache[(first, cnt)] = api.row_block(row, col, rw, b, first, cnt)
        r_, c_, w_, b_ = cache[(first, cnt)]
        return api.LSMR_rows(comm, cnt, m, n, r_, c_, w_, b_, handle=h, **dict(ctl, **over))

    solve(api.split_rows(m, world), itnlim=2)      # first collective of the communicator: NCCL sets its channels up here
    xd, idd = solve(api.split_rows(m, world))
    # every rank holds the same answer
    t = torch.from_numpy(xd.view(np.int32).copy()).cuda()
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    same = all(bool(torch.equal(g[0], gi)) for gi in g)
    rel = float(np.linalg.norm(xd - x1) / np.linalg.norm(x1))
    out.update(m=m, n=n, nnz=int(len(rw)), identical_on_all_ranks=same, rel_to_single=rel, itn_single=i1["itn"], itn_rows=idd["itn"],
               istop_single=i1["istop"], istop_rows=idd["istop"], normr_single=i1["normr"], normr_rows=idd["normr"],
               solve_ms_single=i1["solve_ms"], solve_ms_rows=idd["solve_ms"],
               ms_per_iter_single=i1["solve_ms"] / max(1, i1["itn"]), ms_per_iter_rows=idd["solve_ms"] / max(1, idd["itn"]))
    # uneven split: rank 0 takes two thirds
    if world >= 2 and not huge:
        first0 = (2 * m) // 3
        rest = api.split_rows(m - first0, world - 1)
        blocks = [(0, first0)] + [(first0 + f, c) for f, c in rest]
        xu, iu = solve(blocks)
        out.update(rel_uneven_to_even=float(np.linalg.norm(xu - xd) / np.linalg.norm(xd)), itn_uneven=iu["itn"])
    # without the CUDA graph (kernels and all-reduces enqueued one by one): same operations, same bits
    out["nograph_identical"] = True
    if not huge:
        os.environ["DAZIM_LSMR_NOGRAPH"] = "1"
        xn, inn = solve(api.split_rows(m, world))
        del os.environ["DAZIM_LSMR_NOGRAPH"]
        out.update(nograph_identical=bool(np.array_equal(xn, xd)), ms_per_iter_rows_nograph=inn["solve_ms"] / max(1, inn["itn"]))
    ok = same and rel < 2e-4 and idd["itn"] == i1["itn"] and out.get("rel_uneven_to_even", 0.0) < 2e-4 and out["nograph_identical"]
    out["ok"] = bool(ok)
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
