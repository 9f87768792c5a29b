"""Row-distributed LSMR (lsmrModule.f90:36, aprod.f90:7): one process per GPU, NCCL all-reduce of
one n-vector + one scalar per iteration inside the captured iteration graph.  Needs >= 2 GPUs (the 1-GPU round-end
tier skips it; scripts/gpu_r2zj.sh runs it on a 2-GPU box, profiles/r2_dist_lsmr_2gpu.json)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_rows_and_row_block_partition_the_system():
    from dazimsurftomo_b200 import api
    rng = np.random.default_rng(3)
    m, n, nnz = 37, 11, 400
    row = rng.integers(1, m + 1, nnz).astype(np.int32); col = rng.integers(1, n + 1, nnz).astype(np.int32)
    rw = rng.normal(size=nnz).astype(np.float32); b = rng.normal(size=m).astype(np.float32)
    for world in (1, 2, 3, 5, 40):
        blocks = api.split_rows(m, world)
        assert len(blocks) == world and blocks[0][0] == 0 and sum(c for _, c in blocks) == m
        assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
        # the blocks' products add up to the whole system's: A^T u = sum_r A_r^T u_r, (A x)_r = A_r x
        x = rng.normal(size=n).astype(np.float64); u = rng.normal(size=m).astype(np.float64)
        full_atu = np.zeros(n); np.add.at(full_atu, col - 1, rw * u[row - 1])
        full_ax = np.zeros(m); np.add.at(full_ax, row - 1, rw * x[col - 1])
        atu = np.zeros(n); seen = 0
        for first, cnt in blocks:
            r_, c_, w_, b_ = api.row_block(row, col, rw, b, first, cnt)
            assert np.array_equal(b_, b[first:first + cnt]) and (cnt == 0 or len(r_) == 0 or (r_.min() >= 1 and r_.max() <= cnt))
            np.add.at(atu, c_ - 1, w_ * u[first:first + cnt][r_ - 1])
            ax = np.zeros(cnt); np.add.at(ax, r_ - 1, w_ * x[c_ - 1])
            assert np.allclose(ax, full_ax[first:first + cnt], rtol=0, atol=1e-12)
            seen += len(r_)
        assert seen == nnz and np.allclose(atu, full_atu, rtol=0, atol=1e-9)


def _gloo_id_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from dazimsurftomo_b200 import api
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    q.put((rank, api.Comm.broadcast_id()))
    dist.destroy_process_group()


def test_communicator_id_reaches_every_rank_over_gloo():
    """Host side of Comm.from_torch with world size 2 on CPU: rank 0's NCCL id (128 bytes) arrives on rank 1."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_gloo_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert len(got[0]) == 128 and got[0] == got[1] and any(got[0])


@pytest.mark.gpu
def test_row_distributed_lsmr_two_gpus(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29900 + os.getpid() % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_lsmr_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    d = json.loads(lines[-1])
    from conftest import note
    note("row-distributed LSMR on 2 GPUs: " + json.dumps(d))
    assert d["ok"] and d["identical_on_all_ranks"] and d["rel_to_single"] < 2e-4 and d["itn_rows"] == d["itn_single"]


@pytest.mark.gpu
def test_rows_inversion_two_gpus(gpu, tmp_path):
    """The inversion driver with the rows of G left on the ranks that built them (invert.run(rows=True),
    dazim_plan_iterate_rows) against the gathered / replicated tail, test2 (iso) and test3 (joint) subset cases."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29800 + os.getpid() % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_invert_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    d = json.loads(lines[-1])
    from conftest import note
    note("row-distributed inversion tail on 2 GPUs: " + json.dumps(d))
    assert d["ok"]
