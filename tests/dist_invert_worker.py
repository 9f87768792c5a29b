"""Worker of tests/test_gpu_dist_lsmr.py::test_rows_inversion_two_gpus: torchrun, one process per GPU.

The reference's test2 (isotropic) and test3 (joint) inversions on the 1 240-ray subset, two outer iterations, (a) with
the row blocks all-gathered and the tail replicated (the established multi-GPU path: byte-identical to one GPU) and
(b) with the rows left in place and the tail row-distributed (Plan.iterate_rows).  (b) must give every rank the same
model, the same pre-solve statistics as (a) bit for bit, and a model within LSMR's parity tolerance of (a)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dazimsurftomo_b200 import invert  # noqa: E402

INV = os.path.join(ROOT, "tests", "golden", "inv")


def write_case(d, tag, maxiter, weightVs):
    lines = open(os.path.join(INV, "%s_para.in" % tag)).read().splitlines()
    lines[3] = "surfphase_subset.dat                 c: traveltime data file"
    lines[11] = "%d                                   c: maximum of iteration" % maxiter
    lines[14] = "%g                                  c: smoothing for dVsv" % weightVs
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "para.in"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(d, "MOD"), "w").write(open(os.path.join(INV, "%s_MOD" % tag)).read())
    open(os.path.join(d, "surfphase_subset.dat"), "w").write(
        open(os.path.join(ROOT, "tests", "golden", "test1", "surfphase_subset.dat")).read())


def main():
    base = sys.argv[1]
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    out, ok = dict(world=world), True
    null = open(os.devnull, "w")
    for tag in ("test2", "test3"):
        d = os.path.join(base, tag)
        if rank == 0:
            write_case(d, tag, 2, 2.0)
        dist.barrier()
        a = invert.run(os.path.join(d, "para.in"), write_files=False, log_stream=null, rows=False)
        b = invert.run(os.path.join(d, "para.in"), write_files=False, log_stream=null, rows=True)
        t = torch.from_numpy(np.ascontiguousarray(b["vsf"]).view(np.int32).copy()).cuda()
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        same = all(bool(torch.equal(g[0], x)) for x in g)
        scale = float(np.abs(a["vsf"]).max())
        rel = float(np.abs(a["vsf"] - b["vsf"]).max() / scale)
        rel_g = 0.0 if a["para"].iso_mod else float(max(np.abs(a["gcf"] - b["gcf"]).max(), np.abs(a["gsf"] - b["gsf"]).max()))
        ha, hb = a["history"], b["history"]
        before_same = ha[0]["before"] == hb[0]["before"] and ha[0]["mean_weight"] == hb[0]["mean_weight"] and \
            ha[0]["nar"] == hb[0]["nar"] and ha[0]["count3"] == hb[0]["count3"]
        out[tag] = dict(identical_on_all_ranks=same, rel_vs_to_gathered=rel, abs_gcgs_to_gathered=rel_g,
                        first_iteration_statistics_bit_identical=bool(before_same),
                        itn=[[x["lsmr"]["itn"] for x in ha], [x["lsmr"]["itn"] for x in hb]],
                        after_rms=[ha[-1]["after"]["rms"], hb[-1]["after"]["rms"]],
                        tail_ms=[a["gpu_ms"]["iterate"], b["gpu_ms"]["iterate"]], gather_ms=a["gpu_ms"].get("gather", 0.0),
                        tail_ms_per_iteration=[[x["step_ms"] for x in ha], [x["step_ms"] for x in hb]],
                        lsmr_ms_per_iteration=[[x["lsmr"]["solve_ms"] for x in ha], [x["lsmr"]["solve_ms"] for x in hb]])
        ok = ok and same and before_same and rel < 2e-4 and rel_g < 2e-4 and abs(ha[-1]["after"]["rms"] - hb[-1]["after"]["rms"]) < 1e-4
    out["ok"] = bool(ok)
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
