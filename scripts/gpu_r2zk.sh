#!/bin/bash
# round 2, call 37 (2 GPUs): row-distributed inversion tail -- pytest entry (subset cases), then the reference's full test3
# inversion with the rows left in place against the gathered path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist_lsmr.py -q -x > gpurun_out/r2zk_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r2zk_pytest.log
grep "row-distributed inversion" gpurun_out/parity_notes.txt | tail -1 | cut -c1-1200
python - <<'PY'
import lzma, os
inv = "tests/golden/inv"
for d in ("/tmp/t3a", "/tmp/t3b"):
    os.makedirs(d, exist_ok=True)
    for f in ("para.in", "MOD"):
        open(os.path.join(d, f), "w").write(open(os.path.join(inv, "test3_" + f)).read())
    with lzma.open(os.path.join(inv, "surfphase_forward_RV3th.dat.xz"), "rb") as f:
        open(os.path.join(d, "surfphase_forward_RV3th.dat"), "wb").write(f.read())
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 -m dazimsurftomo_b200.invert /tmp/t3a/para.in --rows ) > gpurun_out/r2zk_invert_test3_2gpu_rows.log 2>&1
tail -4 gpurun_out/r2zk_invert_test3_2gpu_rows.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 -m dazimsurftomo_b200.invert /tmp/t3b/para.in ) > gpurun_out/r2zk_invert_test3_2gpu_gathered.log 2>&1
tail -4 gpurun_out/r2zk_invert_test3_2gpu_gathered.log
python - <<'PY' | tee gpurun_out/r2_invert_test3_2gpu_rows_vs_gathered.json
import json, numpy as np
a = np.loadtxt("/tmp/t3a/Gc_Gs_model.inv"); b = np.loadtxt("/tmp/t3b/Gc_Gs_model.inv")
sh = np.load("tests/golden/inv/test3_iter.npz")["shipped"]
print(json.dumps({"test3 on 2 GPUs, rows left in place vs gathered": {"max_abs_diff_table": float(np.abs(a - b).max()),
      "byte_identical": open("/tmp/t3a/Gc_Gs_model.inv").read() == open("/tmp/t3b/Gc_Gs_model.inv").read()},
      "rows left in place vs shipped Gc_Gs_model.inv": {"max_abs_dVs_mid_km_s": float(np.abs(a[:, 3] - sh[:, 0]).max()),
      "max_abs_dGc_percent": float(np.abs(a[:, 6] - sh[:, 1]).max()), "max_abs_dGs_percent": float(np.abs(a[:, 7] - sh[:, 2]).max())}}))
PY
