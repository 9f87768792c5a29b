#!/bin/bash
# 2 GPUs: the reference's test3 (joint) through the torchrun driver -- strips + period-sharded G build + NCCL all-gather
# of the row blocks + solve on every rank -- against the shipped model and against the 1-GPU run.
mkdir -p gpurun_out /tmp/t3a /tmp/t3b
timeout 300 python -m pytest tests/test_gpu_inversion.py -m gpu -x -q -k "iterate_device" 2>&1 | tail -15
python - <<'PY'
import lzma, os
inv = "tests/golden/inv"
for d in ("/tmp/t3a", "/tmp/t3b"):
    for f in ("para.in", "MOD"):
        open(os.path.join(d, f), "w").write(open(os.path.join(inv, "test3_" + f)).read())
    with lzma.open(os.path.join(inv, "surfphase_forward_RV3th.dat.xz"), "rb") as f:
        open(os.path.join(d, "surfphase_forward_RV3th.dat"), "wb").write(f.read())
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
   -m dazimsurftomo_b200.invert /tmp/t3a/para.in ) > gpurun_out/invert_test3_2gpu.log 2>&1
tail -4 gpurun_out/invert_test3_2gpu.log
( time timeout 600 python -m dazimsurftomo_b200.invert /tmp/t3b/para.in ) > gpurun_out/invert_test3_1gpu.log 2>&1
tail -4 gpurun_out/invert_test3_1gpu.log
python - <<'PY' | tee gpurun_out/invert_test3_2gpu_compare.json
import json, numpy as np
a = np.loadtxt("/tmp/t3a/Gc_Gs_model.inv"); b = np.loadtxt("/tmp/t3b/Gc_Gs_model.inv")
sh = np.load("tests/golden/inv/test3_iter.npz")["shipped"]
same = open("/tmp/t3a/Gc_Gs_model.inv").read() == open("/tmp/t3b/Gc_Gs_model.inv").read()
same_vs = open("/tmp/t3a/DSurfTomo.inv").read() == open("/tmp/t3b/DSurfTomo.inv").read()
print(json.dumps({"test3 on 2 GPUs vs 1 GPU": {"Gc_Gs_model.inv byte-identical": same, "DSurfTomo.inv byte-identical": same_vs,
      "max_abs_diff": float(np.abs(a - b).max())},
      "2 GPUs vs shipped Gc_Gs_model.inv": {"max_abs_dVs_mid_km_s": float(np.abs(a[:, 3] - sh[:, 0]).max()),
      "max_abs_dGc_percent": float(np.abs(a[:, 6] - sh[:, 1]).max()), "max_abs_dGs_percent": float(np.abs(a[:, 7] - sh[:, 2]).max())}}))
PY
