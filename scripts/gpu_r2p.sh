#!/bin/bash
# round 2, call 16 (4 GPUs): N=4 bench, dazim_gbuild_multi on 4 devices, the reference's test3 inversion on 4 GPUs vs 1
mkdir -p gpurun_out /tmp/t3a /tmp/t3b
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log; tail -n 3 gpurun_out/r2p_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r2p_bench_n4.log 2> gpurun_out/r2p_bench_n4.err; echo "rc=$?"; tail -n 1 gpurun_out/r2p_bench_n4.log > gpurun_out/r2_bench_S200_4gpu.json; python scripts/show_bench.py gpurun_out/r2_bench_S200_4gpu.json | cut -c1-400
python - <<'PY'
import lzma, os
inv = "tests/golden/inv"
for d in ("/tmp/t3a", "/tmp/t3b"):
    for f in ("para.in", "MOD"):
        open(os.path.join(d, f), "w").write(open(os.path.join(inv, "test3_" + f)).read())
    with lzma.open(os.path.join(inv, "surfphase_forward_RV3th.dat.xz"), "rb") as f:
        open(os.path.join(d, "surfphase_forward_RV3th.dat"), "wb").write(f.read())
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29642 -m dazimsurftomo_b200.invert /tmp/t3a/para.in ) > gpurun_out/r2p_invert_test3_4gpu.log 2>&1
tail -3 gpurun_out/r2p_invert_test3_4gpu.log
( time timeout 600 python -m dazimsurftomo_b200.invert /tmp/t3b/para.in ) > gpurun_out/r2p_invert_test3_1gpu.log 2>&1
tail -3 gpurun_out/r2p_invert_test3_1gpu.log
python - <<'PY' | tee gpurun_out/r2_invert_test3_4gpu_vs_1gpu.json
import json, numpy as np
a = np.loadtxt("/tmp/t3a/Gc_Gs_model.inv"); b = np.loadtxt("/tmp/t3b/Gc_Gs_model.inv")
sh = np.load("tests/golden/inv/test3_iter.npz")["shipped"]
same = open("/tmp/t3a/Gc_Gs_model.inv").read() == open("/tmp/t3b/Gc_Gs_model.inv").read()
same_vs = open("/tmp/t3a/DSurfTomo.inv").read() == open("/tmp/t3b/DSurfTomo.inv").read()
print(json.dumps({"test3 on 4 GPUs vs 1 GPU": {"Gc_Gs_model.inv byte-identical": same, "DSurfTomo.inv byte-identical": same_vs,
      "max_abs_diff": float(np.abs(a - b).max())},
      "4 GPUs vs shipped Gc_Gs_model.inv": {"max_abs_dVs_mid_km_s": float(np.abs(a[:, 3] - sh[:, 0]).max()),
      "max_abs_dGc_percent": float(np.abs(a[:, 6] - sh[:, 1]).max()), "max_abs_dGs_percent": float(np.abs(a[:, 7] - sh[:, 2]).max())}}))
PY
