#!/bin/bash
# round 2, call 8 (2 GPUs): multi-GPU behind the C ABI, N=2 bench with the per-rank end-to-end leg
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or partition" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -n 6 gpurun_out/r2h_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2h_bench_n2.log 2> gpurun_out/r2h_bench_n2.err; echo "rc=$?"; tail -n 1 gpurun_out/r2h_bench_n2.log | cut -c1-2500; tail -n 3 gpurun_out/r2h_bench_n2.err
# one C-ABI call on 2 devices, S200-lite, timed against the single-device call
python - <<'PY' > gpurun_out/r2h_multi_timing.txt 2>&1
import time, numpy as np
from dazimsurftomo_b200 import api, synthetic
w = synthetic.s200(src_per_period=125)
args = (w.vs, w.depz, w.tRc, w.sublayers, w.goxd, w.gozd, w.dvxd, w.dvzd, w.sv)
for dev in (None, [0, 1]):
    for rep in range(2):
        t0 = time.perf_counter()
        r = api.CalSurfGAnisoJoint(*args, maxnar=260_000_000, devices=dev)
        dt = time.perf_counter() - t0
    print("devices", dev, "wall s %.3f" % dt, "nar", r["nar"], {k: round(float(v), 1) for k, v in r["times"].items() if k.endswith("_ms")})
    if dev is None:
        ref = {k: np.array(r[k], copy=True) for k in ("dsurf", "rw", "row", "col")}
    else:
        print("identical to single device:", all(np.array_equal(ref[k], r[k]) for k in ref))
PY
cat gpurun_out/r2h_multi_timing.txt | tail -5
