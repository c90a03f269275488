#!/bin/bash
# 2-GPU visit: data-parallel parity tests (log kept under profiles/) and the N = 2 bench line
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -v -s --tb=short -p no:cacheprovider > gpurun_out/r2_pytest_2gpu.log 2>&1
echo "pytest rc=$?"; grep -n "PASSED\|FAILED\|ERROR\|passed\|failed\|skipped" gpurun_out/r2_pytest_2gpu.log | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k in d.get("kernels", [])[:14]:
    print("   ", k["name"], k["ms_per_step"], k.get("frac"))
PY
