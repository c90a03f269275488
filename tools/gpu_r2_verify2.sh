#!/bin/bash
# Round-2 verification on 2 GPUs of the final code: the data-parallel parity tests and the config-2 line at 2 ranks.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -q -m gpu -x -s > gpurun_out/r2v_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2v_pytest_2gpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2v_bench_cfg2_n2.json 2> gpurun_out/r2v_bench_cfg2_n2.err; echo "cfg2 n2 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2v_bench_cfg2_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
