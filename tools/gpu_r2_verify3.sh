#!/bin/bash
# one pass of the GPU suite + the headline line (after the tiled weight_split_kernel and the 16-row dw_small_kernel)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x -rs > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2x_pytest.log
timeout 120 python bench.py > gpurun_out/r2x_bench_cfg2.json 2> gpurun_out/r2x_bench_cfg2.err; echo "cfg2 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2x_bench_cfg2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], [ (k['name'],k['ms_per_step']) for k in d['kernels'] if k['name'] in ('weight_split','mlp_chain_fwd','mlp_dw_small')])"
