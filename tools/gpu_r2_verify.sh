#!/bin/bash
# Round-2 verification visit (1 GPU): the GPU suite three times (the multi-iteration comparisons run in lockstep since a
# free-running one was measured flaky, profiles/r02_flaky_graph_test.log), smoke(), the headline bench line.
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest tests -q -m gpu -x -rs > gpurun_out/r2v_pytest_$i.log 2>&1; echo "pytest $i rc=$?"; tail -2 gpurun_out/r2v_pytest_$i.log
done
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2v_smoke.log
timeout 400 python bench.py > gpurun_out/r2v_bench_cfg2.json 2> gpurun_out/r2v_bench_cfg2.err; echo "cfg2 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2v_bench_cfg2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], [ (k['name'],k['ms_per_step']) for k in d['kernels'] if k['name'] in ('fine march','fine feat')])"
