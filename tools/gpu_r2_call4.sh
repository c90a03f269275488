#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_fine_gpu.py -m gpu -x -v --tb=short -p no:cacheprovider -k "elementwise or 3xbf16 or config5 or brick_adam" > gpurun_out/r2_pytest_a.log 2>&1
echo "pytest a rc=$?"; grep -n "FAILED\|ERROR\|passed\|failed\|Aborted\|Fatal\|Error" gpurun_out/r2_pytest_a.log | tail -12
python -m pytest tests/test_training_gpu.py -m gpu -v --tb=short -p no:cacheprovider > gpurun_out/r2_pytest_b.log 2>&1
echo "pytest b rc=$?"; grep -n "FAILED\|ERROR\|passed\|failed\|Aborted\|Fatal\|Error" gpurun_out/r2_pytest_b.log | tail -12
python bench.py --no-cpu-baseline --no-gpu-oracle > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_c.err
python tools/chain_phase_probe.py > gpurun_out/r2_chain_probe_c.log 2>&1; tail -52 gpurun_out/r2_chain_probe_c.log
python - <<'PY'
import json
for f in ("r2_bench_c",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("samples_per_ray"))
        for k in d.get("kernels", [])[:30]:
            print("   ", k["name"], k["ms_per_step"], k.get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
