#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -v --tb=short -p no:cacheprovider > gpurun_out/r2_pytest_full.log 2>&1
echo "pytest rc=$?"
grep -n "FAILED\|ERROR\|passed\|failed\|Aborted\|Fatal" gpurun_out/r2_pytest_full.log | tail -15
python bench.py --no-cpu-baseline --no-gpu-oracle > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_b.err
python tools/chain_phase_probe.py > gpurun_out/r2_chain_probe_b.log 2>&1; tail -40 gpurun_out/r2_chain_probe_b.log
python - <<'PY'
import json
for f in ("r2_bench_b",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("samples_per_ray"))
        for k in d.get("kernels", [])[:30]:
            print("   ", k["name"], k["ms_per_step"], k.get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
