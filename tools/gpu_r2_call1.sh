#!/bin/bash
# Round 2, first GPU visit: parity suite, smoke, bench at configs 2 / 3 / 5 (1 GPU), phase probe of the MLP kernels.
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rs 2>&1 | tail -60 > gpurun_out/r2_pytest1.log
tail -15 gpurun_out/r2_pytest1.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke1.log 2>&1; tail -3 gpurun_out/r2_smoke1.log
python bench.py > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "bench cfg2 rc=$?"; tail -3 gpurun_out/r2_bench_cfg2.err
python bench.py --grid 256 --camera dtu --no-cpu-baseline --no-gpu-oracle > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo "bench cfg3 rc=$?"; tail -3 gpurun_out/r2_bench_cfg3.err
python bench.py --mode render --grid 384 --steps 3 --warmup 1 > gpurun_out/r2_bench_cfg5_n1.json 2> gpurun_out/r2_bench_cfg5_n1.err; echo "bench cfg5 rc=$?"; tail -3 gpurun_out/r2_bench_cfg5_n1.err
python tools/chain_phase_probe.py > gpurun_out/r2_chain_probe.log 2>&1; tail -40 gpurun_out/r2_chain_probe.log
python - <<'PY'
import json
for f in ("r2_bench_cfg2", "r2_bench_cfg3", "r2_bench_cfg5_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_oracle"), d["config"].get("samples_per_ray"))
        for k in d.get("kernels", [])[:30]:
            print("   ", k["name"], k["ms_per_step"], k.get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
