#!/bin/bash
# Round-2 final 1-GPU visit: the GPU test suite, smoke(), and one bench line per single-GPU configuration
# (BASELINE configs 2, 3, 5 at N=1, the coarse stage, the reference arm).  Everything lands in gpurun_out/r2_final_*.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py > gpurun_out/r2_final_bench_cfg2.json 2> gpurun_out/r2_final_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --grid 256 --camera dtu --no-cpu-baseline --no-gpu-oracle > gpurun_out/r2_final_bench_cfg3.json 2> gpurun_out/r2_final_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 300 python bench.py --mode render --grid 384 --steps 5 --warmup 2 > gpurun_out/r2_final_bench_cfg5_n1.json 2> gpurun_out/r2_final_bench_cfg5_n1.err; echo "cfg5 rc=$?"
timeout 300 python bench.py --stage coarse --grid 116 --global-step 3000 > gpurun_out/r2_final_bench_coarse.json 2> gpurun_out/r2_final_bench_coarse.err; echo "coarse rc=$?"
timeout 300 python bench.py --stage geometry_searching --grid 120 --global-step 600 > gpurun_out/r2_final_bench_geometry.json 2> gpurun_out/r2_final_bench_geometry.err; echo "geometry rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "reference rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("metric"), round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), (d.get("gpu_oracle") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
