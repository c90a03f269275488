#!/bin/bash
# Round-2 refresh of the other single-GPU lines with the final build (config 3, config 5 at N=1, the two coarse stages).
mkdir -p gpurun_out
timeout 120 python bench.py --grid 256 --camera dtu --no-cpu-baseline --no-gpu-oracle > gpurun_out/r2v_bench_cfg3.json 2> gpurun_out/r2v_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 120 python bench.py --stage coarse --grid 116 --global-step 3000 > gpurun_out/r2v_bench_coarse.json 2> gpurun_out/r2v_bench_coarse.err; echo "coarse rc=$?"
timeout 120 python bench.py --stage geometry_searching --grid 120 --global-step 600 > gpurun_out/r2v_bench_geometry.json 2> gpurun_out/r2v_bench_geometry.err; echo "geometry rc=$?"
timeout 120 python bench.py --mode render --grid 384 --steps 5 --warmup 2 > gpurun_out/r2v_bench_cfg5_n1.json 2> gpurun_out/r2v_bench_cfg5_n1.err; echo "cfg5 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2v_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
