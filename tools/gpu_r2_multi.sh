#!/bin/bash
# Round-2 multi-GPU visit (N = number of GPUs of the lease, 2 / 4 / 8): BASELINE config 2 (160^3), config 4 (256^3,
# 8192 rays per rank) and config 5 (384^3 full-image render) at N ranks.  usage: bash tools/gpu_r2_multi.sh N
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
timeout 300 $RUN --steps 30 --warmup 5 > gpurun_out/r2_final_bench_cfg2_n$N.json 2> gpurun_out/r2_final_bench_cfg2_n$N.err; echo "cfg2 n$N rc=$?"
timeout 400 $RUN --steps 30 --warmup 5 --grid 256 > gpurun_out/r2_final_bench_cfg4_n$N.json 2> gpurun_out/r2_final_bench_cfg4_n$N.err; echo "cfg4 n$N rc=$?"
timeout 400 $RUN --mode render --grid 384 --steps 5 --warmup 2 > gpurun_out/r2_final_bench_cfg5_n$N.json 2> gpurun_out/r2_final_bench_cfg5_n$N.err; echo "cfg5 n$N rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final_bench_*_n$N.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = {k["name"]: k["ms_per_step"] for k in d.get("kernels", [])}
        print(f, d.get("metric"), round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "allreduce", ks.get("allreduce"), "dw", ks.get("mlp_dw"))
    except Exception as e:
        print(f, "unreadable", e)
PY
