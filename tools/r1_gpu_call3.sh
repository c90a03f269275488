#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 60 python tools/trainer_pool_probe.py gpurun_out/r1_pool_probe.json > gpurun_out/r1_pool_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r1_pool_probe.log )
tail -5 gpurun_out/r1_pool_probe.log
( timeout 100 python bench.py --steps 30 --warmup 5 > gpurun_out/r1_bench_native.json 2> gpurun_out/r1_bench_native.err; echo "rc=$?" >> gpurun_out/r1_bench_native.err )
tail -c 1500 gpurun_out/r1_bench_native.json; tail -3 gpurun_out/r1_bench_native.err
