#!/bin/bash
# The last GPU call of round 1 (~4.9 GPU-minutes left): new tests first, then the stage-op rooflines, then the
# rest of the GPU suite as a regression check of the engine's learning-rate plumbing, then smoke().
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 120 python -m pytest tests/test_training_gpu.py tests/test_dropin_gpu.py -m gpu -q -s -p no:cacheprovider \
    > gpurun_out/r1_new_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r1_new_tests.log )
tail -5 gpurun_out/r1_new_tests.log
( timeout 70 python tools/stage_ops_bench.py --out gpurun_out/r1_stage_ops.json > gpurun_out/r1_stage_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r1_stage_ops.log )
tail -8 gpurun_out/r1_stage_ops.log
( timeout 100 python -m pytest tests/test_fine_gpu.py tests/test_native_ops_gpu.py -m gpu -q -x -p no:cacheprovider \
    > gpurun_out/r1_regression.log 2>&1; echo "rc=$?" >> gpurun_out/r1_regression.log )
tail -3 gpurun_out/r1_regression.log
( timeout 60 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r1_smoke.log )
tail -3 gpurun_out/r1_smoke.log
