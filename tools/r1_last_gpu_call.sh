#!/bin/bash
# GPU calls at the end of round 1 (a few GPU-minutes left): new tests first, then the stage-op rooflines, then the
# rest of the GPU suite as a regression check, then smoke().  Every leg has its own timeout.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 100 python -m pytest tests/test_training_gpu.py tests/test_dropin_gpu.py -m gpu -q -p no:cacheprovider \
    > gpurun_out/r1_new_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r1_new_tests.log )
tail -5 gpurun_out/r1_new_tests.log
( timeout 60 python tools/stage_ops_bench.py --out gpurun_out/r1_stage_ops.json > gpurun_out/r1_stage_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r1_stage_ops.log )
tail -8 gpurun_out/r1_stage_ops.log
( timeout 80 python -m pytest tests/test_fine_gpu.py tests/test_native_ops_gpu.py -m gpu -q -x -p no:cacheprovider \
    > gpurun_out/r1_regression.log 2>&1; echo "rc=$?" >> gpurun_out/r1_regression.log )
tail -3 gpurun_out/r1_regression.log
