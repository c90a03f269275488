#!/bin/bash
mkdir -p gpurun_out
T=tests/test_training_gpu.py::test_graph_launched_step_equals_directly_enqueued_step
for i in 1 2 3 4 5 6; do timeout 120 python -m pytest $T -x -q 2>&1 | grep -E "passed|failed|assert 0\." | tr '\n' ' '; echo " [base $i]"; done > gpurun_out/flaky.log 2>&1
export FGS_MARCH_W=2 FGS_MARCH_MINB=6
for i in 1 2 3 4 5 6; do timeout 120 python -m pytest $T -x -q 2>&1 | grep -E "passed|failed|assert 0\." | tr '\n' ' '; echo " [march26 $i]"; done >> gpurun_out/flaky.log 2>&1
cat gpurun_out/flaky.log
