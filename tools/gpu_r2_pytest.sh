#!/bin/bash
# full verbose GPU parity run, stop at the first failure, whole log kept
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -v --tb=short -p no:cacheprovider "$@" > gpurun_out/r2_pytest_full.log 2>&1
echo "pytest rc=$?"
grep -n "PASSED\|FAILED\|ERROR\|passed\|failed" gpurun_out/r2_pytest_full.log | tail -15
