#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list (shares only).
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
