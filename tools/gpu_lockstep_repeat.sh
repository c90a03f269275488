#!/bin/bash
# flakiness check of the final tests: tests/test_training_gpu.py (every multi-iteration comparison of two step drivers
# lives there) repeated in one process for ~20 s
mkdir -p gpurun_out
timeout 40 python - > gpurun_out/r2x_training_repeat.log 2>&1 <<'PY'
import sys, time, pytest
t0 = time.time()
res = []
while time.time() - t0 < 20.0:
    rc = pytest.main(["-q", "-x", "-p", "no:cacheprovider", "-m", "gpu", "tests/test_training_gpu.py"])
    res.append(int(rc))
    print("repeat", len(res), "rc", int(rc), "t", round(time.time() - t0, 1), flush=True)
    if rc != 0:
        break
print("SUMMARY runs", len(res), "failures", sum(1 for r in res if r != 0))
PY
tail -2 gpurun_out/r2x_training_repeat.log
