#!/bin/bash
# ncu evidence for the current bench command: (1) launch list with device times, (2) full capture
# of the kernel matching $1 (default gemm_kernel).  Each ncu run directly follows a plain run.
PAT=${1:-gemm_kernel}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$PAT -s 40 -c 3 -o gpurun_out/prof_$PAT $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
