#!/bin/bash
# ncu evidence for the current bench command: (1) launch list with device times, (2) full capture
# of the kernels matching $1 (default mlp_chain_kernel) and $2 (default mlp_dw_kernel).  Each ncu run
# directly follows a plain run of the same command.
PAT=${1:-mlp_chain2_kernel}
PAT2=${2:-mlp_dw_kernel}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$PAT -s 6 -c 2 -f -o gpurun_out/prof_$PAT $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture 1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$PAT2 -s 3 -c 1 -f -o gpurun_out/prof_$PAT2 $CMD > gpurun_out/ncu3.log 2>&1
echo "full capture 2 rc=$?"
ls -la gpurun_out | tail -12
