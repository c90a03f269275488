#!/bin/bash
# ncu visit after the march / feat change (1 GPU): the launch list of a short bench run again, and a --set full capture
# of the two kernels that changed (the other 20 kernels of the step are unchanged since profiles/r02_ncu_full.txt).
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-oracle"
$CMD > gpurun_out/r2d_ncu_plain.json 2> gpurun_out/r2d_ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2d_ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2d_launches.csv $CMD > gpurun_out/r2d_ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r2d_ncu_launches.log
ncu --set full --clock-control none --import-source on -k regex:'march_kernel|feat_kernel' -s 12 -c 2 -o /tmp/r2d_full $CMD > gpurun_out/r2d_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/r2d_ncu_full.log; ls -la /tmp/r2d_full.ncu-rep
cp /tmp/r2d_full.ncu-rep gpurun_out/r2d_full.ncu-rep
