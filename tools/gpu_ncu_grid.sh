#!/bin/bash
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"level1_kernel|feat_kernel|feat_bwd_kernel|level1_bwd_kernel|march_kernel|composite_kernel|adam_bricks_kernel" -s 21 -c 7 -f -o gpurun_out/prof_grid_kernels $CMD > gpurun_out/ncu4.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu4.log
