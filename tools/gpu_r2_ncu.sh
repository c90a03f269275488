#!/bin/bash
# ncu visit (1 GPU): launch list of a short bench run + one --set full capture of the hot kernels.
# (ncu on the box first runs the command plain; keep the command short.)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-oracle"
$CMD > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r2_ncu_launches.log
ncu --set full --clock-control none --import-source on \
    -k regex:'mlp_chain2|mlp_dw2|sample1_kernel|sample1_bwd|feat_kernel|feat_bwd|march_kernel|composite_kernel|adam_bricks_kernel|adam_dense_zero|smooth|tv_kernel|weights1' \
    -s 65 -c 14 -o gpurun_out/r2_full $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/r2_full.ncu-rep
