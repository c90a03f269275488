#!/bin/bash
# ncu visit (1 GPU): launch list of a short bench run + one --set full capture of every kernel of a training step.
# (each ncu pass only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-oracle"
$CMD > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r2_ncu_launches.log
# steps 0..5 launch ~21-25 library kernels each (TV steps: +4); skip the first five steps, capture two whole steps
ncu --set full --clock-control none --import-source on \
    -k regex:'mlp_chain2|mlp_dw2|mlp_bwd_planes|dw_small|weight_split|march_kernel|expand1|sample1|weights1|compact2|feat_kernel|feat_bwd|composite_kernel|adam_|brick_ulist|tv_' \
    -s 115 -c 50 -o gpurun_out/r2_full $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/r2_full.ncu-rep
