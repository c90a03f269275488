#!/bin/bash
# ncu visit for the coarse stage: --set full of the stage's own kernels during `bench.py --stage coarse`.
mkdir -p gpurun_out
CMD="python bench.py --stage coarse --grid 116 --global-step 3000 --steps 3 --warmup 3"
$CMD > gpurun_out/r2_ncu_coarse_plain.json 2> gpurun_out/r2_ncu_coarse_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_coarse_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on \
    -k regex:'coarse_|conv3d_replicate|l1_tv|tv_grad_grid|tv_scatter' -s 40 -c 11 -o /tmp/r2_coarse $CMD > gpurun_out/r2_ncu_coarse.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/r2_ncu_coarse.log
python tools/ncu_digest.py /tmp/r2_coarse.ncu-rep r02_coarse > gpurun_out/r2_ncu_coarse_digest.log 2>&1; echo "digest rc=$?"
cp profiles/r02_coarse_ncu_full.txt gpurun_out/ 2>/dev/null
