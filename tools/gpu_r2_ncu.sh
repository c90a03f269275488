#!/bin/bash
# ncu visit (1 GPU): launch list of a short bench run + one --set full capture of every kernel of a training step.
# (each ncu pass only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values)
# The digest is made ON the box (the report of ~26 kernels with source is ~40 MB; gpurun_out/ comes back only below 64 MiB).
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-oracle"
$CMD > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r2_ncu_launches.log
# steps launch 21 library kernels (+4 on TV steps: global steps 5001, 5004, 5007 ...); skip seven steps (7 x 21 + 2 x 4 = 155 launches), capture step 7 = a TV step (25 launches) + margin
ncu --set full --clock-control none --import-source on \
    -k regex:'mlp_chain2|mlp_dw2|mlp_bwd_planes|dw_small|weight_split|march_kernel|expand1|sample1|weights1|compact2|feat_kernel|feat_bwd|composite_kernel|adam_|brick_ulist|tv_' \
    -s 155 -c 27 -o /tmp/r2_full $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/r2_ncu_full.log; ls -la /tmp/r2_full.ncu-rep
python tools/ncu_digest.py /tmp/r2_full.ncu-rep r02 > gpurun_out/r2_ncu_digest.log 2>&1; echo "digest rc=$?"
cp profiles/r02_ncu_full.txt profiles/r02_ncu_traffic.json gpurun_out/
ncu -i /tmp/r2_full.ncu-rep --page details --csv 2>/dev/null | grep -iE "mlp_chain2|mlp_bwd_planes|mlp_dw2" | grep -iE "Tensor|Warp Cycles Per Issued|Stall" | head -60 > gpurun_out/r2_ncu_mlp_details.csv
sz=$(stat -c %s /tmp/r2_full.ncu-rep); if [ "$sz" -lt 45000000 ]; then cp /tmp/r2_full.ncu-rep gpurun_out/r2_full.ncu-rep; else echo "report $sz bytes: kept on the box only"; fi
