#!/bin/bash
# Evidence run for profiles/ (one gpurun call, one GPU):  bash tools/collect_profiles.sh <tag>
#   1. the bench command exits 0 WITHOUT ncu, 2. ncu launch list of the same command, 3. DRAM / L2 / tensor-pipe table of
#   the conv launches of one DDIM step, 4. `--set full` of one whole UNet forward (raw page exported to csv on the box:
#   the .ncu-rep itself is too large to travel), 5. `--set full --import-source on` of the head kernel and of the
#   256 px conv1 / conv2 (kept as .ncu-rep for the source page).
set -u
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 3 --ddim-steps 2 --no-cpu-baseline --no-extras"
OUT=gpurun_out
K='conv_igemm|head_conv|gn_apply|gn_affine|stem_im2col|ddim_step|linear_kernel|cond_combine'
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_a_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
    --clock-control none -k regex:"conv_igemm|head_conv" -s 216 -c 36 --csv --log-file $OUT/convs_$TAG.csv $CMD > $OUT/ncu_b_$TAG.log 2>&1
[ "${FULL:-1}" = "0" ] && { ls -la $OUT/*_$TAG* | tail -6; exit 0; }   # FULL=0: launch list + conv table only (2 min instead of 13)
# one whole DDIM step (conditioning, stem, 35 convs, 28 GroupNorm applies, gn_affine, head incl. the DDIM update = 68 kernels
# of the filter): the 7th step of the run starts at filtered launch 414 (2 steps per sample() = 137 filtered launches)
ncu --set full --clock-control none -k regex:"$K" -s 414 -c 68 -f -o /tmp/full_$TAG $CMD > $OUT/ncu_c_$TAG.log 2>&1
ncu -i /tmp/full_$TAG.ncu-rep --page raw --csv > $OUT/full_$TAG.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"head_conv" -s 6 -c 1 -f -o $OUT/prof_${TAG}_head $CMD > $OUT/ncu_d_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 217 -c 2 -f -o $OUT/prof_${TAG}_conv256 $CMD > $OUT/ncu_e_$TAG.log 2>&1
ls -la $OUT/*_$TAG* | tail -12
