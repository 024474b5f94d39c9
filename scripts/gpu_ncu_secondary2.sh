#!/bin/bash
# `ncu --set full` captures (one launch each) of the kernels added or rewritten in r1l-r1s -> gpurun_out/<tag>_sec_<name>.ncu-rep
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_secondary2.sh r1w'; then python scripts/summarize_ncu.py r1w
TAG=${1:-r1w}
mkdir -p gpurun_out
cap() { # name, kernel regex, launches to skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${TAG}_sec_$name "$@" > gpurun_out/${TAG}_sec_$name.log 2>&1
    echo "$name rc=$?"
}
cap sep_f8_rastrigin cec14_sep 1 python scripts/run_cec14.py 8 --reps 1
cap sep_f10_schwefel cec14_sep 1 python scripts/run_cec14.py 10 --reps 1
cap stage_f1_pure_rotation cec14_stage 1 python scripts/run_cec14.py 1 --reps 1
cap fnds_count_sorted fnds_count_sorted 1 python scripts/run_secondary_kernels.py fnds
cap fnds_peel_sorted fnds_peel_sorted 40 python scripts/run_secondary_kernels.py fnds
cap lj_reg lj_reg_kernel 2 python scripts/run_secondary_kernels.py lj
ls -la gpurun_out/${TAG}_sec_*.ncu-rep
