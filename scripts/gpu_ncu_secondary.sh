#!/bin/bash
# One gpurun call: `ncu --set full` captures of the secondary kernels (one launch each) -> gpurun_out/<tag>_sec_<name>.ncu-rep
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_secondary.sh r1'
TAG=${1:-r1}
mkdir -p gpurun_out
cap() { # name, kernel regex, launches to skip, workload
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${TAG}_sec_$1 \
        python scripts/run_secondary_kernels.py $4 > gpurun_out/${TAG}_sec_$1.log 2>&1
    echo "$1 rc=$?"
}
cap cec13_rot cec13_kernel 6 cec13
cap lj lj_kernel 2 lj
cap fnds_count fnds_count 1 fnds
cap fnds_peel fnds_peel 40 fnds
cap gram gram_partial 1 gram
cap sample cmaes_sample 1 sample
cap de_trial de_trial 2 de
cap hv_sweep hv_sweep 0 hv
ls -la gpurun_out/${TAG}_sec_*.ncu-rep
