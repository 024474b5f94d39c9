mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cec14_sep -c 1 -f -o gpurun_out/r1m_sep_f8 python scripts/run_cec14.py 8 --reps 1 > gpurun_out/r1m_ncu_sep.log 2>&1
tail -3 gpurun_out/r1m_ncu_sep.log
python scripts/run_cec14.py 8 10 23 24 --reps 5
