#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench (both arms), ncu launch list + one full capture.
# Usage (from the authoring container):  gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r1}
mkdir -p gpurun_out
{ nvidia-smi; nproc; free -g | head -2; } > gpurun_out/${TAG}_box.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cec14_stage -s 54 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | head -30
head -c 1500 gpurun_out/${TAG}_bench.json
