#!/bin/bash
# Fast iteration on the headline kernel only: CEC2014 parity tests, bench (no e2e / cpu legs), phase profile.
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_eval.py tests/test_gpu_cec2013.py -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python scripts/phase_profile.py 1 6 8 10 12 > gpurun_out/${TAG}_phase.log 2>&1; cat gpurun_out/${TAG}_phase.log
python - <<PY
import json
b=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.4g ms/step %.2f frac %.3f rot %.3f'%(b['value'],b['ms_per_step'],b['roofline']['frac'],b['roofline']['rotation_only_frac']))
print(' '.join('f%d:%.3f'%(p['f'],p['ms']) for p in b['per_function']))
PY
