python -m pytest tests/test_gpu_eval.py -m gpu -q -x 2>&1 | tail -2
python scripts/run_cec14.py 8 10 23 24 1 --reps 10 | tr '\n' ' '; echo
python - <<PY
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in r.items():
    if isinstance(v,dict) and 'cec2014' in k:
        m=max(v.items(),key=lambda t:t[1]); print(k,m)
PY
