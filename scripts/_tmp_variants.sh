python -m pytest tests/test_gpu_eval.py tests/test_gpu_cec2013.py -m gpu -q -x 2>&1 | tail -2
python scripts/run_cec14.py 6 12 22 --reps 10 | tr '\n' ' '; echo
python - <<PY
import sys, time, numpy as np
sys.path.insert(0,'.')
from pagmo2_b200 import capi, synth
ctx=capi.Context(0)
n=1<<20
mr,os_=synth.cec2013_tables(50)
x=ctx.to_device(np.random.default_rng(1).uniform(-100,100,(n,50))); f=ctx.malloc(8*n)
for func in (9,16,8,11):
    p=capi.Problem(ctx,"cec2013",prob_id=func,dim=50,rotation=mr,shift=os_)
    p.eval_device(x,n,f); ctx.synchronize(); t0=time.perf_counter()
    for _ in range(5): p.eval_device(x,n,f)
    ctx.synchronize(); dt=(time.perf_counter()-t0)/5
    print("cec2013 f%d D=50: %.3f ms  %.3g evals/s"%(func,dt*1e3,n/dt)); p.close()
PY
