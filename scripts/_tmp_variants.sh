python -m pytest tests/test_gpu_mo_utils.py tests/test_gpu_nsga2.py tests/test_gpu_migration.py -m gpu -q -x 2>&1 | tail -4
for f in 1 0; do echo "== fuse $f"
PGC_FNDS_FUSE=$f python scripts/bench_mo.py 65536 2>&1 | grep -E "^zdt|^dtlz" | python -c "
import sys,json
for l in sys.stdin:
    k,_,j=l.partition(' '); d=json.loads(j); print(k, 'fnds_N %.2f fnds_2N %.2f select %.2f'%(d['fnds_N_ms'],d['fnds_2N_ms'],d['select_best_2N_to_N_ms']))"
PGC_FNDS_FUSE=$f python scripts/bench_nsga2.py 65536 2>&1 | grep -oE '"generations_per_s": [0-9.]+|"launches_per_generation": [0-9.]+' | tr '\n' ' '; echo
done
