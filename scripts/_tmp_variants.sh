python -m pytest tests/test_gpu_mo_utils.py tests/test_gpu_nsga2.py tests/test_gpu_migration.py -m gpu -q -x 2>&1 | tail -4
PGC_FNDS_COOP=0 python -m pytest tests/test_gpu_mo_utils.py tests/test_gpu_nsga2.py -m gpu -q -x 2>&1 | tail -2
echo "== coop"; python scripts/bench_mo.py 65536 2>&1 | tail -30 | grep -E "^zdt|^dtlz" | cut -c1-420
python scripts/bench_nsga2.py 65536 2>&1 | grep -E "generations_per_s|launches"
echo "== launch per level"; PGC_FNDS_COOP=0 python scripts/bench_mo.py 65536 2>&1 | tail -30 | grep -E "^zdt|^dtlz" | cut -c1-420
PGC_FNDS_COOP=0 python scripts/bench_nsga2.py 65536 2>&1 | grep -E "generations_per_s|launches"
