echo "== default"; python scripts/run_cec14.py 1 2 13 23 5 8 10 --reps 10 | tr '\n' ' '; echo
echo "== sep3 (previous)"; PGC_LIBRARY_PATH=$PWD/pagmo2_b200/_variants/libpgc_sep3.so python scripts/run_cec14.py 1 2 13 23 5 8 10 --reps 10 | tr '\n' ' '; echo
bash scripts/gpu_quick14.sh r1p
