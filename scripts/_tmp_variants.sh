echo "== default"; python scripts/run_cec14.py 1 2 13 23 5 26 --reps 10 | tr '\n' ' '; echo
for v in s0 s05 s15 s2 s4; do
echo "== $v"; PGC_LIBRARY_PATH=$PWD/pagmo2_b200/_variants/libpgc_$v.so python scripts/run_cec14.py 1 2 13 23 5 26 --reps 10 | tr '\n' ' '; echo
done
