#!/bin/bash
# cfg5 on one GPU for forced tile sizes of the island-sized cec2013 launches (PGC_CEC13_TI) - experiment
for t in 1 2 4 8; do
  PGC_CEC13_TI=$t python bench.py --workload cfg5 --steps 6 --warmup 2 2>/dev/null | tail -1 > /tmp/cfg5_$t.json
  python -c "import json; d=json.load(open('/tmp/cfg5_$t.json')); print('ti', $t, round(d['value']), d.get('ms_per_step'))"
done
