for slots in 2 4; do for ti in 0 2 4 8; do
  if [ $ti = 0 ]; then unset PGC_CEC13_TI; else export PGC_CEC13_TI=$ti; fi
  export PGC_CEC13_SHARE_SLOTS=$slots PARTS_MODES=eval_cec2013,loop_cec2013 PARTS_KS=8
  echo -n "slots=$slots ti=$ti: "; python scripts/concurrent_parts.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print({m: {k: round(x['us_per_island_step'],1) for k,x in v.items()} for m,v in d.items()})"
done; done
