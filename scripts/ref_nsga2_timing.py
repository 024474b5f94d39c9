#!/usr/bin/env python
"""Wall time of ONE generation of the unmodified reference nsga2::evolve (oracle/_ref, reference src/algorithms/nsga2.cpp:91-307,
sequential fitness, one thread) at several population sizes, the fitted exponent t ~ N^p, and the extrapolation to BASELINE cfg3's
pop 65 536 - the CPU figure quoted next to the device's NSGA-II generations/s.  Run where /root/reference was compiled (CPU only):

    python scripts/ref_nsga2_timing.py > profiles/r2_ref_nsga2_timing.json
"""
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.pyoracle import reference  # noqa: E402

R = reference()
out = {"what": "reference nsga2::evolve, 1 generation (2 generations timed, halved), defaults cr 0.95 eta_c 10 m 0.01 eta_m 50, one host thread",
       "host": {"cores": os.cpu_count()}, "cases": {}}
for name, fam, args in (("zdt1_nx30", "zdt", (1, 30)), ("dtlz2_nx12_m3", "dtlz", (2, 12, 3, 100))):
    rows = []
    for NP in (2048, 4096, 8192, 16384):
        p = R.problem(fam, *args)
        secs, _, _, fe = p.evolve("nsga2", NP, 2, pop_seed=31, algo_seed=7)
        rows.append({"pop": NP, "seconds_per_generation": secs / 2, "fevals": fe})
        print(name, NP, secs / 2, file=sys.stderr)
    lx, ly = np.log([r["pop"] for r in rows[1:]]), np.log([r["seconds_per_generation"] for r in rows[1:]])
    p_exp, c = np.polyfit(lx, ly, 1)
    t65536 = float(np.exp(c) * 65536 ** p_exp)
    out["cases"][name] = {"rows": rows, "fitted_exponent": float(p_exp), "extrapolated_seconds_per_generation_at_65536": t65536,
                          "extrapolated_generations_per_s_at_65536": 1.0 / t65536}
print(json.dumps(out, indent=1))
