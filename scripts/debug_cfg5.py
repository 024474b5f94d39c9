"""Small cfg5-like archipelago for compute-sanitizer runs: 4 islands x sade on cec2013 f12 D=50, pop 256, 3 rounds of 6 generations."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi, synth  # noqa: E402
from pagmo2_b200.archipelago import ResidentArchipelago  # noqa: E402

mr, os_ = synth.cec2013_tables(50)
spec = [dict(device=0, family="cec2013", problem_kw=dict(prob_id=12, dim=50, rotation=mr, shift=os_),
             algo=capi.algo_desc("sade", gens=6, seed=11 + g, ftol=0.0, xtol=0.0), pop_size=256, seed=200 + g, r_rate=1, s_rate=1, owner=0)
        for g in range(4)]
a = ResidentArchipelago(spec, topology="ring", seed=1, comm=None, my_ranks=(0,), log=True)
a.evolve(3)
a.synchronize()
print("champions", a.champions_f().tolist())
