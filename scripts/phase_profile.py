"""Per-phase cycle breakdown of the CEC2014 stage kernel (debug aid; run on the GPU box)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import oracle  # noqa: E402  (synthetic tables only)
from pagmo2_b200 import capi  # noqa: E402

n, dim = 1 << 20, 100
ctx = capi.Context(0)
O = oracle()
rng = np.random.default_rng(1)
d_x = ctx.to_device(rng.uniform(-100, 100, (n, dim)))
d_f = ctx.malloc(8 * n)
out = {}
for f in [int(a) for a in sys.argv[1:]] or [1, 2, 5, 6, 8, 11, 12, 17, 23, 27]:
    mr, os_c, s = O.cec2014_problem_tables(f, dim)
    p = capi.Problem(ctx, "cec2014", prob_id=f, dim=dim, rotation=mr, shift=os_c, shuffle=s)
    p.phase_cycles(d_x, n, d_f)
    c = p.phase_cycles(d_x, n, d_f)
    t = c.pop("warp_tiles")
    out[f] = {k: round(v / t) for k, v in c.items()}
    out[f]["total"] = sum(out[f].values())
    print(f, out[f], flush=True)
    p.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "phase_cycles.json").write_text(json.dumps(out, indent=1))
