"""cec2013 strict mode against the restated reference on the GPU box: worst relative error per function and dimension in the
default (tensor/tiled rotation) mode and in strict mode (reference accumulation order), and how many points exceed 1e-12."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import oracle  # noqa: E402
from pagmo2_b200 import capi  # noqa: E402

orc = oracle()
ctx = capi.Context(0)
out = {}
for dim in (10, 30, 50, 100):
    rng = np.random.default_rng(1400 + dim)
    mr, os_ = orc.cec2013_tables(dim)
    n = 403
    xs = np.vstack([rng.uniform(-100, 100, (n - 103, dim)), os_[:dim] + rng.normal(0, 1.0, (100, dim)), os_[None, :dim], np.zeros((1, dim)),
                    rng.uniform(-100, 100, (1, dim))])
    for func in range(1, 29):
        prob = capi.Problem(ctx, "cec2013", prob_id=func, dim=dim, rotation=mr, shift=os_)
        want = orc.cec2013(func, xs)
        loose = prob.eval_host(xs)[:, 0]
        prob.set_strict(True)
        got = prob.eval_host(xs)[:, 0]
        prob.close()
        scale = 4.189828872724338e+002 * dim if func in (23, 28) else 0.0
        den = np.maximum(np.abs(want), scale)
        rs, rl = np.abs(got - want) / den, np.abs(loose - want) / den
        out[f"d{dim}_f{func}"] = {"strict_max": float(rs.max()), "strict_over": int((rs > 1e-12).sum()), "loose_max": float(rl.max()),
                                  "loose_over": int((rl > 1e-12).sum()), "argmax": int(rs.argmax())}
        if rs.max() > 1e-12:
            print(dim, func, out[f"d{dim}_f{func}"])
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "cec2013_strict_report.json").write_text(json.dumps(out, indent=1))
