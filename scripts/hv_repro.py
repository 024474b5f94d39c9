import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(63)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
f = rng.uniform(0, 1, (n, 3))
f = f / np.linalg.norm(f, axis=1, keepdims=True)
c = ctx.hv_contributions(f, np.full(3, 1.25))
print("ok", c.sum())
