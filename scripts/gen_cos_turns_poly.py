"""Coefficients of cos(2 pi f) = sum_k c_k (f^2)^k on |f| <= 1/4 (Chebyshev interpolation in u = f^2, 40-digit arithmetic),
and a float64 Horner check against mpmath.  Output pasted into pagmo2_b200/csrc/cec_device.cuh (cos_turns)."""
import mpmath as mp
import numpy as np

mp.mp.dps = 50
N = 8  # degree in u
a, b = mp.mpf(0), mp.mpf(1) / 16
nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * i + 1) / (2 * (N + 1))) for i in range(N + 1)]
vals = [mp.cos(2 * mp.pi * mp.sqrt(u)) for u in nodes]
V = mp.matrix(N + 1, N + 1)
for i, u in enumerate(nodes):
    for k in range(N + 1):
        V[i, k] = u ** k
c = mp.lu_solve(V, mp.matrix(vals))
coef = [float(x) for x in c]
print("coefficients:")
for k, x in enumerate(coef):
    print(f"  c{k} = {x!r}  ({x.hex()})")
rng = np.random.default_rng(0)
f = np.concatenate([rng.uniform(-0.25, 0.25, 200000), np.linspace(-0.25, 0.25, 20001)])
u = f * f
p = np.full_like(u, coef[N])
for k in range(N - 1, -1, -1):
    p = p * u + coef[k]  # numpy: separate roundings (worse than fma)
truth = np.array([float(mp.cos(2 * mp.pi * mp.mpf(x))) for x in f[::50]])
print("max abs err (float64 Horner, no fma):", np.max(np.abs(p[::50] - truth)))
