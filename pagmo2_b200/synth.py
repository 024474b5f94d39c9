"""Seeded synthetic CEC2013 / CEC2014 data tables for benchmarks and examples (numpy only).

The reference's real tables (src/problems/cec2014_data.cpp, cec2013_data.cpp) are not part of the checkout this repository was
built against, so workloads use tables of the same shape and statistics: orthogonal rotation matrices (QR of a Gaussian matrix),
shifts ~ U[-80, 80), random shuffles of 1..dim.  Layouts are those of the reference members (see include/pagmo_cuda/pgc.h):
rotation component i at i*dim*dim, shift component i at i*dim, shuffle component i at i*dim (1-based).
These are NOT the tables of oracle/cec_synth.c (the tests use those on both sides); throughput does not depend on the values."""
from __future__ import annotations

import numpy as np

NCOMP = 10


def _rotations(rng, dim: int, k: int = NCOMP) -> np.ndarray:
    out = np.empty((k, dim, dim))
    for i in range(k):
        q, r = np.linalg.qr(rng.standard_normal((dim, dim)))
        out[i] = q * np.sign(np.diag(r))  # fix the signs so that the distribution is Haar
    return out.reshape(-1)


def cec2014_tables(func: int, dim: int, seed: int = 2014):
    """(rotation [10*dim*dim], shift [10*dim] compacted as the cec2014 ctor does, shuffle [10*dim] 1-based)."""
    rng = np.random.default_rng([seed, func, dim])
    shuffle = np.concatenate([rng.permutation(dim) + 1 for _ in range(NCOMP)]).astype(np.int32)
    return _rotations(rng, dim), rng.uniform(-80.0, 80.0, NCOMP * dim), shuffle


def cec2013_tables(dim: int, seed: int = 2013):
    """(rotation = MD[dim]: 10 matrices, shift_data: 10 lines of 100 values, addressed at i*dim by the reference)."""
    rng = np.random.default_rng([seed, dim])
    return _rotations(rng, dim), rng.uniform(-80.0, 80.0, NCOMP * 100)
