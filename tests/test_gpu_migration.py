"""GPU parity of the migration path: device select_best / fair_replace / population init against the restated oracle (bit-exact:
index and copy work), and a whole device archipelago against the same archipelago on oracle-backed islands."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parent))


@pytest.fixture(scope="module")
def capi():
    from pagmo2_b200 import capi as m
    return m


def dev_select(capi, ctx, ids, x, f, rate):
    n, nx, nf = x.shape[0], x.shape[1], f.shape[1]
    d = [ctx.to_device(a) for a in (ids, x, f)]
    o = [ctx.malloc(8 * max(n, 1) * w) for w in (1, nx, nf)]
    k = C.c_size_t()
    try:
        capi.check(capi.lib().pgc_select_best_device(ctx._h, d[0], d[1], d[2], n, nx, nf, int(isinstance(rate, float)), float(rate), o[0], o[1],
                                                     o[2], C.byref(k), None))
        ctx.synchronize()
        k = k.value
        return ctx.from_device(o[0], (k,), np.uint64), ctx.from_device(o[1], (k, nx)), ctx.from_device(o[2], (k, nf))
    finally:
        for p in d + o:
            ctx.free(p)


def dev_replace(capi, ctx, ids, x, f, rate, mids, mx, mf):
    n, nx, nf, nm = x.shape[0], x.shape[1], f.shape[1], mx.shape[0]
    d = [ctx.to_device(a) for a in (ids, x, f)]
    m = [ctx.to_device(a) for a in (mids, mx, mf)] if nm else [None, None, None]
    try:
        capi.check(capi.lib().pgc_fair_replace_device(ctx._h, d[0], d[1], d[2], n, nx, nf, int(isinstance(rate, float)), float(rate), m[0], m[1],
                                                      m[2], nm, None))
        ctx.synchronize()
        return ctx.from_device(d[0], (n,), np.uint64), ctx.from_device(d[1], (n, nx)), ctx.from_device(d[2], (n, nf))
    finally:
        for p in d + [q for q in m if q]:
            ctx.free(p)


@pytest.mark.parametrize("nobj", (1, 2, 3))
def test_policies_match_oracle(capi, ctx, orc, nobj):
    rng = np.random.default_rng(40 + nobj)
    for n, nm, rate in ((20, 5, 1), (20, 5, 3), (20, 0, 2), (7, 9, 0.5), (16, 16, 1.0), (5, 3, 0), (1024, 2, 1), (1024, 300, 0.25),
                        (5000, 5000, 1.0)):
        ids = rng.integers(0, 2**63, n, dtype=np.uint64)
        x, f = rng.normal(size=(n, 7)), rng.normal(size=(n, nobj))
        mids = rng.integers(0, 2**63, nm, dtype=np.uint64)
        mx, mf = rng.normal(size=(nm, 7)), rng.normal(size=(nm, nobj))
        if nobj == 1 and n > 6:
            f[3, 0] = np.nan  # NaN sorts last
            f[5, 0] = f[6, 0]  # a tie keeps index order
        for a, b in zip(dev_select(capi, ctx, ids, x, f, rate), orc.select_best(ids, x, f, rate)):
            assert np.array_equal(a, b, equal_nan=True), (n, nm, rate)
        for a, b in zip(dev_replace(capi, ctx, ids, x, f, rate, mids, mx, mf), orc.fair_replace(ids, x, f, rate, mids, mx, mf)):
            assert np.array_equal(a, b, equal_nan=True), (n, nm, rate)
    ids = np.arange(4, dtype=np.uint64)
    with pytest.raises(capi.PgcError):
        dev_select(capi, ctx, ids, np.zeros((4, 2)), np.zeros((4, nobj)), 5)
    with pytest.raises(capi.PgcError):
        dev_replace(capi, ctx, ids, np.zeros((4, 2)), np.zeros((4, nobj)), 1.5, ids, np.zeros((4, 2)), np.zeros((4, nobj)))


@pytest.mark.parametrize("nec,nic", [(2, 0), (0, 3), (2, 2), (1, 4)])
def test_constrained_policies_match_oracle(capi, ctx, orc, nec, nic):
    """the single-objective constrained branches (select_best.cpp:137-152, fair_replace.cpp:158-188; sort_population_con): bit-exact
    rows against the restatement, which tests/test_archipelago.py pins to the compiled reference.  Individuals violate constraints of
    one kind only, where compare_fc's two norms coincide (pgc.h)."""
    from test_archipelago import _constrained_group
    rng = np.random.default_rng(70 + nec + 10 * nic)
    tol = np.concatenate([np.full(nec, 1e-2), np.full(nic, 1e-2)])
    L = capi.lib()
    vp, sz = C.c_void_p, C.c_size_t
    L.pgc_select_best_con_device.argtypes = [vp, vp, vp, vp, sz, sz, sz, sz, vp, C.c_int, C.c_double, vp, vp, vp, C.POINTER(sz), vp]
    L.pgc_fair_replace_con_device.argtypes = [vp, vp, vp, vp, sz, sz, sz, sz, vp, C.c_int, C.c_double, vp, vp, vp, sz, vp]
    L.pgc_sort_population_con_device.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
    for n, nm, rate in ((20, 5, 1), (20, 5, 3), (12, 9, 0.5), (16, 16, 1.0), (1024, 300, 0.25), (4000, 4000, 1.0)):
        nx, nf = 5, 1 + nec + nic
        ids = rng.integers(0, 2**63, n, dtype=np.uint64)
        mids = rng.integers(0, 2**63, nm, dtype=np.uint64)
        x, mx = rng.normal(size=(n, nx)), rng.normal(size=(nm, nx))
        f, mf = _constrained_group(rng, n, nec, nic, False), _constrained_group(rng, nm, nec, nic, False)
        f[5] = f[2]  # a tie keeps the input order
        if n == 1024:  # NaN constraints are never satisfied and make the violation norm NaN (std::max semantics, constrained.hpp:55,73)
            f[17, 1] = np.nan
            f[300, nec + nic] = np.nan
        d = [ctx.to_device(a) for a in (ids, x, f)]
        o = [ctx.malloc(8 * n * w) for w in (1, nx, nf)]
        m = [ctx.to_device(a) for a in (mids, mx, mf)]
        dord = ctx.malloc(4 * n)
        k = sz()
        try:
            capi.check(L.pgc_sort_population_con_device(ctx._h, d[2], n, nec, nic, tol.ctypes.data, dord, None))
            assert np.array_equal(ctx.from_device(dord, (n,), np.uint32), orc.sort_population_con(f, nec, nic, tol))
            if n == 1024:
                continue  # (row comparisons below use array_equal: NaN rows are covered by the order)
            capi.check(L.pgc_select_best_con_device(ctx._h, d[0], d[1], d[2], n, nx, nec, nic, tol.ctypes.data, int(isinstance(rate, float)),
                                                    float(rate), o[0], o[1], o[2], C.byref(k), None))
            got = ctx.from_device(o[0], (k.value,), np.uint64), ctx.from_device(o[1], (k.value, nx)), ctx.from_device(o[2], (k.value, nf))
            for a, b in zip(got, orc.select_best_con(ids, x, f, rate, nec, nic, tol)):
                assert np.array_equal(a, b), (n, rate)
            capi.check(L.pgc_fair_replace_con_device(ctx._h, d[0], d[1], d[2], n, nx, nec, nic, tol.ctypes.data, int(isinstance(rate, float)),
                                                     float(rate), m[0], m[1], m[2], nm, None))
            got = ctx.from_device(d[0], (n,), np.uint64), ctx.from_device(d[1], (n, nx)), ctx.from_device(d[2], (n, nf))
            for a, b in zip(got, orc.fair_replace_con(ids, x, f, rate, mids, mx, mf, nec, nic, tol)):
                assert np.array_equal(a, b), (n, nm, rate)
        finally:
            for p in d + o + m + [dord]:
                ctx.free(p)
    with pytest.raises(capi.PgcError):  # no constraints: the unconstrained entry points are the ones to call
        L.pgc_sort_population_con_device.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
        capi.check(L.pgc_sort_population_con_device(ctx._h, ctx.malloc(64), 4, 0, 0, None, ctx.malloc(64), None))


def test_population_init_matches_oracle(capi, ctx, orc):
    prob = capi.Problem(ctx, "rastrigin", dim=9)
    lb, ub = prob.bounds()
    n = 1000
    dx, df, di = ctx.malloc(8 * n * 9), ctx.malloc(8 * n), ctx.malloc(8 * n)
    capi.check(capi.lib().pgc_population_init_device(prob._h, n, 77, dx, df, di, None))
    ctx.synchronize()
    x, f, ids = ctx.from_device(dx, (n, 9)), ctx.from_device(df, (n,)), ctx.from_device(di, (n,), np.uint64)
    xo, io = orc.population_init(lb, ub, n, 77)
    assert np.array_equal(x, xo) and np.array_equal(ids, io)
    assert np.allclose(f, orc.simple("rastrigin", x), rtol=1e-12)
    assert (x >= lb).all() and (x < ub).all() and len(set(ids.tolist())) == n
    for p in (dx, df, di):
        ctx.free(p)
    prob.close()


@pytest.mark.parametrize("mtype,handling", (("p2p", "preserve"), ("broadcast", "evict")))
def test_device_archipelago_matches_oracle_archipelago(capi, ctx, orc, mtype, handling):
    from oracle_island import OracleIsland
    from pagmo2_b200.archipelago import Archipelago, DeviceIsland
    kw = dict(topology="ring", weight=0.75, migration_type=mtype, migrant_handling=handling, seed=5)

    def dev(g):
        return DeviceIsland(0, "rastrigin", capi.algo_desc("sade", gens=2, seed=7 + g, ftol=0.0, xtol=0.0), 16, seed=100 + g, r_rate=2, s_rate=2, dim=6)

    def cpu(g):
        return OracleIsland(orc, "rastrigin", 6, 16, seed=100 + g, algo="sade", gens=2, algo_seed=7 + g, s_rate=2, r_rate=2, ftol=0.0, xtol=0.0)

    a, b = Archipelago(4, dev, **kw), Archipelago(4, cpu, **kw)
    for r in range(4):
        a.evolve(1)
        b.evolve(1)
        for ia, ib in zip(a.islands, b.islands):
            pa, pb = ia.population(), ib.population()
            assert np.array_equal(pa.ids, pb.ids), r
            assert np.allclose(pa.x, pb.x, rtol=1e-9, atol=1e-12) and np.allclose(pa.f, pb.f, rtol=1e-9)
    assert [(e.round, e.id, e.src, e.dst) for e in a.log] == [(e.round, e.id, e.src, e.dst) for e in b.log] and a.log


def test_cfg5_archipelago_runs_and_improves(capi, ctx, orc):
    """BASELINE cfg5 in small: cec2013 D=50, 8 islands x 1024, sade, ring - champions never get worse, migrants get in."""
    from pagmo2_b200.archipelago import Archipelago, DeviceIsland
    mr, os_ = orc.cec2013_tables(50)

    def dev(g):
        return DeviceIsland(0, "cec2013", capi.algo_desc("sade", gens=10, seed=11 + g), 1024, seed=200 + g, prob_id=12, dim=50, rotation=mr, shift=os_)

    a = Archipelago(8, dev, topology="ring", seed=1)
    f0 = a.champions_f()
    a.evolve(3)
    f1 = a.champions_f()
    assert (f1 <= f0).all() and (f1 < f0).any() and a.log
    p = a.islands[3].population()
    assert np.allclose(p.f[:, 0], orc.cec2013(12, p.x), rtol=1e-12)
