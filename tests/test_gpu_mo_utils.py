"""GPU parity of the multi-objective utilities (fast_non_dominated_sorting, crowding_distance, select_best_N_mo,
sort_population_mo) through the C ABI against the oracle: ranks, dominator counts and fronts INCLUDING the order inside
every front must be bit-exact; crowding distances bit-exact on tie-free fronts; selections exact on distinct keys
(the reference leaves the order of equal keys unspecified)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EX1 = np.array([[0, 7], [1, 5], [2, 3], [4, 2], [7, 1], [10, 0], [2, 6], [4, 4], [10, 2], [6, 6], [9, 5]], dtype=float)


def same_fnds(a, b):
    return (np.array_equal(a["rank"], b["rank"]) and np.array_equal(a["dom_count"], b["dom_count"])
            and len(a["fronts"]) == len(b["fronts"]) and all(np.array_equal(x, y) for x, y in zip(a["fronts"], b["fronts"])))


def test_reference_known_answers(ctx):
    # reference tests/multi_objective.cpp:96-209
    r = ctx.fnds(EX1)
    assert [list(x) for x in r["fronts"]] == [[0, 1, 2, 3, 4, 5], [6, 7, 8], [9, 10]]
    assert list(r["dom_count"]) == [0, 0, 0, 0, 0, 0, 2, 2, 3, 5, 5] and list(r["rank"]) == [0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2]
    r = ctx.fnds(np.array([[1, 2, 3], [-2, 3, 7], [-1, -2, -3], [0, 0, 0]], dtype=float))
    assert [list(x) for x in r["fronts"]] == [[1, 2], [3], [0]] and list(r["rank"]) == [2, 0, 0, 1]
    r = ctx.fnds(np.empty((4, 0)))  # {{}, {}, {}, {}}: everything in front 0
    assert [list(x) for x in r["fronts"]] == [[0, 1, 2, 3]]
    inf = np.inf
    assert list(ctx.crowding_distance(np.array([[0, 0], [-1, 1], [2, -2]], dtype=float))) == [2, inf, inf]
    assert list(ctx.crowding_distance(np.array([[0.25, 0.25, 0.25], [-1, 1, 2], [2, -2, -2]], dtype=float))) == [3, inf, inf]
    assert list(ctx.crowding_distance(np.array([[0, 0], [1, -1], [2, -2], [4, -4]], dtype=float))) == [inf, 1.0, 1.5, inf]
    assert list(ctx.crowding_distance(np.zeros((2, 2)))) == [inf, inf]
    assert list(ctx.sort_population_mo(np.array([[0.25, 0.25], [-1, 1], [2, -2]]))) == [1, 2, 0]
    assert list(ctx.sort_population_mo(EX1)) == [0, 5, 4, 3, 1, 2, 6, 8, 7, 9, 10]
    assert list(ctx.sort_population_mo(np.arange(11, dtype=float)[:, None])) == list(range(11))
    assert list(ctx.sort_population_mo(np.array([[1.0, 5, 2, 3]]))) == [0] and len(ctx.sort_population_mo(np.empty((0, 2)))) == 0


def test_errors(ctx):
    from pagmo2_b200 import capi
    for bad in (np.zeros((1, 3)), np.empty((0, 2))):
        with pytest.raises(capi.PgcError):
            ctx.fnds(bad)
    for bad in (np.zeros((1, 2)), np.zeros((2, 1)), np.empty((2, 0))):
        with pytest.raises(capi.PgcError):
            ctx.crowding_distance(bad)


@pytest.mark.parametrize("n,m", [(2, 2), (17, 2), (255, 2), (256, 3), (257, 3), (1000, 2), (5000, 2), (4096, 3), (3000, 4), (2000, 8),
                                 (300, 1), (9000, 3), (6000, 5), (4500, 1), (4096, 2), (8192, 2)])  # the last two: sort-based count pass (n % 1024 == 0, m == 2)
def test_fnds_bit_exact_vs_oracle(ctx, orc, n, m):
    rng = np.random.default_rng(n * 10 + m)
    f = rng.uniform(0, 1, (n, m))
    assert same_fnds(ctx.fnds(f), orc.fnds(f))
    # duplicated points and shared coordinates (ties in single objectives) change nothing about exactness
    g = np.round(f * 8) / 8
    assert same_fnds(ctx.fnds(g), orc.fnds(g))
    # NaN-aware dominance (custom_comparisons.hpp:54-88)
    h = f.copy()
    h[rng.integers(0, n, max(1, n // 50)), rng.integers(0, m, max(1, n // 50))] = np.nan
    assert same_fnds(ctx.fnds(h), orc.fnds(h))


def test_fnds_converged_population_big_fronts(ctx, orc):
    """A population sitting on two fronts (the NSGA-II end game): fronts larger than the shared-memory sort."""
    rng = np.random.default_rng(77)
    n = 12000
    t = rng.uniform(0, 1, n)
    f = np.stack([t, 1 - t], axis=1)
    f[n // 2:] += 0.25
    assert same_fnds(ctx.fnds(f), orc.fnds(f))


@pytest.mark.parametrize("n,m", [(3, 2), (100, 2), (1000, 3), (5000, 2)])
def test_crowding_select_sort_vs_oracle(ctx, orc, n, m):
    rng = np.random.default_rng(n + m)
    f = rng.uniform(0, 1, (n, m))
    assert np.array_equal(ctx.crowding_distance(f), orc.crowding_distance(f))  # tie-free keys: bit exact
    ro = orc.fnds(f)
    cd = np.zeros(n)
    for fi in ro["fronts"]:
        cd[fi] = 0.0 if len(fi) == 1 else orc.crowding_distance(f[fi])
    so, sg = orc.sort_population_mo(f), ctx.sort_population_mo(f)
    assert np.array_equal(sg, so)  # both stable: identical, including the +inf ties
    for N in (0, 1, n // 3, n // 2, n - 1, n, n + 3):
        assert np.array_equal(ctx.select_best_N_mo(f, N), orc.select_best_N_mo(f, N)), N


def test_nsga2_scale_population(ctx, orc):
    """BASELINE cfg3 scale (2N = 131072 points would take the O(N^2) CPU oracle minutes): check 65536 points through
    size-independent properties and a 8192-point prefix against the oracle."""
    rng = np.random.default_rng(31)
    n, m = 65536, 2
    f = rng.uniform(0, 1, (n, m))
    r = ctx.fnds(f)
    rank = r["rank"]
    assert sorted(np.concatenate(r["fronts"]).tolist()) == list(range(n))            # fronts partition the population
    assert all((rank[fr] == k).all() for k, fr in enumerate(r["fronts"]))             # ranks agree with fronts
    assert np.array_equal(r["fronts"][0], np.flatnonzero(r["dom_count"] == 0))        # front 0 ascending, no dominators
    # nobody inside a front dominates anybody else there (sampled), and every point of front k>0 has a dominator in k-1
    for k in (0, 1, len(r["fronts"]) // 2):
        fr = r["fronts"][k][:400]
        a = f[fr]
        dom = (a[:, None, :] <= a[None, :, :]).all(-1) & (a[:, None, :] < a[None, :, :]).any(-1)
        assert not dom.any()
        if k:
            prev = f[r["fronts"][k - 1]]
            for q in fr[:50]:
                assert ((prev <= f[q]).all(-1) & (prev < f[q]).any(-1)).any()
    sub = f[:8192]
    assert same_fnds(ctx.fnds(sub), orc.fnds(sub))
    best = ctx.select_best_N_mo(f, n // 2)
    assert len(best) == n // 2 and len(set(best.tolist())) == n // 2
    assert rank[best].max() <= rank[np.setdiff1d(np.arange(n), best)].min()           # no worse rank kept over a better one


@pytest.mark.parametrize("m", (2, 3))
def test_duplicate_points_follow_the_stable_restatement(ctx, orc, m):
    """clones of existing individuals (an offspring that escaped crossover and mutation) tie in every comparison: fronts, crowding
    distances and the truncated selection must still agree with the restated (stable) reference semantics."""
    rng = np.random.default_rng(70 + m)
    for n in (12, 64, 512):
        f = rng.uniform(0, 1, (n, m))
        f[n // 2] = f[1]          # a clone inside the same front
        f[n - 1] = f[n // 3]      # another one
        f[3] = f[2] + 0.0         # and a third
        f[3, 0] = f[2, 0]
        got, want = ctx.fnds(f), orc.fnds(f)
        assert np.array_equal(got["rank"], want["rank"]) and len(got["fronts"]) == len(want["fronts"])
        for a, b in zip(got["fronts"], want["fronts"]):
            assert np.array_equal(a, b)
        assert np.array_equal(ctx.crowding_distance(f[got["fronts"][0]]), orc.crowding_distance(f[got["fronts"][0]]))
        for N in (1, n // 4, n // 2, n - 1):
            assert np.array_equal(ctx.select_best_N_mo(f, N), orc.select_best_N_mo(f, N)), (n, N)
        assert np.array_equal(ctx.sort_population_mo(f), orc.sort_population_mo(f))


def test_fnds_sorted_space_degenerate_inputs(ctx, orc):
    """inputs that stress the large-input (sorted-space) level loop: every point identical (one front, nothing left to peel), two
    distinct points repeated (two fronts larger than the shared-memory sort, runs of equal first ranks spanning many blocks), a
    total order (one point per front: thousands of levels, the compacted list shrinking by one per level), and select_best_N_mo
    stopping inside such a sort."""
    n = 4608
    same = np.full((n, 2), 0.25)
    assert same_fnds(ctx.fnds(same), orc.fnds(same))
    two = np.where((np.arange(n) % 2 == 0)[:, None], np.array([[0.1, 0.2]]), np.array([[0.3, 0.4]]))
    assert same_fnds(ctx.fnds(two), orc.fnds(two))
    chain = np.stack([np.arange(n, dtype=float), np.arange(n, dtype=float)[::-1].copy()[::-1]], axis=1)  # point i dominates i + 1
    rng = np.random.default_rng(5)
    chain = chain[rng.permutation(n)]
    got = ctx.fnds(chain)
    assert same_fnds(got, orc.fnds(chain)) and len(got["fronts"]) == n
    for N in (1, 100, n // 2):
        assert np.array_equal(ctx.select_best_N_mo(chain, N), orc.select_best_N_mo(chain, N))
    # three objectives, half the points cloned
    f = rng.uniform(0, 1, (n, 3))
    f[n // 2:] = f[:n // 2]
    assert same_fnds(ctx.fnds(f), orc.fnds(f))


# ---------------------------------------------------------------- BASELINE cfg3 at its own size, in full
def _cfg3_points(orc, kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "zdt1":      # ZDT1 objectives of a random population, nx = 30: ~300-400 small fronts
        return orc.zdt(1, rng.uniform(0, 1, (n, 30)))
    if kind == "dtlz2":     # DTLZ2, three objectives, nx = 12: ~80 big fronts
        return orc.dtlz(2, rng.uniform(0, 1, (n, 12)), 3, 100)
    raise ValueError(kind)


@pytest.mark.parametrize("kind,n", [("zdt1", 65536), ("dtlz2", 65536), ("zdt1", 131072), ("dtlz2", 131072)])
def test_fnds_full_size_bit_exact_vs_oracle(ctx, orc, kind, n):
    """The sort NSGA-II runs every generation at pop 65 536 (N for the tournament ranking, 2N = 131 072 for select_best_N_mo):
    ranks, dominator counts, and every front in the reference's order, compared IN FULL with the restated reference
    (oracle_fnds_nolist: the same append order without the O(N^2) list, ~10-50 s on 8 host cores)."""
    f = _cfg3_points(orc, kind, n, n % 1000 + len(kind))
    assert same_fnds(ctx.fnds(f), orc.fnds(f))


@pytest.mark.parametrize("switch", ["PGC_FNDS_PERSIST", "PGC_FNDS_COUNT2", "PGC_FNDS_BIG_INKERNEL", "PGC_FNDS_FUSE"])
@pytest.mark.parametrize("kind", ["zdt1", "dtlz2"])
def test_fnds_fallback_paths_equal_default_path(ctx, orc, monkeypatch, switch, kind):
    """Every A/B switch of the large-input FNDS is product code: launch-per-level loop instead of the resident kernel
    (PERSIST=0, alone and with FUSE=0), all-pairs count pass (COUNT2=0), big levels through the host / CUB path
    (BIG_INKERNEL=0).  Each must give the default path's result bit for bit at the cfg3 size, and the oracle's at 12 288 points."""
    n = 65536
    f = _cfg3_points(orc, kind, n, 7)
    small = f[:12288]
    want_small = orc.fnds(small)
    default = ctx.fnds(f)
    monkeypatch.setenv(switch, "0")
    if switch == "PGC_FNDS_FUSE":
        monkeypatch.setenv("PGC_FNDS_PERSIST", "0")   # FUSE only exists in the launch-per-level loop
    assert same_fnds(ctx.fnds(f), default)
    assert same_fnds(ctx.fnds(small), want_small)
    sel = ctx.select_best_N_mo(f, n // 2)
    monkeypatch.delenv(switch)
    monkeypatch.delenv("PGC_FNDS_PERSIST", raising=False)
    assert np.array_equal(sel, ctx.select_best_N_mo(f, n // 2))


def test_fnds_without_cooperative_launch(ctx, orc, monkeypatch):
    """A device that refuses the cooperative launch (partitioned / MIG): the driver falls back to the launch-per-level loop."""
    f = _cfg3_points(orc, "zdt1", 65536, 9)
    default = ctx.fnds(f)
    monkeypatch.setenv("PGC_FNDS_FORCE_NOCOOP", "1")
    assert same_fnds(ctx.fnds(f), default)
    g = f[:9000]
    assert same_fnds(ctx.fnds(g), orc.fnds(g))


def test_fnds_above_resident_capacity(ctx, orc):
    """More points than one thread per position of the resident grid (148 x 1024 = 151 552): the launch-per-level loop is the
    only path; full compare with the oracle."""
    n = 155648
    f = _cfg3_points(orc, "zdt1", n, 3)
    assert same_fnds(ctx.fnds(f), orc.fnds(f))


def test_fnds_watchdog_surfaces_as_error(ctx, orc, monkeypatch):
    """A lost wake-up in the resident level loop must end as PGC_ERR_CUDA with a message, not as a hung device: block 0 is made to
    drop out (PGC_FNDS_INJECT_STUCK=1), the other blocks' spin watchdogs fire, the host reports it, and the context stays usable."""
    from pagmo2_b200 import capi
    f = _cfg3_points(orc, "zdt1", 65536, 4)
    monkeypatch.setenv("PGC_FNDS_INJECT_STUCK", "1")
    with pytest.raises(capi.PgcError, match="lost a wake-up"):
        ctx.fnds(f)
    monkeypatch.delenv("PGC_FNDS_INJECT_STUCK")
    g = f[:20000]
    assert same_fnds(ctx.fnds(g), orc.fnds(g))


@pytest.mark.parametrize("m", (2, 3, 4))
def test_crowding_ties_follow_the_carried_index_vector(ctx, orc, m):
    """crowding_distance sorts ONE index vector objective after objective (multi_objective.cpp:296-313): among equal values of an
    objective the order left by the previous objective decides who is the boundary point.  Quantised objectives (many exact ties,
    DTLZ-like zeros) through crowding_distance, sort_population_mo and select_best_N_mo, device == oracle exactly."""
    rng = np.random.default_rng(100 + m)
    for n in (5, 64, 257, 1500):
        f = np.round(rng.uniform(0, 1, (n, m)) * 6) / 6      # ~7 distinct values per objective
        f[rng.integers(0, n, n // 3), m - 1] = 0.0
        front = f[orc.fnds(f)["fronts"][0]]
        if len(front) >= 2:
            assert np.array_equal(ctx.crowding_distance(front), orc.crowding_distance(front), equal_nan=True), (m, n)
        assert np.array_equal(ctx.sort_population_mo(f), orc.sort_population_mo(f)), (m, n)
        for N in (1, n // 3, n // 2, n - 1):
            if N >= 1:
                assert np.array_equal(ctx.select_best_N_mo(f, N), orc.select_best_N_mo(f, N)), (m, n, N)
