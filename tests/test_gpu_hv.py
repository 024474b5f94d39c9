"""GPU parity of the hypervolume path (compute and exclusive contributions, 2 and 3 objectives) against the reference's own fixtures,
golden outputs of the compiled reference (hv2d / hv3d / HyCon3D) and the restated oracle."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
REL = 1e-12


def cases(g, prefix):
    return sorted({k[:-2] for k in g.files if k.startswith(prefix) and k.endswith("_p")})


def test_reference_fixtures(ctx):
    g = np.load(GOLD / "hv_ref.npz")
    for k in cases(g, "compute_"):  # tests/hypervolume_test_data/testcases_list.txt:23-28, eps 1e-8 there
        assert abs(ctx.hv_compute(g[k + "_p"], g[k + "_r"]) - g[k + "_a"][0]) < 1e-8, k
    for k in cases(g, "exclusive_"):
        idx, want = int(g[k + "_a"][0]), g[k + "_a"][1]
        assert abs(ctx.hv_contributions(g[k + "_p"], g[k + "_r"])[idx] - want) < 1e-8, k
    for k in cases(g, "least_"):
        assert int(np.argmin(ctx.hv_contributions(g[k + "_p"], g[k + "_r"]))) == int(g[k + "_a"][0]), k


def test_golden_outputs_of_the_compiled_reference(ctx):
    g = np.load(GOLD / "hv_ref.npz")
    for k in cases(g, "ref_"):
        p, r, hv, c = g[k + "_p"], g[k + "_r"], g[k + "_hv"][0], g[k + "_c"]
        assert abs(ctx.hv_compute(p, r) - hv) <= REL * hv, k
        got = ctx.hv_contributions(p, r)
        # non-dominated fronts: HyCon3D and the device both add positive boxes -> relative agreement per point.  With dominated points
        # the reference switches to hvwfg (hv_hv3d.cpp:236-238), which subtracts hypervolumes: its own values carry ~1e-16 * HV.
        nondominated = "front" in k or len(c) <= 2
        tol = REL * np.abs(c) if nondominated else np.maximum(REL * np.abs(c), 4e-15 * hv)
        assert np.all(np.abs(got - c) <= tol), (k, np.abs(got - c).max())


@pytest.mark.parametrize("m", (2, 3))
def test_against_oracle_and_properties(ctx, orc, m):
    rng = np.random.default_rng(60 + m)
    for n, kind in ((1, "random"), (2, "random"), (65, "random"), (400, "front"), (400, "random"), (257, "ties")):
        f = rng.uniform(0, 1, (n, m))
        if kind == "front":
            f = f / np.linalg.norm(f, axis=1, keepdims=True)
        if kind == "ties":
            f = np.round(f * 8) / 8  # many equal coordinates, duplicates and dominated points
        r = np.full(m, 1.25)
        hv, c = ctx.hv_compute(f, r), ctx.hv_contributions(f, r)
        hv_o = orc.hv_compute(f, r)
        assert abs(hv - hv_o) <= REL * hv_o, (n, kind)
        assert np.abs(c - orc.hv_contributions(f, r)).max() <= 1e-14 * hv_o, (n, kind)
        assert (c >= 0).all() and c.sum() <= hv * (1 + 1e-12)
        # definition: removing point i loses exactly its exclusive contribution
        for i in rng.choice(n, size=min(n, 5), replace=False):
            if n > 1:
                assert abs((hv - ctx.hv_compute(np.delete(f, i, axis=0), r)) - c[i]) <= 1e-13 * hv
    # large front: permutation invariance (every point has its own sweep) and additivity of a far-away clone
    f = rng.uniform(0, 1, (20000, m))
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    r = np.full(m, 1.25)
    c = ctx.hv_contributions(f, r)
    perm = rng.permutation(len(f))
    assert np.allclose(ctx.hv_contributions(f[perm], r), c[perm], rtol=1e-12, atol=0)
    assert abs(ctx.hv_compute(f[perm], r) - ctx.hv_compute(f, r)) <= 1e-12 * ctx.hv_compute(f, r)


@pytest.mark.parametrize("m", (4, 5, 7))
def test_wfg_against_oracle_and_properties(ctx, orc, m):
    """4 and more objectives: the device WFG (hv.cu, replaces hvwfg hv_hvwfg.cpp:64-117) against the restated WFG, which
    tests/test_oracle.py pins to the compiled reference."""
    rng = np.random.default_rng(80 + m)
    for n, kind in ((1, "random"), (2, "random"), (3, "random"), (40, "random"), (150 if m < 7 else 60, "front"), (120 if m < 7 else 50, "ties")):
        f = rng.uniform(0, 1, (n, m))
        if kind == "front":
            f = f / np.linalg.norm(f, axis=1, keepdims=True)
        if kind == "ties":
            f = np.round(f * 4) / 4
        r = np.full(m, 1.25)
        hv, c = ctx.hv_compute(f, r), ctx.hv_contributions(f, r)
        hv_o, c_o = orc.hv_compute(f, r), orc.hv_contributions(f, r)
        assert abs(hv - hv_o) <= REL * hv_o, (n, kind)
        assert np.abs(c - c_o).max() <= 1e-13 * hv_o, (n, kind)
        assert (c >= -1e-15).all() and c.sum() <= hv * (1 + 1e-12)
        for i in rng.choice(n, size=min(n, 4), replace=False):
            if n > 1:
                assert abs((hv - ctx.hv_compute(np.delete(f, i, axis=0), r)) - c[i]) <= 1e-13 * hv
        perm = rng.permutation(n)
        assert abs(ctx.hv_compute(f[perm], r) - hv) <= 1e-13 * hv
    # a product structure with a known answer: points (a_i, b_i, 0, 0) in 4 objectives -> hv = hv2d(a, b) * r_3 * r_4
    ab = rng.uniform(0, 1, (300, 2))
    f = np.column_stack([ab, np.zeros((300, m - 2))])
    r = np.full(m, 1.5)
    want = ctx.hv_compute(ab, r[:2]) * 1.5 ** (m - 2)
    assert abs(ctx.hv_compute(f, r) - want) <= 1e-12 * want


def test_wfg_golden_outputs_of_the_compiled_reference(ctx):
    """the reference's own WFG fixtures (5 objectives x 1024 points, 7 x 64, the exclusive cases) and hvwfg outputs on seeded sets"""
    g = np.load(Path(__file__).parent / "golden" / "hv_wfg_ref.npz")
    cases = lambda prefix: sorted({k[:-2] for k in g.files if k.startswith(prefix) and k.endswith("_p")})
    for k in cases("compute_"):
        got, carried, compiled = ctx.hv_compute(g[k + "_p"], g[k + "_r"]), g[k + "_a"][0], g[k + "_hv"][0]
        assert abs(got - carried) < (1e-3 if "_d7_" in k else 1e-10) * carried, k  # the unlisted d7 file carries an approximate answer
        assert abs(got - compiled) <= 1e-13 * compiled, k
    for k in cases("exclusive_"):
        idx, want = int(g[k + "_a"][0]), g[k + "_a"][1]
        assert abs(ctx.hv_contributions(g[k + "_p"], g[k + "_r"])[idx] - want) < 1e-8, k
    for k in cases("ref_"):
        hv = g[k + "_hv"][0]
        assert abs(ctx.hv_compute(g[k + "_p"], g[k + "_r"]) - hv) <= 1e-13 * hv, k
        assert np.abs(ctx.hv_contributions(g[k + "_p"], g[k + "_r"]) - g[k + "_c"]).max() <= 1e-13 * hv, k


def test_deep_staircase_takes_the_second_pass(ctx, orc):
    """one point that dominates, in the xy-plane, a 2D front of 699 points lying below it in z: its sweep needs a staircase deeper than
    the first pass's per-point capacity (512), so it is rerun with the worst-case scratch."""
    t = np.linspace(0.05, 0.95, 699)
    f = np.vstack([np.column_stack([t, 1.0 - t, np.full_like(t, 0.1) + 1e-4 * t]), [[0.0, 0.0, 0.5]]])
    r = np.full(3, 1.25)
    hv_o = orc.hv_compute(f, r)
    assert abs(ctx.hv_compute(f, r) - hv_o) <= REL * hv_o
    assert np.abs(ctx.hv_contributions(f, r) - orc.hv_contributions(f, r)).max() <= 1e-14 * hv_o


def test_errors_and_empty(ctx, capi=None):
    from pagmo2_b200 import capi
    with pytest.raises(capi.PgcError):  # hv_algorithm.cpp:226-258
        ctx.hv_compute(np.array([[0.5, 2.0]]), [1.0, 1.0])
    with pytest.raises(capi.PgcError):
        ctx.hv_contributions(np.array([[1.0, 1.0]]), [1.0, 1.0])
    with pytest.raises(capi.PgcError):
        ctx.hv_compute(np.zeros((3, 13)), np.ones(13))  # the device WFG keeps its recursion stack for at most 12 objectives
    with pytest.raises(capi.PgcError):
        ctx.hv_compute(np.array([[0.5, 0.5, 0.5, 2.0]]), np.ones(4))
    assert ctx.hv_compute(np.zeros((0, 3)), np.ones(3)) == 0.0
    assert ctx.hv_contributions(np.array([[0.25, 0.5, 0.0]]), np.ones(3))[0] == 0.75 * 0.5 * 1.0
