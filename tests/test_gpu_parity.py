"""GPU parity tests: the sm_100a path, called through the C ABI, against the CPU
oracle on the same seeded inputs and the same permutation matrix, and against
the reference's golden vectors.  Bars (BASELINE.json north_star): gene scores
and ranks bit-exact, leading-edge indices exact, ES/NES within 1e-10 absolute,
p-value / FDR counts identical."""
import json
import os

import numpy as np
import pytest

import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "weighted_ks.json")))
ES_TOL = 1e-10  # absolute, north_star
NA = np.iinfo(np.int32).min


@pytest.fixture(scope="module")
def ctx():
    c = fg.Context(0)
    # the tests below are about the production kernels: keep small problems off the
    # all-exact path they would otherwise take (test_random_small_pipelines turns it on)
    c.set_option("es_exact_small", 0)
    yield c
    c.close()


def _case(G=1000, n=20, S=50, P=96, sizes=(10, 50), dist="uniform", seed=0):
    off, idx = synth.gene_sets_csr(G, S, sizes, dist, seed=seed + 1)
    x, y = synth.two_class(G, n, off, idx, seed=seed + 2)
    perm = synth.perm_matrix(n, P, seed=seed + 4)
    return x, y, perm, off, idx


def _ranks(oracle, scores):  # scores [G, C]
    return np.stack([oracle.rank_desc(scores[:, c]) for c in range(scores.shape[1])], axis=1)


def _assert_lm_null(oracle, grk, flat, null, ref, G, off):
    """lm parity is tolerance-based (a GEMM here, a Householder restatement in the oracle, LINPACK
    in R -- lm scores are not bit-pinned anywhere): ranks may only differ between genes whose
    oracle scores lie within 1e-12, and the null ES is asserted for EVERY column: within 1e-10
    where the ranks agree, and within 1e-10 + (rank differences of the column) x (1 / (G - m) +
    the weight step of one position) where they do not (a flipped pair moves one hit by one rank)."""
    S, R, P = ref.shape
    rr = _ranks(oracle, flat)
    bad = np.argwhere(grk != rr)
    for g, c in bad:
        other = np.where(rr[:, c] == grk[g, c])[0][0]
        assert abs(flat[g, c] - flat[other, c]) <= 1e-12
    ndiff = (grk != rr).sum(axis=0)                      # per score column c = p * R + r
    assert (ndiff > 0).mean() <= 0.05, "rank differences in %d of %d columns" % ((ndiff > 0).sum(), ndiff.size)
    m_max = int(np.diff(off).max())
    err = np.abs(null - ref).reshape(S, R * P, order="F").max(axis=0)
    bound = ES_TOL + ndiff * (2.0 / (G - m_max))
    assert (err <= bound).all(), (err.max(), int(np.argmax(err - bound)))
    assert err[ndiff == 0].max() <= ES_TOL


# ---------------------------------------------------------------- a4: s2n_C
def test_s2n_bit_exact(ctx, oracle):
    from test_oracle_golden import X_TOY
    y = np.array([[1, 0, 0, 1], [0, 0, 1, 1]], dtype=np.int32).T
    assert np.array_equal(ctx.s2n(X_TOY, y), oracle.s2n_C(X_TOY, y))
    rng = np.random.default_rng(5)
    for n, G, R in ((37, 211, 3), (200, 1000, 1), (1, 5, 1), (33, 64, 2), (901, 70, 1)):
        x = rng.normal(6, 2, size=(n, G))
        yb = (rng.random((n, R)) < 0.4).astype(np.int32)
        if n > 3:
            yb[3, R - 1] = NA
        got = ctx.s2n(x, yb)
        want = oracle.s2n_C(x, yb)
        assert np.array_equal(got, want, equal_nan=True), (n, G, R)
        if oracle.ref_lib() is not None:
            assert np.array_equal(got, oracle.s2n_C(x, yb, use_ref=True), equal_nan=True)
    # row mismatch -> zeros (src/ggsea.cpp:15-17)
    assert not ctx.s2n(rng.normal(size=(10, 4)), np.ones((9, 1), dtype=np.int32)).any()


# ------------------------------------------------- a1/a2/a3/a6/a7: null loop
@pytest.mark.parametrize("s2n_mode", [0, 1])
@pytest.mark.parametrize("R", [1, 2])
@pytest.mark.parametrize("abs_flag", [False, True])
def test_perm_null_s2n_wks(ctx, oracle, R, abs_flag, s2n_mode):
    """s2n_mode 0: bit-faithful score kernel (scores bit-exact); s2n_mode 1
    (default): FP64 tensor-core contraction + certification (scores to 1e-11
    relative, ranks still bit-exact)."""
    ctx.set_option("s2n_mode", s2n_mode)
    x, y, perm, off, idx = _case()
    if R == 2:
        rng = np.random.default_rng(9)
        y = np.column_stack([y[:, 0], rng.permutation(y[:, 0])])
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    P = perm.shape[1]
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm, abs_flag=abs_flag)
    rc, want, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS,
                                    abs_flag=abs_flag, return_scores=True)
    assert rc == 0
    gsc, grk = ctx.debug_last(P * R)
    ctx.set_option("s2n_mode", 1)
    want_sc = sc.reshape(sc.shape[0], -1, order="F")
    if abs_flag:  # the device keeps signed scores and applies abs when ranking
        gsc = np.abs(gsc)
    if s2n_mode == 0:
        assert np.array_equal(gsc, want_sc)                     # bit-exact scores
    else:
        assert np.allclose(gsc, want_sc, rtol=1e-11, atol=1e-13)
    assert np.array_equal(grk, _ranks(oracle, want_sc))         # bit-exact ranks
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= ES_TOL


@pytest.mark.parametrize("p_exp", [0.0, 0.5, 2.0])
def test_weighted_ks_weight_exponent(ctx, oracle, p_exp):
    """w = |s|^p (R/flexgsea.R:819) for p != 1: p = 0 is the classic unweighted KS
    (0^0 = 1 as in R), fractional and > 1 exponents go through pow()."""
    x, y, perm, off, idx = _case(G=600, S=40, P=32, sizes=(5, 90))
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 32, perm=perm, p=p_exp)
    rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS, p=p_exp)
    assert rc == 0 and np.abs(got - want).max() <= ES_TOL


@pytest.mark.parametrize("G", [11, 21, 64])
def test_tiny_universe_structural_ties(ctx, oracle, G):
    """few genes and sets of 1-4 members: es_pos = -es_neg exactly for many units
    (a single gene at the median rank of an odd universe, symmetric pairs), with
    p = 1 and p = 0, signed and abs scores.  The sign must be the reference's."""
    rng = np.random.default_rng(G)
    n, P = 12, 120
    sets = [rng.choice(G, m, replace=False) + 1 for m in (1, 1, 1, 2, 2, 3, 4, 1, 2, 3, G - 1, G // 2)]
    off, idx = oracle.csr(sets)
    x = np.asfortranarray(rng.normal(size=(n, G)))
    y = np.zeros((n, 1), dtype=np.int32); y[: n // 2] = 1
    perm = synth.perm_matrix(n, P, seed=G + 1)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    for p_exp in (1.0, 0.0):
        for abs_flag in (False, True):
            got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm, p=p_exp, abs_flag=abs_flag)
            rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS,
                                        p=p_exp, abs_flag=abs_flag)
            assert rc == 0
            assert np.array_equal(np.isnan(got), np.isnan(want))
            ok = ~np.isnan(want)
            assert np.abs(got[ok] - want[ok]).max() <= ES_TOL, (p_exp, abs_flag)


def test_random_small_shapes(ctx):
    """2 200 random small problems (profiles/stress_parity.py: 7-3000 genes, 4-40
    samples, sets with duplicates, tied / rounded data, weight exponents 0 / 1 / 2,
    abs, mean / maxmean / weighted KS, both score modes) against the oracle, NaN
    pattern included.  This sweep found, and now pins: exact-zero scores that the
    tensor path made 1e-17 (weight 0 -> NaN ES in the reference), +-Inf scores
    re-certified after the 1.1 * max patch, G == m reached through duplicated
    members, |s|^2 as s * s."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "stress_parity", os.path.join(os.path.dirname(__file__), "..", "profiles", "stress_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(2200, 7, ctx) == 0


def test_random_small_pipelines(ctx):
    """1 500 random small flexgsea() pipelines (profiles/stress_pipeline.py; weighted KS,
    maxmean and mean; s2n and user-made score columns with ties): observed
    scores bit-exact, observed ES / max_es_at / le_prop / leading edge / running_es_at
    exact, and the p-value COUNTS identical to the oracle pipeline even with 4-8
    samples and a dozen genes, where the null takes a handful of values and equality
    with the observed ES is decided by the reference's rounding (units within rounding
    of the observed ES are redone in the reference's arithmetic); pooled-FDR counts
    identical from 150 genes and 6 samples up."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "stress_pipeline", os.path.join(os.path.dirname(__file__), "..", "profiles", "stress_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(1500, 5, ctx) == 0
    ctx.set_option("es_exact_small", 1)      # the default: small problems entirely in the reference's arithmetic
    try:
        assert mod.run(1500, 5, ctx, fdr_strict_always=True) == 0
    finally:
        ctx.set_option("es_exact_small", 0)


def test_tensor_scores_certified_ranks(ctx, oracle):
    """near ties under the tensor-core score path: duplicated genes (exact ties in
    the reference), a gene and its 10x copy (1 ulp apart in the reference,
    tests/testthat/test_s2n.R:44-48), and a gene whose classes are far apart
    relative to its variance (cancellation in E[x^2] - mean^2).  Ranks must equal
    the reference's; also when the certification lists overflow (fallback)."""
    G, n, P = 800, 24, 40
    x, y = synth.two_class(G, n, seed=11)
    x = x.copy()
    x[:, 100] = x[:, 7]; x[:, 101] = x[:, 7]; x[:, 300] = x[:, 299]          # duplicates
    x[:, 400] = 10.0 * x[:, 8]; x[:, 401] = 1e3 * x[:, 9]                     # scaled copies
    x[:, 500] = np.where(y[:, 0] == 1, 1e4, -1e4) + 1e-3 * x[:, 10]           # cancellation
    off, idx = synth.gene_sets_csr(G, 20, (10, 60), "uniform", seed=5)
    perm = synth.perm_matrix(n, P, seed=2)
    perm[:, 0] = np.arange(1, n + 1)                                          # identity: the far-apart labelling
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    rc, want, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS,
                                    return_scores=True)
    assert rc == 0
    want_rk = _ranks(oracle, sc[:, 0, :])
    for cap in (4 << 20, 16):                                                 # 16: forces the overflow fallback
        ctx.set_option("cert_cap", cap)
        got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
        gsc, grk = ctx.debug_last(P)
        assert np.array_equal(grk, want_rk), cap
        assert np.abs(got - want).max() <= ES_TOL
    ctx.set_option("cert_cap", 4 << 20)


def test_perm_null_blocks_and_generic_kernel(ctx, oracle):
    """permutation blocking must not change results; the order-free kernel
    (used for duplicate / oversized sets) agrees with the bitmap kernel."""
    x, y, perm, off, idx = _case(P=70)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    a = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 70, perm=perm)
    ctx.set_option("perm_block", 16)
    b = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 70, perm=perm)
    ctx.set_option("perm_block", 0)
    assert np.array_equal(a, b)
    ctx.set_option("force_generic_es", 1)
    g = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 70, perm=perm)
    ctx.set_option("force_generic_es", 0)
    assert np.abs(a - g).max() <= 1e-13


def test_duplicates_large_and_edge_sets(ctx, oracle):
    """sets with duplicated members (R keeps them, m counts them), a set above
    the bitmap capacity, a single-gene set and the all-genes set (ES = NaN)."""
    G, n, P = 700, 16, 24
    rng = np.random.default_rng(3)
    sets = [rng.choice(G, 30, replace=False) + 1,
            np.array([5, 9, 5, 77, 9, 200]),              # duplicates
            rng.choice(G, 600, replace=False) + 1,        # > 512 members
            np.array([42]),
            np.arange(1, G + 1),                          # G - m == 0 -> NaN
            rng.choice(G, 512, replace=False) + 1]        # exactly the capacity
    off, idx = oracle.csr(sets)
    x, y = synth.two_class(G, n, seed=8)
    perm = synth.perm_matrix(n, P, seed=1)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
    rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS)
    assert np.isnan(want[4]).all() and np.isnan(got[4]).all()
    keep = [0, 1, 2, 3, 5]
    assert np.abs(got[keep] - want[keep]).max() <= ES_TOL


def test_wide_sets_and_class_boundaries(ctx, oracle):
    """every size class of the in-register weighted-KS kernel at its boundaries
    (8 lanes x {4..32} ranks, 16 x 32, a whole warp x {32, 64}), a set above
    ES_MAX_SET = 2048 (order-free kernel), quads that mix classes, S % 4 != 0."""
    G, n, P = 5000, 24, 20
    rng = np.random.default_rng(11)
    sizes = [2049, 2048, 1500, 1025, 1024, 700, 513, 512, 300, 257, 256, 255, 129, 128, 127, 65, 64, 63,
             33, 32, 31, 17, 16, 9, 8, 7, 5, 4, 3, 2, 1, 40, 90, 200, 400, 1000, 12]
    assert len(sizes) % 4 != 0
    sets = [rng.choice(G, m, replace=False) + 1 for m in sizes]
    off, idx = oracle.csr(sets)
    x, y = synth.two_class(G, n, seed=21)
    perm = synth.perm_matrix(n, P, seed=4)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
    rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS)
    assert rc == 0 and np.abs(got - want).max() <= ES_TOL
    ctx.set_option("force_generic_es", 1)
    try:
        gen = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
    finally:
        ctx.set_option("force_generic_es", 0)
    assert np.abs(got - gen).max() <= ES_TOL


@pytest.mark.parametrize("abs_flag", [False, True])
@pytest.mark.parametrize("es_kind", ["maxmean", "mean"])
def test_perm_null_mean_maxmean(ctx, oracle, es_kind, abs_flag):
    """two responses, sets of 1 .. 700 members (more than one 16-byte member block
    per lane), an empty-ish tail (S % 4 != 0), signed and abs scores"""
    G, n, P = 1500, 20, 40
    rng = np.random.default_rng(17)
    sizes = [1, 2, 7, 8, 9, 63, 64, 65, 700, 300, 33, 5, 120, 16, 17]
    sets = [rng.choice(G, m, replace=False) + 1 for m in sizes] + [np.array([4, 4, 9, 4])]  # duplicates count
    off, idx = oracle.csr(sets)
    x, y = synth.two_class(G, n, seed=12)
    y = np.column_stack([y[:, 0], rng.permutation(y[:, 0])])
    perm = synth.perm_matrix(n, P, seed=13)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    k = nat.ES_MAXMEAN if es_kind == "maxmean" else nat.ES_MEAN
    ok = oracle.ES_MAXMEAN if es_kind == "maxmean" else oracle.ES_MEAN
    got = ctx.perm_null(nat.SCORE_S2N, k, P, perm=perm, abs_flag=abs_flag)
    rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, ok, abs_flag=abs_flag)
    assert rc == 0 and got.shape == want.shape and np.abs(got - want).max() <= ES_TOL
    assert np.array_equal(np.sign(got), np.sign(want))


def test_philox_permutations_match_definition(ctx, oracle):
    """device-drawn permutations: same Philox4x32-10 Fisher-Yates as the oracle's
    definition, keyed by the GLOBAL permutation id (shard independent)."""
    x, y, _, off, idx = _case(P=1)
    n = x.shape[0]
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    seed, P = 1234567, 48
    a = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, seed=seed, perm_id0=0)
    perm = np.stack([oracle.philox_perm(seed, p, n) for p in range(P)], axis=1)
    b = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
    assert np.array_equal(a, b)
    lo = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 20, seed=seed, perm_id0=0)
    hi = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P - 20, seed=seed, perm_id0=20)
    assert np.array_equal(np.concatenate([lo, hi], axis=2), a)


def test_nonfinite_scores_raise_like_reference(ctx):
    """a gene with zero variance in both classes and equal means gives 0/0:
    the reference stops with "Gene scoring function returned NA or Inf."
    (R/flexgsea.R:251-253, :400-402)."""
    x, y, perm, off, idx = _case(P=8)
    x = x.copy(); x[:, 7] = 3.0
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    with pytest.raises(fg.NonFiniteScoreError):
        ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 8, perm=perm)
    with pytest.raises(fg.NonFiniteScoreError):
        ctx.score_observed(nat.SCORE_S2N)


def test_inf_patch(ctx, oracle):
    """zero variance in both classes but different means -> +-Inf, patched to
    1.1 * max/min finite over the permutation's score matrix (:715-716)."""
    G, n = 300, 12
    x, y = synth.two_class(G, n, seed=5)
    x = x.copy()
    x[:, 3] = np.where(y[:, 0] == 1, 2.0, 1.0)   # +Inf
    x[:, 11] = np.where(y[:, 0] == 1, 1.0, 2.0)  # -Inf
    off, idx = synth.gene_sets_csr(G, 10, (10, 30), "uniform", seed=2)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    got = ctx.score_observed(nat.SCORE_S2N)
    want = oracle.s2n_post(oracle.s2n_C(x, y))
    assert np.isfinite(want).all() and np.array_equal(got, want)
    perm = np.arange(1, n + 1, dtype=np.int32).reshape(-1, 1)
    null = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 1, perm=perm)
    rc, ref = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS)
    assert rc == 0 and np.abs(null - ref).max() <= ES_TOL


# ------------------------------------------------ a6: ranking (ties, -0.0)
def test_rank_ties_and_user_scores(ctx, oracle):
    rng = np.random.default_rng(2)
    G, R, P = 517, 2, 9
    sc = rng.normal(size=(G, R, P))
    sc[rng.integers(0, G, 60), 0, :] = 0.25        # exact ties
    sc[10, :, :] = 0.0; sc[20, :, :] = -0.0        # -0.0 ties with 0.0
    sc[:, 1, 3] = np.round(sc[:, 1, 3], 1)         # heavy ties
    off, idx = synth.gene_sets_csr(G, 20, (5, 80), "uniform", seed=6)
    ctx.set_gene_sets(off, idx)
    got = ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
    _, rk = ctx.debug_last(R * P, G=G)
    flat = sc.reshape(G, -1, order="F")
    assert np.array_equal(rk, _ranks(oracle, flat))
    rc, want = oracle.es_from_scores(oracle.ES_WKS, sc, off, idx)
    assert rc == 0 and np.abs(got - want).max() <= ES_TOL
    for kind, ok in ((nat.ES_MEAN, oracle.ES_MEAN), (nat.ES_MAXMEAN, oracle.ES_MAXMEAN)):
        g2 = ctx.es_from_scores(kind, sc)
        _, w2 = oracle.es_from_scores(ok, sc, off, idx)
        assert np.abs(g2 - w2).max() <= ES_TOL
    bad = sc.copy(); bad[3, 0, 0] = np.inf
    with pytest.raises(fg.NonFiniteScoreError):
        ctx.es_from_scores(nat.ES_WEIGHTED_KS, bad)


def test_sort_paths_agree(ctx, oracle):
    """the on-chip 32-bit-key sort (+ low-half fix-up, + 64-bit redo of columns
    with long runs) and the 64-bit global sort give R's ranks on adversarial
    inputs: values that share their top 32 key bits, exact duplicates, zeros."""
    rng = np.random.default_rng(12)
    G = 3000
    cols = [1.0 + rng.permutation(G) * 2.0 ** -44,                    # one giant run, all distinct
            np.repeat(rng.normal(size=G // 4), 4)[:G],                # exact ties in groups of 4
            np.where(rng.random(G) < 0.5, 0.0, -0.0),                 # +-0 only
            rng.normal(size=G),
            np.concatenate([rng.normal(size=G - 40), 2.5 + np.arange(40) * 2.0 ** -50])]  # run of 40 < 48
    sc = np.stack(cols, axis=1)[:, None, :]                           # [G, 1, 5]
    off, idx = synth.gene_sets_csr(G, 8, (10, 40), "uniform", seed=3)
    ctx.set_gene_sets(off, idx)
    want = _ranks(oracle, sc[:, 0, :])
    for force64 in (0, 1):
        ctx.set_option("force_sort64", force64)
        ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
        _, rk = ctx.debug_last(5, G=G)
        assert np.array_equal(rk, want), force64
    ctx.set_option("force_sort64", 0)


# --------------------------------------- a8: observed extras + the goldens
def test_weighted_ks_golden_on_device(ctx, oracle):
    """tests/testthat/test_weighted_ks.R through the device path."""
    rng = fg.RRNG(GOLD["seed"], "Rounding")
    scores = rng.rnorm(3000).reshape(3, 1000).T
    gs = rng.sample_int(1000, 20)
    ctx.set_gene_sets(np.array([0, 20], dtype=np.int32), gs)
    o = ctx.observed_es(nat.ES_WEIGHTED_KS, scores)
    assert np.abs(o["es"][0] - np.array(GOLD["es"]["values"])).max() <= 1e-12
    assert list(o["max_es_at"][0].astype(int)) == GOLD["max_es_at"]["values"]
    assert list(o["le_prop"][0]) == GOLD["le_prop"]["values"]
    null = ctx.es_from_scores(nat.ES_WEIGHTED_KS, scores[:, :, None])
    assert np.abs(null[0, :, 0] - np.array(GOLD["es"]["values"])).max() <= 1e-12
    res = fg.flexgsea_weighted_ks.run(scores[:, :, None], gs, None, return_stats=("le_prop", "max_es_at"))
    assert np.abs(res["es"][:, 0] - np.array(GOLD["es"]["values"])).max() <= 1e-12
    assert list(res["max_es_at"][:, 0].astype(int)) == GOLD["max_es_at"]["values"]


def test_observed_extras_exact(ctx, oracle):
    x, y, _, off, idx = _case(S=40)
    dup = np.array([5, 9, 5, 77, 9, 200], dtype=np.int32)
    off = np.concatenate([off, [off[-1] + dup.size]]).astype(np.int32)
    idx = np.concatenate([idx, dup]).astype(np.int32)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    sc = ctx.score_observed(nat.SCORE_S2N)
    o = ctx.observed_es(nat.ES_WEIGHTED_KS, sc)
    rank = oracle.rank_desc(sc[:, 0])
    for s in range(off.size - 1):
        w = oracle.wks_column(sc[:, 0], rank, idx[off[s]:off[s + 1]], extras=True)
        sl = slice(off[s], off[s + 1])
        assert abs(o["es"][s, 0] - w["es"]) <= 1e-15
        assert o["max_es_at"][s, 0] == w["max_es_at"] and o["le_prop"][s, 0] == w["le_prop"]
        assert np.array_equal(o["running_es_at"][sl, 0], w["running_es_at"])
        assert np.abs(o["running_es_pos"][sl, 0] - w["running_es_pos"]).max() <= 1e-15
        assert np.abs(o["running_es_neg"][sl, 0] - w["running_es_neg"]).max() <= 1e-15
        le = o["running_es_at"][off[s] + o["le_first"][s, 0]: off[s] + o["le_first"][s, 0] + o["le_count"][s, 0], 0]
        assert np.array_equal(le, w["leading_edge"])           # leading-edge indices exact


# ------------------------------------------------------- a5: flexgsea_lm
def test_lm_scores_and_null(ctx, oracle):
    G, n, P = 600, 40, 30
    off, idx = synth.gene_sets_csr(G, 25, (10, 60), "uniform", seed=3)
    x, _ = synth.two_class(G, n, seed=4)
    design = synth.continuous_response(n, k=2, seed=5)
    perm = synth.perm_matrix(n, P, seed=6)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_lm(design, True)
    got = ctx.score_observed(nat.SCORE_LM)
    rc, want = oracle.lm_scores(x, design, True)
    assert rc == 0 and got.shape == want.shape == (G, 3)
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    null = ctx.perm_null(nat.SCORE_LM, nat.ES_WEIGHTED_KS, P, perm=perm)
    rc, ref, sc = oracle.perm_null(x, oracle.SCORE_LM, design, True, perm, off, idx, oracle.ES_WKS,
                                   return_scores=True)
    assert rc == 0
    gsc, grk = ctx.debug_last(P * 3)
    flat = sc.reshape(G, -1, order="F")
    assert np.abs(gsc - flat).max() <= 1e-12
    _assert_lm_null(oracle, grk, flat, null, ref, G, off)
    with pytest.raises(fg.NonFiniteScoreError):
        ctx.set_response_lm(np.column_stack([design, design[:, 1]]), True)


# ------------------------------------------------- a10-a12: significance
@pytest.mark.parametrize("abs_flag", [False, True])
@pytest.mark.parametrize("simple", [False, True])
def test_calc_sig_counts_identical(ctx, oracle, simple, abs_flag):
    rng = np.random.default_rng(21)
    S, P = 300, 700
    es = rng.normal(size=S); en = rng.normal(size=(S, P))
    if abs_flag:
        es, en = np.abs(es), np.abs(en)
    es[4] = en[4].max() + 1; es[7] = 0.0; en[7, :5] = 0.0
    en[9, :] = np.abs(en[9, :]); es[9] = -0.3          # no negative null: NaN mean fix-up (:615-616)
    ctx.set_null(en)
    got = ctx.calc_sig(es, 0, not simple, not simple, abs_flag)
    want = oracle.calc_sig(es, en, not simple, not simple, abs_flag)
    assert np.array_equal(got["p_counts"], want["p_counts"])
    assert np.array_equal(got["p"], want["p"])
    assert np.array_equal(got["fwer"], want["fwer"])
    if simple:
        assert np.array_equal(got["p_high"], want["p_high"])
        if not abs_flag:
            assert np.array_equal(got["p_low"], want["p_low"])
        assert np.allclose(got["fdr"], want["fdr"], rtol=1e-14, atol=0)
    else:
        assert np.abs(got["nes"] - want["nes"]).max() <= ES_TOL
        assert np.array_equal(got["fdr_counts"], want["fdr_counts"])
        assert np.array_equal(got["glob_counts"], want["glob_counts"])
        assert np.allclose(got["fdr"], want["fdr"], rtol=1e-12, atol=0)


def test_small_n_label_preserving_permutations(ctx, oracle):
    """Six samples, 3 + 3: one permutation in twenty reproduces the observed
    labelling, and the reference then gets a null ES bit-identical to the observed
    one (same code, same numbers), which its `<=` / `>=` counts see as equal.  Here
    the observed ES comes from the exact kernel and the null from the tensor-core /
    in-register path, so the significance kernels take null values within rounding
    of the observed ES as equal to it: p and FDR counts must match the reference."""
    G, n, S, P = 300, 6, 25, 400
    off, idx = synth.gene_sets_csr(G, S, (8, 40), "uniform", seed=3)
    x, y = synth.two_class(G, n, off, idx, seed=4)
    perm = synth.perm_matrix(n, P, seed=6)
    yb = y.astype(np.int32)
    same = sum(np.array_equal(yb[perm[:, p] - 1, 0], yb[:, 0]) for p in range(P))
    assert same >= 5                                    # the case is really exercised
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(yb)
    sc = ctx.score_observed(nat.SCORE_S2N)
    obs = ctx.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=False)["es"][:, 0]
    ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm, want_host=False)
    got = ctx.calc_sig(obs, 0, True, True, False)
    o_sc = oracle.s2n_post(oracle.s2n_C(x, yb))
    rc, o_obs = oracle.es_from_scores(oracle.ES_WKS, o_sc.reshape(G, 1, 1), off, idx)
    rc2, o_null = oracle.perm_null(x, oracle.SCORE_S2N, yb, False, perm, off, idx, oracle.ES_WKS)
    assert rc == 0 and rc2 == 0
    want = oracle.calc_sig(o_obs[:, 0, 0], o_null[:, 0, :], True, True, False)
    assert np.abs(obs - o_obs[:, 0, 0]).max() <= ES_TOL
    assert np.array_equal(got["p_counts"], want["p_counts"])
    assert np.array_equal(got["fdr_counts"], want["fdr_counts"])
    assert np.array_equal(got["glob_counts"], want["glob_counts"])
    assert np.allclose(got["p"], want["p"], rtol=1e-14, atol=0)
    assert np.allclose(got["fdr"], want["fdr"], rtol=1e-10, atol=0)


@pytest.mark.parametrize("exact_small", [0, 1])
@pytest.mark.parametrize("es_kind", ["wks", "maxmean", "mean"])
def test_two_responses_small_n_pipeline(ctx, oracle, es_kind, exact_small):
    """two responses, six samples: the observed ES handed to the null loop is indexed by
    (response, set); p / FDR / global counts of every response match the oracle pipeline,
    on the production kernels (+ exact redo of the flagged units) and on the all-exact
    small-problem path"""
    G, n, S, P = 400, 6, 30, 300
    off, idx = synth.gene_sets_csr(G, S, (5, 60), "uniform", seed=31)
    x, y0 = synth.two_class(G, n, off, idx, seed=32)
    y = np.column_stack([y0[:, 0], np.array([1, 0, 1, 0, 0, 1])]).astype(np.int32)
    perm = synth.perm_matrix(n, P, seed=33)
    k = {"wks": nat.ES_WEIGHTED_KS, "maxmean": nat.ES_MAXMEAN, "mean": nat.ES_MEAN}[es_kind]
    ok = {"wks": oracle.ES_WKS, "maxmean": oracle.ES_MAXMEAN, "mean": oracle.ES_MEAN}[es_kind]
    ctx.set_option("es_exact_small", exact_small)
    try:
        ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
        sc = ctx.score_observed(nat.SCORE_S2N)
        obs = ctx.observed_es(k, sc, ragged=False)["es"]
        if exact_small == 0:  # same hand-over through the explicit entry point
            ctx.set_gene_sets(off, idx)      # (forgets the observed ES)
            ctx.set_observed_es(k, obs)
        ctx.perm_null(nat.SCORE_S2N, k, P, perm=perm, want_host=False)
        o_sc = oracle.s2n_post(oracle.s2n_C(x, y))
        assert np.array_equal(sc, o_sc)
        rc, o_obs = oracle.es_from_scores(ok, o_sc.reshape(G, 2, 1), off, idx)
        rc2, o_null = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, ok)
        assert rc == 0 and rc2 == 0
        assert np.abs(obs - o_obs[:, :, 0]).max() <= ES_TOL
        for r in range(2):
            got = ctx.calc_sig(obs[:, r], r, True, True, False)
            want = oracle.calc_sig(o_obs[:, r, 0], o_null[:, r, :], True, True, False)
            assert np.array_equal(got["p_counts"], want["p_counts"]), r
            assert np.array_equal(got["fdr_counts"], want["fdr_counts"]), r
            assert np.array_equal(got["glob_counts"], want["glob_counts"]), r
            assert np.allclose(got["fdr"], want["fdr"], rtol=1e-10, atol=0, equal_nan=True)
    finally:
        ctx.set_option("es_exact_small", 0)


def test_fdr_reference_properties_on_device(ctx):
    """tests/testthat/test_fdr.R:20-54 through the device path."""
    from test_oracle_golden import _fdr_inputs
    es, en = _fdr_inputs()
    for fn in (fg.flexgsea_calc_sig, fg.flexgsea_calc_sig_simple):
        sig = fn(es, en)
        for k in ("p", "fdr", "fwer"):
            assert sig[k].min() > 0 and sig[k].max() <= 1
        assert sig["p"][3] <= .1 and sig["p"][4] <= .1 and sig["fdr"][3] <= .1 and sig["fdr"][4] <= .1
        assert all(sig["fdr"][i] > .1 for i in (0, 1, 2, 5, 6, 7, 8, 9))


def test_staged_sig_equals_single(ctx):
    """the multi-GPU staged significance (two shards emulated on one GPU)
    reproduces the single-pass result exactly."""
    rng = np.random.default_rng(33)
    S, P = 120, 500
    es = rng.normal(size=S); en = rng.normal(size=(S, 1, P))
    ctx.set_null(en)
    one = ctx.calc_sig(es)
    parts, counts, ncs, gls = [], [], [], []
    shards = [en[:, :, :230], en[:, :, 230:]]
    for sh in shards:
        ctx.set_null(sh)
        s, c = ctx.sig_stage1(es)
        parts.append(s); counts.append(c)
    total = counts[0] + counts[1]
    for sh in shards:
        ctx.set_null(sh)
        r = ctx.sig_stage2(es, np.stack(parts), total, P)
        ncs.append(r["null_counts"]); gls.append(r["glob"])
    fin = ctx.sig_finish(r["nes"], ncs[0] + ncs[1], gls[0] + gls[1])
    assert np.array_equal(r["p"], one["p"]) and np.array_equal(r["nes"], one["nes"])
    assert np.array_equal(fin["fdr_counts"], one["fdr_counts"])
    assert np.array_equal(fin["fdr"], one["fdr"])


# ------------------------------------------ BASELINE.json sizes (full G, S)
def test_c2_full_size_parity(ctx, oracle):
    """BASELINE configs[1] at its real shape (20,000 genes x 200 samples, 8,000
    sets of 15-500 genes) on 64 of the permutations: ranks bit-exact, ES within
    1e-10, against the oracle on the same permutation matrix."""
    c = synth.CONFIGS["C2"]
    off, idx = synth.gene_sets_csr(c["G"], c["S"], c["sizes"], c["size_dist"], seed=1)
    x, y = synth.two_class(c["G"], c["n"], off, idx, seed=2)
    P = 64
    perm = synth.perm_matrix(c["n"], P, seed=4)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
    gsc, grk = ctx.debug_last(P)
    oracle.use_all_cores()
    rc, want, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS,
                                    return_scores=True)
    assert rc == 0
    assert np.allclose(gsc, sc[:, 0, :], rtol=1e-11, atol=1e-13)
    assert np.array_equal(grk, _ranks(oracle, sc[:, 0, :]))
    assert np.abs(got - want).max() <= ES_TOL


def test_c2_full_nperm_properties(ctx):
    """size-independent properties at the full nperm = 10,000 of configs[1]:
    the null of two half shards (global permutation ids) equals the one-shot null
    bit for bit; every ES is finite and in [-1, 1]; p-values in (0, 1]."""
    c = synth.CONFIGS["C2"]
    off, idx = synth.gene_sets_csr(c["G"], c["S"], c["sizes"], c["size_dist"], seed=1)
    x, y = synth.two_class(c["G"], c["n"], off, idx, seed=2)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    P, seed = c["P"], 99
    full = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, seed=seed, perm_id0=0)
    assert np.isfinite(full).all() and np.abs(full).max() <= 1.0
    sc = ctx.score_observed(nat.SCORE_S2N)
    es = ctx.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=False)["es"][:, 0]
    sig = ctx.calc_sig(es)
    assert (sig["p"] > 0).all() and (sig["p"] <= 1).all() and (sig["fdr"] > 0).all() and (sig["fdr"] <= 1).all()
    assert int(sig["glob_counts"][2] + sig["glob_counts"][3]) <= c["S"] * P
    lo = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P // 2, seed=seed, perm_id0=0)
    assert np.array_equal(lo, full[:, :, :P // 2])
    hi = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P - P // 2, seed=seed, perm_id0=P // 2)
    assert np.array_equal(hi, full[:, :, P // 2:])


def test_c3_full_shape_lm_parity(ctx, oracle):
    """BASELINE configs[2] at its real shape: flexgsea_lm with a continuous response and three
    covariates (design of 5 columns, R = 4 score columns), 20,000 genes x 500 samples, 8,000 sets
    of 15-500 genes, on 24 of the permutations against the oracle on the same permutation matrix:
    scores within 1e-12 of their scale, ranks equal outside 1e-12 score groups, every null ES
    asserted (see _assert_lm_null)."""
    c = synth.CONFIGS["C3"]
    G, n, S = c["G"], c["n"], c["S"]
    off, idx = synth.gene_sets_csr(G, S, c["sizes"], c["size_dist"], seed=1)
    x, _ = synth.two_class(G, n, off, idx, seed=2)
    design = synth.continuous_response(n, k=3, seed=3)
    P = 24
    perm = synth.perm_matrix(n, P, seed=4)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_lm(design, True)
    assert ctx.n_response(nat.SCORE_LM) == 4
    oracle.use_all_cores()
    got_obs = ctx.score_observed(nat.SCORE_LM)
    rc, want_obs = oracle.lm_scores(x, design, True)
    assert rc == 0 and np.abs(got_obs - want_obs).max() <= 1e-12 * max(1.0, np.abs(want_obs).max())
    null = ctx.perm_null(nat.SCORE_LM, nat.ES_WEIGHTED_KS, P, perm=perm)
    gsc, grk = ctx.debug_last(P * 4)
    rc, ref, sc = oracle.perm_null(x, oracle.SCORE_LM, design, True, perm, off, idx, oracle.ES_WKS,
                                   return_scores=True)
    assert rc == 0 and null.shape == ref.shape == (S, 4, P)
    flat = sc.reshape(G, -1, order="F")
    assert np.abs(gsc - flat).max() <= 1e-12 * max(1.0, np.abs(flat).max())
    _assert_lm_null(oracle, grk, flat, null, ref, G, off)


@pytest.mark.parametrize("es_kind", ["maxmean", "mean"])
def test_c4_full_shape_mean_maxmean_parity(ctx, oracle, es_kind):
    """BASELINE configs[3] at its real shape (20,000 genes x 200 samples, 20,000 sets of 15-300
    genes) on 32 of the permutations: flexgsea_maxmean / flexgsea_mean against the oracle
    (R/flexgsea.R:725-767), scores within 1e-11 (tensor-core contraction), ES within 1e-10; a
    permutation count that leaves a partial last round (tail split on) and abs on the second half."""
    c = synth.CONFIGS["C4"]
    G, n, S = c["G"], c["n"], c["S"]
    off, idx = synth.gene_sets_csr(G, S, c["sizes"], c["size_dist"], seed=1)
    x, y = synth.two_class(G, n, off, idx, seed=2)
    P = 32
    perm = synth.perm_matrix(n, P, seed=4)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    oracle.use_all_cores()
    gk = nat.ES_MAXMEAN if es_kind == "maxmean" else nat.ES_MEAN
    ok = oracle.ES_MAXMEAN if es_kind == "maxmean" else oracle.ES_MEAN
    for abs_flag in (False, True):
        got = ctx.perm_null(nat.SCORE_S2N, gk, P, perm=perm, abs_flag=abs_flag)
        rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, ok, abs_flag=abs_flag)
        assert rc == 0 and got.shape == want.shape == (S, 1, P)
        assert np.abs(got - want).max() <= ES_TOL
        # sign decisions (maxmean: neg > pos, strict) agree wherever the ES is not within rounding of 0
        big = np.abs(want) > 1e-9
        assert np.array_equal(np.sign(got[big]), np.sign(want[big]))


def test_c5_shape_parity(ctx, oracle):
    """BASELINE configs[4] at its real shape (60,000 transcripts x 1,000 samples -- 32 label words
    per column in the score contraction --, 20,000 sets of 15-500) on 12 of the permutations:
    ranks bit-exact (score-range bucketed on-chip sort), ES within 1e-10 (weights gathered through
    L1/L2: a 60,000-gene column does not fit on chip)."""
    c = synth.CONFIGS["C5"]
    G, n, S = c["G"], c["n"], c["S"]
    off, idx = synth.gene_sets_csr(G, S, c["sizes"], c["size_dist"], seed=1)
    x, y = synth.two_class(G, n, off, idx, seed=2)
    P = 12
    perm = synth.perm_matrix(n, P, seed=4)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
    gsc, grk = ctx.debug_last(P)
    oracle.use_all_cores()
    rc, want, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS,
                                    return_scores=True)
    assert rc == 0
    assert np.allclose(gsc, sc[:, 0, :], rtol=1e-11, atol=1e-13)
    assert np.array_equal(grk, _ranks(oracle, sc[:, 0, :]))
    assert np.abs(got - want).max() <= ES_TOL
    # the significance stage at S = 20,000 (thresholds no longer fit next to the histogram on chip)
    es = ctx.observed_es(nat.ES_WEIGHTED_KS, ctx.score_observed(nat.SCORE_S2N), ragged=False)["es"][:, 0]
    sig = ctx.calc_sig(es)
    osig = oracle.calc_sig(es, got[:, 0, :])
    assert np.array_equal(sig["p_counts"], osig["p_counts"])
    assert np.array_equal(sig["fdr_counts"], osig["fdr_counts"])
    assert np.array_equal(sig["glob_counts"], osig["glob_counts"])


@pytest.mark.parametrize("es_kind", [nat.ES_WEIGHTED_KS, nat.ES_MAXMEAN, nat.ES_MEAN])
def test_tail_round_split_is_bit_identical(ctx, es_kind):
    """A block whose column count is not a multiple of the SM count ends in a partial round; the
    ES kernels split those columns over the idle SMs (option tail_split).  Every unit is computed by
    the same code either way: the null must be bit-identical with the split on and off."""
    x, y, perm, off, idx = _case(G=3000, n=24, S=300, P=148 + 37, sizes=(10, 400), dist="loguniform", seed=41)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    a = ctx.perm_null(nat.SCORE_S2N, es_kind, perm.shape[1], perm=perm)
    b37 = ctx.perm_null(nat.SCORE_S2N, es_kind, 37, perm=perm[:, :37])       # fewer columns than SMs
    ctx.set_option("tail_split", 0)
    try:
        b = ctx.perm_null(nat.SCORE_S2N, es_kind, perm.shape[1], perm=perm)
    finally:
        ctx.set_option("tail_split", 1)
    assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(b37, b[:, :, :37], equal_nan=True)


@pytest.mark.parametrize("G", [30000, 40000, 60000, 65534])
def test_long_column_user_scores(ctx, oracle, G):
    """columns too long for the on-chip sort in one piece (score-range buckets, each
    sorted on chip): ranks bit-exact on user scores with heavy ties, -0.0, a value
    shared by a third of the genes (a bucket that cannot fit: 64-bit fallback for
    that column only) and plain continuous columns; force_sort64 agrees."""
    rng = np.random.default_rng(G)
    P = 5
    sc = rng.normal(size=(G, 1, P))
    sc[:, 0, 1] = np.round(sc[:, 0, 1], 2)                 # heavy ties
    sc[rng.integers(0, G, G // 3), 0, 2] = 0.25            # one value for a third of the genes
    sc[::7, 0, 3] = -0.0; sc[1::7, 0, 3] = 0.0
    sc[:, 0, 4] = rng.normal(size=G) * 1e-3 + 1.0          # every gene inside a few key bins
    S = 40
    off, idx = synth.gene_sets_csr(G, S, (15, 300), "loguniform", seed=3)
    ctx.set_gene_sets(off, idx)
    got = ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
    _, grk = ctx.debug_last(P, G=G)
    oracle.use_all_cores()
    assert np.array_equal(grk, _ranks(oracle, sc[:, 0, :]))
    rc, want = oracle.es_from_scores(oracle.ES_WKS, sc, off, idx)
    assert rc == 0 and np.abs(got - want).max() <= ES_TOL
    ctx.set_option("force_sort64", 1)
    try:
        got64 = ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
        _, grk64 = ctx.debug_last(P, G=G)
    finally:
        ctx.set_option("force_sort64", 0)
    assert np.array_equal(grk, grk64) and np.array_equal(got, got64)


def test_large_gene_count_paths(ctx, oracle):
    """G > 29696 (configs[4] has 60,000 transcripts): ranks come from the bucketed
    on-chip sort and the ES kernel reads its weights from global memory."""
    G, n, S, P = 40000, 48, 120, 6
    off, idx = synth.gene_sets_csr(G, S, (15, 500), "loguniform", seed=7)
    x, y = synth.two_class(G, n, off, idx, seed=8)
    perm = synth.perm_matrix(n, P, seed=9)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
    _, grk = ctx.debug_last(P)
    oracle.use_all_cores()
    rc, want, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS,
                                    return_scores=True)
    assert rc == 0 and np.array_equal(grk, _ranks(oracle, sc[:, 0, :]))
    assert np.abs(got - want).max() <= ES_TOL


# ----------------------------------------------------- the R-API mirror
def test_flexgsea_end_to_end(ctx, oracle):
    """flexgsea() with flexgsea_s2n + flexgsea_weighted_ks (BASELINE configs[0]
    shape) against the oracle pipeline on the same permutation matrix."""
    import pandas as pd
    G, n, S, P = 1000, 20, 50, 200
    off, idx = synth.gene_sets_csr(G, S, (10, 50), "uniform", seed=1)
    x, y = synth.two_class(G, n, off, idx, seed=2)
    genes = ["g%d" % (i + 1) for i in range(G)]
    sets = {"set%d" % s: [genes[i - 1] for i in idx[off[s]:off[s + 1]]] for s in range(S)}
    perm = synth.perm_matrix(n, P, seed=4)
    res = fg.flexgsea(pd.DataFrame(x, columns=genes), np.where(y[:, 0] == 1, "a", "b"), sets, nperm=P,
                      verbose=False, perm=perm, return_values=["es_null", "leading_edge", "running_es_at"])
    assert list(res["table"].keys()) == ["Response 1"]
    tab = res["table"]["Response 1"]
    assert list(tab.columns) == ["es", "p", "nes", "fdr", "fwer", "max_es_at", "le_prop", "GeneSet"]
    assert list(tab["GeneSet"]) == list(sets.keys())
    rc, null = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS)
    assert np.abs(res["es_null"] - null).max() <= ES_TOL
    sc = oracle.s2n_post(oracle.s2n_C(x, y))
    rank = oracle.rank_desc(sc[:, 0])
    es = np.array([oracle.wks_column(sc[:, 0], rank, idx[off[s]:off[s + 1]]) for s in range(S)])
    assert np.abs(tab["es"].to_numpy() - es).max() <= 1e-12
    sig = oracle.calc_sig(tab["es"].to_numpy(), res["es_null"][:, 0, :])
    assert np.array_equal(tab["p"].to_numpy(), sig["p"])
    assert np.abs(tab["nes"].to_numpy() - sig["nes"]).max() <= ES_TOL
    assert np.allclose(tab["fdr"].to_numpy(), sig["fdr"], rtol=1e-12)
    assert ((tab["p"] >= 0) & (tab["p"] <= 1)).all()
    le = res["leading_edge"]["Response 1"]["set0"]
    w = oracle.wks_column(sc[:, 0], rank, idx[off[0]:off[1]], extras=True)
    assert np.array_equal(le, w["leading_edge"])


def test_flexgsea_lm_structure(ctx):
    """tests/testthat/test_main.R: matrix / data.frame / model.matrix y with
    flexgsea_lm, maxmean and weighted KS, single gene set, extras."""
    import pandas as pd
    x = pd.DataFrame({
        "x1": [-0.91036224, -0.23633198, 0.07675882, -0.14270400],
        "x2": [-0.0005694878, -0.9734813268, -2.0230448058, 0.3552175208],
        "x2.big": [-0.005694878, -9.734813268, -20.230448058, 3.552175208],
        "x3": [0.1979443, -1.0453220, -0.1528467, 0.0897873],
        "x4": [-0.1979443, 1.0453220, 0.1528467, -0.0897873],
        "x5": [0.1979672, -1.0453921, -0.1528823, 0.0897534]})
    y = pd.DataFrame({"y1": [0.19801108, -1.05384953, -0.15942901, 0.09702427],
                      "y2": [1.793303, 1.589571, -1.412860, 2.192915]})
    gs = {"pw1": ["x1", "x2", "x3"], "pw2": ["x2", "x2.big"], "pw3": ["x3", "x5"]}
    mm = fg.ModelMatrix(np.column_stack([np.ones(4), y.to_numpy()]), ["(Intercept)", "y1", "y2"])
    for yy in (y.to_numpy(), y, mm):
        for esf in (fg.flexgsea_maxmean, fg.flexgsea_weighted_ks):
            res = fg.flexgsea(x, yy, gs, es_fn=esf, nperm=100, verbose=False, gs_size_min=1, block_size=11,
                              gene_score_fn=fg.flexgsea_lm, return_values=["es_null"], seed=1)
            names = ["Response1", "Response2"] if isinstance(yy, np.ndarray) else ["y1", "y2"]
            assert list(res["table"].keys()) == names
            for t in res["table"].values():
                assert list(t["GeneSet"]) == list(gs.keys())
                assert ((t["p"] >= 0) & (t["p"] <= 1)).all()
            assert res["es_null"].shape == (3, 2, 100)
    res = fg.flexgsea(x, y, {"pw1": gs["pw1"]}, nperm=100, verbose=False, gs_size_min=1,
                      gene_score_fn=fg.flexgsea_lm, return_values=["es_null", "running_es_pos"], seed=2)
    assert res["es_null"].shape == (1, 2, 100)
    assert isinstance(res["running_es_pos"], dict)


@pytest.mark.gpu
@pytest.mark.parametrize("G", [3001, 3000, 20000])
def test_bulk_copy_staging_same_bits(ctx, oracle, G):
    """The ES kernels stage a column's weights / scores with one bulk copy by the copy engine
    (cp.async.bulk + mbarrier; option bulk_stage) instead of a load / store loop per thread.  More
    columns than SMs, so every CTA takes several columns and the barrier's phase flips; odd G (the
    last weight goes by hand; the mean kernels' columns are then not 16-byte aligned and keep the
    loop).  Same bits either way, and the oracle's values."""
    rng = np.random.default_rng(77 + G)
    P = 330 if G < 20000 else 160
    sc = np.asfortranarray(rng.normal(size=(G, 1, P)))
    off, idx = synth.gene_sets_csr(G, 60, (5, 600), "loguniform", seed=5)
    ctx.set_gene_sets(off, idx)
    try:
        for kind, ok in ((nat.ES_WEIGHTED_KS, oracle.ES_WKS), (nat.ES_MAXMEAN, oracle.ES_MAXMEAN),
                         (nat.ES_MEAN, oracle.ES_MEAN)):
            ctx.set_option("bulk_stage", 1)
            a = ctx.es_from_scores(kind, sc)
            ctx.set_option("bulk_stage", 0)
            b = ctx.es_from_scores(kind, sc)
            assert a.tobytes() == b.tobytes(), kind
            if G < 20000:
                rc, want = oracle.es_from_scores(ok, sc[:, :, :40], off, idx)
                assert rc == 0 and np.abs(a[:, :, :40] - want).max() <= ES_TOL
    finally:
        ctx.set_option("bulk_stage", 1)


@pytest.mark.gpu
def test_user_scores_stream_in_blocks(ctx, oracle):
    """user score columns arrive block by block (upload of block k+1 while block k is ranked,
    two score buffers): every block count, incl. a ragged last block, gives the one-block result"""
    rng = np.random.default_rng(41)
    G, R, P = 700, 2, 23
    sc = rng.normal(size=(G, R, P))
    sc[:, 1, 5] = np.round(sc[:, 1, 5], 1)
    off, idx = synth.gene_sets_csr(G, 30, (5, 90), "uniform", seed=9)
    ctx.set_gene_sets(off, idx)
    rc, want = oracle.es_from_scores(oracle.ES_WKS, sc, off, idx)
    assert rc == 0
    try:
        for blk in (0, 1, 4, 5, 23, 64):      # score columns per block (0: default); 46 columns in all
            ctx.set_option("perm_block", blk)
            got = ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
            assert np.abs(got - want).max() <= ES_TOL, blk
            _, rk = ctx.debug_last(1, G=G)     # the resident buffer is the last block's
        ctx.set_option("perm_block", 3)
        for kind, ok in ((nat.ES_MEAN, oracle.ES_MEAN), (nat.ES_MAXMEAN, oracle.ES_MAXMEAN)):
            g2 = ctx.es_from_scores(kind, sc)
            _, w2 = oracle.es_from_scores(ok, sc, off, idx)
            assert np.abs(g2 - w2).max() <= ES_TOL
        bad = sc.copy(); bad[3, 1, P - 1] = np.nan    # in the last block
        with pytest.raises(fg.NonFiniteScoreError):
            ctx.es_from_scores(nat.ES_WEIGHTED_KS, bad)
    finally:
        ctx.set_option("perm_block", 0)


@pytest.mark.gpu
def test_user_scores_threaded_upload(ctx, oracle):
    """blocks of 32 MB and more go up through the pinned staging ring filled by several host
    threads; same ES as the plain copy, and as the oracle"""
    rng = np.random.default_rng(43)
    G, P = 20000, 560
    sc = np.asfortranarray(rng.normal(size=(G, 1, P)))
    off, idx = synth.gene_sets_csr(G, 40, (15, 300), "loguniform", seed=2)
    ctx.set_gene_sets(off, idx)
    try:
        ctx.set_option("perm_block", 250)          # 40 MB blocks, ragged last block of 9.6 MB
        res = {}
        for th in (0, 1, 3, 8):
            ctx.set_option("h2d_threads", th)
            res[th] = ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
        for th in (1, 3, 8):
            assert np.array_equal(res[th], res[0]), th
        cols = [0, 249, 250, 499, 500, 559]
        rc, want = oracle.es_from_scores(oracle.ES_WKS, np.asfortranarray(sc[:, :, cols]), off, idx)
        assert rc == 0 and np.abs(res[3][:, :, cols] - want).max() <= ES_TOL
    finally:
        ctx.set_option("perm_block", 0)
        ctx.set_option("h2d_threads", 8)


@pytest.mark.gpu
def test_null_streams_back_in_blocks(ctx, oracle):
    """es.null returns to the host block by block while the next block is computed; blocks of
    32 MB and more through the pinned ring and several host threads.  Same array as the plain
    copy, for the permutation loop and for user scores, and as the oracle on sampled columns."""
    G, n, S, P = 300, 24, 2000, 9000
    off, idx = synth.gene_sets_csr(G, S, (5, 40), "uniform", seed=4)
    x, y = synth.two_class(G, n, seed=3)
    perm = synth.perm_matrix(n, P, seed=5)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    sc = np.asfortranarray(np.random.default_rng(8).normal(size=(G, 1, P)))
    try:
        res, res_u = {}, {}
        for th in (0, 3, 8):
            ctx.set_option("h2d_threads", th)
            res[th] = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
            res_u[th] = ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
        for th in (3, 8):
            assert np.array_equal(res[th], res[0]) and np.array_equal(res_u[th], res_u[0]), th
        cols = [0, 4143, 4144, 8287, 8288, 8999]          # block edges of the default 4 144
        rc, want = oracle.es_from_scores(oracle.ES_WKS, np.asfortranarray(sc[:, :, cols]), off, idx)
        assert rc == 0 and np.abs(res_u[8][:, :, cols] - want).max() <= ES_TOL
        rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, np.asfortranarray(perm[:, cols]), off, idx, oracle.ES_WKS)
        assert rc == 0 and np.abs(res[8][:, :, cols] - want).max() <= ES_TOL
    finally:
        ctx.set_option("h2d_threads", 8)


@pytest.mark.gpu
def test_progress_hook_reports_and_cancels(ctx, oracle):
    x, y = synth.two_class(300, 24, seed=3)
    off, idx = synth.gene_sets_csr(300, 12, (5, 40), "uniform", seed=4)
    perm = synth.perm_matrix(24, 40, seed=5)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    want = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 40, perm=perm)
    seen = []
    try:
        ctx.set_option("perm_block", 8)
        ctx.set_progress(lambda done, total: seen.append((done, total)) or False)
        got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 40, perm=perm)
        assert np.array_equal(got, want)
        assert len(seen) == 5 and all(t == 40 for _, t in seen)
        assert [d for d, _ in seen] == sorted(d for d, _ in seen) and seen[-1][0] <= 40
        ctx.set_progress(lambda done, total: done >= 8)
        with pytest.raises(fg.FlexgseaError) as ei:
            ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 40, perm=perm)
        assert ei.value.code == 7
        sc = np.random.default_rng(1).normal(size=(300, 1, 40))
        with pytest.raises(fg.FlexgseaError) as ei:
            ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
        assert ei.value.code == 7
        ctx.set_progress(None)
        assert np.array_equal(ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 40, perm=perm), want)
    finally:
        ctx.set_progress(None)
        ctx.set_option("perm_block", 0)


def test_near_duplicated_genes_take_the_block_fallback(ctx, oracle):
    """Thousands of genes that are copies of others up to the last bits (a rescaled copy: s2n is
    scale invariant) nearly tie in every permutation: the near-tie list of the tensor-core path
    overflows and the BLOCK is rescored by the bit-faithful kernel and ranked again (device-side,
    no whole-call rerun).  Ranks bit-exact, scores bit-identical to the reference arithmetic in
    that block, ES within 1e-10.  Bit-identical copies need none of that: the reference ties them
    exactly too, index order decides, and the block keeps the tensor path."""
    G0, ndup, n, S, P = 4000, 2200, 20, 60, 2072
    off, idx = synth.gene_sets_csr(G0 + ndup, S, (10, 300), "loguniform", seed=51)
    x0, y = synth.two_class(G0, n, seed=52)
    perm = synth.perm_matrix(n, P, seed=54)
    oracle.use_all_cores()
    for scale, want_fallback in ((1.0 + 2.0 ** -30, 2), (1.0, 0)):
        x = np.asfortranarray(np.concatenate([x0, x0[:, :ndup] * scale], axis=1))
        ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
        assert ctx.get_stat("duplicated_rows") == (1 if scale == 1.0 else 0)
        ctx.set_option("perm_block", 1036)        # two blocks, each 2 * 2200 * 1036 = 4.6 M entries > the 4 M list
        try:
            got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
            gsc, grk = ctx.debug_last(P - 1036)   # the last block
        finally:
            ctx.set_option("perm_block", 0)
        assert ctx.get_stat("fallback_blocks") == want_fallback
        rc, want, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS,
                                        return_scores=True)
        assert rc == 0
        ref_last = sc[:, 0, 1036:]
        assert np.array_equal(grk, _ranks(oracle, ref_last))
        if want_fallback:
            assert np.array_equal(gsc, ref_last)  # the fallback block holds the reference's own bits
        else:
            assert np.allclose(gsc, ref_last, rtol=1e-11, atol=1e-13)
        assert np.abs(got - want).max() <= ES_TOL


@pytest.mark.parametrize("es_kind", ["wks", "maxmean"])
def test_few_samples_many_genes_dedup_counts_identical(oracle, es_kind):
    """a LARGE problem with very few samples (4 samples: six labellings;
    20,000 genes, 2,000 sets).  Every distinct labelling is computed once, in the reference's
    operation order, and copied to the permutations that share it: p AND pooled-FDR counts are
    identical to the oracle's, the null is bit-identical across repeated labellings."""
    G, n, S, P = 20000, 4, 2000, 1500
    off, idx = synth.gene_sets_csr(G, S, (15, 300), "loguniform", seed=61)
    x, y = synth.two_class(G, n, off, idx, seed=62)
    perm = synth.perm_matrix(n, P, seed=64)
    gk = nat.ES_WEIGHTED_KS if es_kind == "wks" else nat.ES_MAXMEAN
    ok = oracle.ES_WKS if es_kind == "wks" else oracle.ES_MAXMEAN
    c = fg.Context(0)      # default options: es_exact_small and dedup_labels on
    try:
        c.set_x(x); c.set_gene_sets(off, idx); c.set_response_s2n(y)
        sc = c.score_observed(nat.SCORE_S2N)
        es = c.observed_es(gk, sc, ragged=False)["es"][:, 0]
        null = c.perm_null(nat.SCORE_S2N, gk, P, perm=perm)
        assert c.get_stat("dedup_labellings") == 6
        sig = c.calc_sig(es)
        oracle.use_all_cores()
        rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, ok)
        assert rc == 0 and np.abs(null - want).max() <= ES_TOL
        osig = oracle.calc_sig(es, want[:, 0, :])
        assert np.array_equal(sig["p_counts"], osig["p_counts"])
        assert np.array_equal(sig["fdr_counts"], osig["fdr_counts"])
        assert np.array_equal(sig["glob_counts"], osig["glob_counts"])
        # permutations that give the same labelling hold the same column, bit for bit
        lab = y[:, 0][perm - 1]                                   # [n x P] labels after permutation
        _, first, inv = np.unique(lab.T, axis=0, return_index=True, return_inverse=True)
        assert np.array_equal(null, null[:, :, first[inv.ravel()]], equal_nan=True)
        # the Philox path and a sharded call deduplicate the same way
        c.set_option("dedup_labels", 0)
        plain = c.perm_null(nat.SCORE_S2N, gk, P, perm=perm)
        assert c.get_stat("dedup_labellings") == 0
        assert np.abs(plain - null).max() <= ES_TOL
    finally:
        c.close()


# ------------------------------------ f3: user score functions in tiles, EList, limma-style client
def test_user_score_tiles_equal_one_shot(ctx, oracle):
    """fgs_es_from_scores_tile + fgs_null_finish: the tiles of a user gene.score.fn give the null the
    one-shot call gives, bit for bit, the resident null serves calc_sig, and a non-finite score in
    any tile is reported at the finish with the reference's meaning (R/flexgsea.R:400-402)."""
    rng = np.random.default_rng(71)
    G, R, P, S = 3000, 2, 230, 150
    off, idx = synth.gene_sets_csr(G, S, (10, 300), "loguniform", seed=72)
    sc = rng.normal(size=(G, R, P))
    sc[:, 0, 7] = np.round(sc[:, 0, 7], 1)                     # ties
    ctx.set_gene_sets(off, idx)
    one = ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc)
    for tile in (1, 37, 100, 230):
        for t0 in range(0, P, tile):
            ctx.es_from_scores_tile(nat.ES_WEIGHTED_KS, np.asfortranarray(sc[:, :, t0:t0 + tile]), t0, P)
        got = ctx.null_finish()
        assert np.array_equal(got, one, equal_nan=True), tile
    es = rng.normal(size=S) * 0.3
    a = ctx.calc_sig(es, 1)
    ctx.set_null(one[:, 1, :])
    b = ctx.calc_sig(es, 0)
    assert np.array_equal(a["p_counts"], b["p_counts"]) and np.array_equal(a["fdr_counts"], b["fdr_counts"])
    bad = sc.copy(); bad[5, 1, 200] = np.inf
    for t0 in range(0, P, 100):
        ctx.es_from_scores_tile(nat.ES_WEIGHTED_KS, np.asfortranarray(bad[:, :, t0:t0 + 100]), t0, P)
    with pytest.raises(fg.NonFiniteScoreError):
        ctx.null_finish()
    with pytest.raises(fg.FlexgseaError):     # a tile out of order: the null was never started
        ctx.set_gene_sets(off[:50], idx[:off[49]])
        ctx.es_from_scores_tile(nat.ES_WEIGHTED_KS, np.asfortranarray(sc[:, :, 10:20]), 10, P)
        ctx.null_finish()                     # (a failed tile is reported by the next call)


def test_tiles_hide_the_device_behind_a_slow_host_score_fn(ctx):
    """SURVEY 8f row 3 / VERDICT r1 item 8: a host score function that costs 5 ms per permutation
    (limma's moderated t costs about that at 20,000 genes).  Handing tile k over only queues its
    upload, ranking and ES, so that work runs while the host scores tile k + 1: at least 80 % of
    the device / transfer time of the tiled run disappears behind the host function."""
    import time
    c = synth.CONFIGS["C5"]      # 60,000 transcripts, 20,000 sets: about 75 us of device / transfer work per column
    G, S = c["G"], c["S"]
    off, idx = synth.gene_sets_csr(G, S, c["sizes"], c["size_dist"], seed=1)
    rng = np.random.default_rng(73)
    P, tile = 320, 16
    tiles = [np.asfortranarray(rng.normal(size=(G, 1, tile))) for _ in range(P // tile)]
    ctx.set_gene_sets(off, idx)

    def run(host_ms):
        t_host = 0.0
        t0 = time.perf_counter()
        for k, sc in enumerate(tiles):
            h0 = time.perf_counter()
            while (time.perf_counter() - h0) * 1e3 < host_ms * tile:   # the host "scores" this tile
                pass
            t_host += time.perf_counter() - h0
            ctx.es_from_scores_tile(nat.ES_WEIGHTED_KS, sc, k * tile, P)
        ctx.null_finish(want_host=False)
        return time.perf_counter() - t0, t_host

    run(0.0)                                   # warm-up (allocations, first launches)
    t_dev = min(run(0.0)[0] for _ in range(3))   # the tiled device / transfer work on its own
    best = -1.0
    for _ in range(3):                         # (wall-clock based: the best of three keeps a busy host out of it)
        total, t_host = run(5.0)
        exposed = total - t_host
        best = max(best, 1.0 - exposed / t_dev)
        if best >= 0.8:
            break
    assert best >= 0.8, "device time %.1f ms, exposed %.1f ms of it (hidden %.0f %%)" % (
        1e3 * t_dev, 1e3 * exposed, 100 * best)


def test_flexgsea_with_elist_and_limma_style_client(ctx):
    """x as a limma EList (t.x = TRUE, R/flexgsea.R:159-173) with the limma adaptor as
    gene.score.fn (R/flexgsea_limma.R:15-31): moderated-t columns are made on the host per
    permutation, shipped in tiles of block.size, ranked and scored on the device; the result is
    the one the untiled user-score path gives, and the built-in lm scores on the same EList use
    the transposed layout."""
    rng = np.random.default_rng(75)
    G, n, S = 1200, 18, 60
    names = ["g%d" % i for i in range(G)]
    off, idx = synth.gene_sets_csr(G, S, (10, 120), "loguniform", seed=76)
    sets = {"s%d" % k: [names[i - 1] for i in idx[off[k]:off[k + 1]]] for k in range(S)}
    E = rng.normal(size=(G, n)) + rng.uniform(4, 12, size=(G, 1))
    el = fg.EList(E, genes=names, weights=rng.uniform(0.5, 2.0, size=(G, n)))
    D = np.column_stack([np.ones(n), rng.normal(size=n), (np.arange(n) % 2).astype(float)])
    mm = fg.ModelMatrix(D, ["(Intercept)", "dose", "batch"])
    kw = dict(gene_score_fn=fg.flexgsea_limma, nperm=90, gs_size_min=1, gs_size_max=500, verbose=False,
              seed=5, rng="R", return_values=["es_null"])
    a = fg.flexgsea(el, mm, sets, block_size=25, **kw)         # four tiles
    b = fg.flexgsea(el, mm, sets, block_size=1000, **kw)       # one tile
    assert list(a["table"]) == ["dose", "batch"]
    assert np.array_equal(a["es_null"], b["es_null"], equal_nan=True)
    for r in ("dose", "batch"):
        for col in ("es", "p", "nes", "fdr"):
            assert np.array_equal(np.asarray(a["table"][r][col]), np.asarray(b["table"][r][col]), equal_nan=True)
    # the observed moderated t drives the observed ES: same as scoring the EList directly
    t_obs = fg.flexgsea_limma(el, mm).to_numpy()
    ctx.set_gene_sets(off, idx)
    es = ctx.observed_es(nat.ES_WEIGHTED_KS, t_obs, ragged=False)["es"]
    assert np.allclose(np.asarray(a["table"]["dose"]["es"]), es[:, 0], rtol=0, atol=1e-12)
    # built-in lm on an EList: genes x samples is transposed for the device
    c = fg.flexgsea(el, mm, sets, gene_score_fn=fg.flexgsea_lm, nperm=40, gs_size_min=1, gs_size_max=500,
                    verbose=False, seed=6)
    d = fg.flexgsea(E.T, mm, sets, gene_names=names, gene_score_fn=fg.flexgsea_lm, nperm=40, gs_size_min=1,
                    gs_size_max=500, verbose=False, seed=6)
    assert np.array_equal(np.asarray(c["table"]["dose"]["fdr"]), np.asarray(d["table"]["dose"]["fdr"]))


def test_custom_es_fn_plugin_with_default_arguments(ctx):
    """a user-defined es.fn record (list(run, prepare, extra_stats, extra),
    R/flexgsea.R:60-93) with a built-in gene.score.fn and the default rng: the host plugin loop
    needs the permutations themselves."""
    G, n, S = 300, 12, 8
    off, idx = synth.gene_sets_csr(G, S, (5, 30), "uniform", seed=81)
    x, y = synth.two_class(G, n, off, idx, seed=82)
    names = ["g%d" % i for i in range(G)]
    sets = {"s%d" % k: [names[i - 1] for i in idx[off[k]:off[k + 1]]] for k in range(S)}
    mean_fn = {"prepare": lambda gs: {}, "extra_stats": (), "extra": (),
               "run": lambda gs, gene_set, prep, return_stats=(), return_values=(): {
                   "es": gs[np.asarray(gene_set) - 1].mean(axis=0)}}
    a = fg.flexgsea(x, y[:, 0], sets, es_fn=mean_fn, gene_names=names, nperm=50, gs_size_min=1, verbose=False, seed=3)
    b = fg.flexgsea(x, y[:, 0], sets, es_fn=fg.flexgsea_mean, gene_names=names, nperm=50, gs_size_min=1,
                    verbose=False, seed=3, rng="R")
    ta, tb = a["table"]["Response 1"], b["table"]["Response 1"]
    assert np.allclose(np.asarray(ta["es"]), np.asarray(tb["es"]), rtol=0, atol=1e-12)
    assert np.array_equal(np.asarray(ta["p"]), np.asarray(tb["p"]))


def test_observed_es_is_copied_only_when_it_belongs_to_the_null(ctx, oracle):
    """label-preserving permutations get the observed ES copied verbatim -- but only an
    observed ES made from fgs_score_observed with the null loop's score kind, abs flag and weight
    exponent.  With few samples (many such permutations) an observed pass with abs = TRUE must not
    leak into a signed null."""
    x, y, perm, off, idx = _case(G=400, n=6, S=30, P=300, sizes=(5, 40), seed=91)
    ctx.set_option("dedup_labels", 0)
    try:
        ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
        sc_abs = ctx.score_observed(nat.SCORE_S2N, True)
        ctx.observed_es(nat.ES_WEIGHTED_KS, sc_abs, ragged=False)      # |scores|: not the signed null's observed ES
        got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, perm.shape[1], perm=perm, abs_flag=False)
    finally:
        ctx.set_option("dedup_labels", 1)
    rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS)
    assert rc == 0 and np.abs(got - want).max() <= ES_TOL
