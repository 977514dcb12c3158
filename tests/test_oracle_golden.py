"""Pins the CPU oracle against everything the reference's own tests hold for
the hot path: the three .rds goldens of
test_weighted_ks.R, the verbatim-compiled s2n_C, and the test_fdr.R /
test_s2n.R / test_lm.R properties."""
import json
import os

import numpy as np
import pytest

import flexgsea_r_b200 as fg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "weighted_ks.json")))

# data of tests/testthat/test_s2n.R:5-14 and test_lm.R:5-14
X_TOY = np.array([
    [-0.91036224, -0.23633198, 0.07675882, -0.14270400],
    [-0.0005694878, -0.9734813268, -2.0230448058, 0.3552175208],
    [10 * -0.0005694878, 10 * -0.9734813268, 10 * -2.0230448058, 10 * 0.3552175208],
    [0.1979443, -1.0453220, -0.1528467, 0.0897873],
    [-0.1979443, 1.0453220, 0.1528467, -0.0897873]]).T  # [4 samples x 5 genes]


def test_rrng_reproduces_reference_inputs():
    rng = fg.RRNG(GOLD["seed"], "Rounding")
    sc = rng.rnorm(3000)
    assert np.allclose(sc[:3], [0.50424188, 1.65056542, -1.28302844], atol=1e-8)
    gs = rng.sample_int(1000, 20)
    assert list(gs) == GOLD["gene_set"]


def test_weighted_ks_golden(oracle):
    """tests/testthat/test_weighted_ks.R:17-45 against the decoded .rds."""
    rng = fg.RRNG(GOLD["seed"], "Rounding")
    scores = rng.rnorm(3000).reshape(3, 1000).T  # column-major fill [1000 x 3]
    gs = rng.sample_int(1000, 20)
    for r in range(3):
        rank = oracle.rank_desc(scores[:, r])
        res = oracle.wks_column(scores[:, r], rank, gs, extras=True)
        assert abs(res["es"] - GOLD["es"]["values"][r]) <= 1e-12
        assert int(res["max_es_at"]) == GOLD["max_es_at"]["values"][r]
        assert res["le_prop"] == GOLD["le_prop"]["values"][r]
        fast = oracle.wks_column(scores[:, r], rank, gs)
        assert fast == res["es"]
        assert len(res["leading_edge"]) == round(res["le_prop"] * 20)


def test_rank_tie_rule(oracle):
    """R/flexgsea.R:929-931: G - rank(s,'first') + 1; among ties the later
    index ranks earlier; -0.0 ties with 0.0."""
    s = np.array([1.0, 2.0, 2.0, -0.0, 0.0, 3.0])
    assert list(oracle.rank_desc(s)) == [4, 3, 2, 6, 5, 1]


def test_s2n_matches_verbatim_reference(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    y = np.array([[1, 0, 0, 1], [0, 0, 1, 1]], dtype=np.int32).T  # y1=='a', y2==0
    a = oracle.s2n_C(X_TOY, y)
    b = oracle.s2n_C(X_TOY, y, use_ref=True)
    assert np.array_equal(a, b)
    # values recorded in SURVEY.md section 4 from the verbatim s2n_C
    assert a[0, 0] == -0.8267350207408003
    assert a[1, 0] == 2.3845824903956858
    assert a[3, 0] == 1.4849613589327466 and a[4, 0] == -a[3, 0]
    assert a[0, 1] == 1.2095774037779901
    assert a[1, 1] == -0.20702489220693349 and a[2, 1] == -0.2070248922069334
    rng = np.random.default_rng(5)
    x = rng.normal(6, 2, size=(37, 211)); yb = (rng.random((37, 3)) < 0.4).astype(np.int32)
    yb[3, 1] = np.iinfo(np.int32).min  # NA_LOGICAL is skipped (src/ggsea.cpp:29)
    assert np.array_equal(oracle.s2n_C(x, yb), oracle.s2n_C(x, yb, use_ref=True))
    # row mismatch -> zeros (src/ggsea.cpp:15-17)
    assert not oracle.s2n_C(x, yb[:30]).any()
    assert not oracle.s2n_C(x, yb[:30], use_ref=True).any()


def test_s2n_properties(oracle):
    """tests/testthat/test_s2n.R:44-48: x2 and 10*x2 score equal (all.equal)."""
    y = np.array([[1, 0, 0, 1], [0, 0, 1, 1]], dtype=np.int32).T
    a = oracle.s2n_post(oracle.s2n_C(X_TOY, y))
    assert np.allclose(a[1], a[2], rtol=1.5e-8)
    ab = oracle.s2n_post(oracle.s2n_C(X_TOY, y), abs_flag=True)
    assert np.array_equal(ab, np.abs(a))


def test_s2n_inf_patch(oracle):
    """R/flexgsea.R:715-716."""
    c = np.array([[1.0, np.inf], [-2.0, -np.inf], [0.5, 3.0]])
    out = oracle.s2n_post(c)
    assert out[0, 1] == 3.0 * 1.1 and out[1, 1] == -2.0 * 1.1


def test_lm_properties(oracle):
    """tests/testthat/test_lm.R:44-62."""
    y = np.array([[0.19801108, -1.05384953, -0.15942901, 0.09702427],
                  [1.793303, 1.589571, -1.412860, 2.192915]]).T
    design = np.column_stack([np.ones(4), y])
    rc, sc = oracle.lm_scores(X_TOY, design, True)
    assert rc == 0 and sc.shape == (5, 2) and np.isfinite(sc).all()
    assert int(np.argmax(sc[:, 0])) == 3 and int(np.argmin(sc[:, 0])) == 4
    assert np.allclose(sc[1], sc[2], rtol=1.5e-8)
    # against numpy least squares
    beta = np.linalg.lstsq(design, X_TOY, rcond=None)[0][1:].T
    assert np.allclose(sc, beta / X_TOY.std(axis=0, ddof=1)[:, None], rtol=1e-10)
    rc1, sc1 = oracle.lm_scores(X_TOY, design[:, :2], True)
    assert rc1 == 0 and int(np.argmax(sc1[:, 0])) == 3
    # rank-deficient design -> NA in R
    rc2, _ = oracle.lm_scores(X_TOY, np.column_stack([design, design[:, 1]]), True)
    assert rc2 == oracle.ERR_NONFINITE


def _fdr_inputs():
    rng = fg.RRNG(7242, "Rounding")
    es = rng.rnorm(10)
    en = rng.rnorm(10 * 1000).reshape(1000, 10).T.copy()
    es[3] = en[3].min() - 1
    es[4] = en[4].max() + 1
    return es, en


@pytest.mark.parametrize("simple", [False, True])
def test_fdr_properties(oracle, simple):
    """tests/testthat/test_fdr.R:20-54 for flexgsea_calc_sig and _simple."""
    es, en = _fdr_inputs()
    sig = oracle.calc_sig(es, en, split_p=not simple, calc_nes=not simple)
    for k in ("p", "fdr", "fwer"):
        assert sig[k].min() > 0.0 and sig[k].max() <= 1.0
    assert sig["p"][3] <= .1 and sig["p"][4] <= .1 and sig["fdr"][3] <= .1 and sig["fdr"][4] <= .1
    for i in (0, 1, 2, 5, 6, 7, 8, 9):
        assert sig["fdr"][i] > .1
    if not simple:
        # restatement values recorded during the survey
        assert np.allclose(sig["p"], [.6189, .1202, .1823, .00209, .00193, .0536, .5246, .1389, .4067, .5467], atol=6e-5)
        assert np.allclose(sig["nes"], [.6228, 2.0639, -1.6817, -5.2378, 5.2677, 2.3963, -.8285, 1.8817, 1.0280, -.7562], atol=6e-5)
        assert np.allclose(sig["fdr"], [.6196, .2011, .3635, .000805, .001192, .1645, .5535, .2051, .5001, .5535], atol=6e-5)


def test_fdr_bruteforce_numpy(oracle):
    """calc_fdr_nes / adj_fdr_nes (R/flexgsea.R:474-533) against a direct
    numpy transcription on a random case with both signs."""
    rng = np.random.default_rng(11)
    S, P = 23, 57
    es = rng.normal(size=S); en = rng.normal(size=(S, P))
    sig = oracle.calc_sig(es, en, want_nes_null=True)
    mpos = np.array([e[e >= 0].mean() for e in en]); mneg = np.array([e[e < 0].mean() for e in en])
    nes = es / np.where(es >= 0, mpos, -mneg)
    nn = en / np.where(en >= 0, mpos[:, None], -mneg[:, None])
    assert np.allclose(sig["nes"], nes, rtol=1e-13) and np.allclose(sig["nes_null"], nn, rtol=1e-13)
    fdr = np.zeros(S)
    for i in range(S):
        if nes[i] >= 0:
            fdr[i] = ((nn > nes[i]).sum() + 1) / (nn > 0).sum() / (((nes > nes[i]).sum() + 1) / (nes > 0).sum())
        else:
            fdr[i] = ((nn < nes[i]).sum() + 1) / (nn < 0).sum() / (((nes < nes[i]).sum() + 1) / (nes < 0).sum())
    adj = fdr.copy()
    for sel, order in ((nes >= 0, lambda v: np.argsort(v, kind="stable")),
                       (nes < 0, lambda v: np.argsort(-v, kind="stable"))):
        idx = np.where(sel)[0]; mn = 1.0
        for k in order(nes[idx]):
            if adj[idx[k]] <= mn:
                mn = adj[idx[k]]
            else:
                adj[idx[k]] = mn
    assert np.allclose(sig["fdr"], adj, rtol=1e-12)
    p = np.array([((es[i] <= en[i][en[i] >= 0]).sum() + 1) / ((en[i] >= 0).sum() + 1) if es[i] >= 0 else
                  ((es[i] >= en[i][en[i] <= 0]).sum() + 1) / ((en[i] <= 0).sum() + 1) for i in range(S)])
    assert np.array_equal(sig["p"], p)
    assert np.allclose(sig["fwer"], np.minimum(1, p * S))


def test_maxmean_mean_small(oracle):
    """R/flexgsea.R:725-741, :755-767 on a hand-checkable case."""
    sc = np.array([1.0, -2.0, 3.0, -4.0, 0.5]).reshape(5, 1, 1)
    off, idx = oracle.csr([[1, 2, 3], [2, 4], [1, 2]])
    _, mm = oracle.es_from_scores(oracle.ES_MAXMEAN, sc, off, idx)
    _, mn = oracle.es_from_scores(oracle.ES_MEAN, sc, off, idx)
    assert np.allclose(mm[:, 0, 0], [4 / 3, -3.0, -1.0])
    assert np.allclose(mn[:, 0, 0], [2 / 3, -3.0, -0.5])


def test_philox_perm_is_permutation(oracle):
    for n in (1, 2, 7, 200):
        p = oracle.philox_perm(123, 45, n)
        assert sorted(p) == list(range(1, n + 1))
    assert not np.array_equal(oracle.philox_perm(1, 0, 50), oracle.philox_perm(1, 1, 50))
    assert np.array_equal(oracle.philox_perm(9, 3, 50), oracle.philox_perm(9, 3, 50))
