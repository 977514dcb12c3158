"""The R boundary without R: r-package/src/flexgsea_native.cpp is compiled
against a functional stand-in of the R / Rcpp API and its .Call routines are looked up in the
registered table and called the way R/flexgsea_native.R calls them -- same order, same
arguments.  CPU part: registration, arity of every call site in the R file, the perm.fun
signature, the patch to R/flexgsea.R.  GPU part: the whole flexgsea() sequence through .Call
against the oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import rstandin
from rstandin import ROOT

RFILE = os.path.join(ROOT, "r-package", "R", "flexgsea_native.R")


@pytest.fixture(scope="module")
def shim():
    return rstandin.Shim()


def test_registration_keeps_s2n_C_and_turns_dynamic_lookup_off(shim):
    r = shim.routines()
    assert r["_flexgsea_s2n_C"] == 2            # reference src/RcppExports.cpp:21-24
    assert shim.L.standin_dynamic_symbols() == 0   # R_useDynamicSymbols(dll, FALSE), :28
    assert len(r) >= 15 and all(k.startswith("_flexgsea_") for k in r)


def test_every_call_site_in_the_R_file_matches_a_registered_arity(shim):
    reg = shim.routines()
    sites = rstandin.r_call_sites(RFILE)
    assert len(sites) >= 12
    for name, nargs in sites:
        assert name in reg, "%s is called from R but not registered" % name
        assert reg[name] == nargs, "%s: R passes %d arguments, the routine takes %d" % (name, nargs, reg[name])
    called = {n for n, _ in sites}
    # everything registered is reachable from R (s2n_C through the reference's own R/RcppExports.R:4-6)
    assert set(reg) - called <= {"_flexgsea_s2n_C", "_flexgsea_session_score_observed_n", "_flexgsea_session_set_null",
                                 "_flexgsea_session_es_from_scores"}


def test_wrong_arity_and_unknown_routine_are_errors_like_in_R(shim):
    with pytest.raises(rstandin.RError, match="Incorrect number of arguments"):
        shim.call("_flexgsea_s2n_C", np.zeros((2, 2)))
    with pytest.raises(rstandin.RError, match="not available for .Call"):
        shim.call("_flexgsea_nope")
    with pytest.raises(rstandin.RError, match="not a GPU session"):
        shim.call("_flexgsea_session_response_s2n", 1.0, np.zeros((2, 1), dtype=bool))


def test_perm_fun_signature_is_the_one_flexgsea_calls():
    """R/flexgsea.R:305-307: perm.fun(x, y, gene.sets, gene.names, nperm, block.size, gene.score.fn,
    es.fn, responses, abs=abs, verbose=verbose)"""
    f = rstandin.r_formals(RFILE, "flexgsea_perm_native")
    assert f[:9] == ["x", "y", "gene.sets", "gene.names", "nperm", "block.size", "gene.score.fn", "es.fn",
                     "responses"]
    assert set(f[9:]) == {"abs", "verbose"}


def test_patch_applies_to_the_reference(tmp_path):
    ref = "/root/reference/R/flexgsea.R"
    if not os.path.exists(ref) or shutil.which("patch") is None:
        pytest.skip("reference not present")
    os.makedirs(tmp_path / "R")
    shutil.copy(ref, tmp_path / "R" / "flexgsea.R")
    out = subprocess.run(["patch", "-p1", "--fuzz=0", "-i", os.path.join(ROOT, "r-package", "flexgsea.R.patch")],
                         cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    txt = open(tmp_path / "R" / "flexgsea.R").read()
    assert txt.count("flexgsea_native_begin(") == 1 and txt.count("perm.fun <- flexgsea_perm_native") == 1
    assert txt.count("flexgsea_native_sig(") == 1 and txt.count("flexgsea_native_observed(") == 1
    assert txt.count("{") == txt.count("}")


def _ybin_like_R(y):
    """flexgsea_s2n_ybin in flexgsea_native.R: logical -> classes c(TRUE, FALSE); else sort(unique)[1]"""
    y = np.asarray(y)
    if y.dtype == np.bool_:
        return y.astype(np.int32)
    return np.stack([(y[:, r] == np.sort(np.unique(y[:, r]))[0]).astype(np.int32) for r in range(y.shape[1])], axis=1)


@pytest.mark.gpu
def test_flexgsea_sequence_through_dot_call(shim, oracle):
    """flexgsea_native_begin -> observed -> flexgsea_perm_native -> flexgsea_native_sig, as .Call
    sequences with the arguments the R code passes; against the oracle on the same permutation matrix."""
    from flexgsea_r_b200 import synth
    G, n, S, P = 1500, 24, 120, 200
    off, idx = synth.gene_sets_csr(G, S, (10, 200), "loguniform", seed=11)
    x, y = synth.two_class(G, n, off, idx, seed=12)
    ylog = y.astype(bool)                      # a logical y: class 1 = TRUE (not sort(unique)[1])
    perm = synth.perm_matrix(n, P, seed=14)
    ybin = _ybin_like_R(ylog)
    assert np.array_equal(ybin, y)

    assert shim.call("_flexgsea_n_gpu")[0] >= 1
    sess = shim.call("_flexgsea_session_create", 1)
    shim.call("_flexgsea_session_upload", sess, x, False, off.astype(np.int32), idx.astype(np.int32))
    shim.call("_flexgsea_session_option", sess, "nperm_total", float(P))
    shim.call("_flexgsea_session_response_s2n", sess, shim.logical(ybin))
    shim.call("_flexgsea_session_interruptible", sess, True)
    # observed scores: R's own flexgsea_s2n -> s2n_C, the arity-2 routine
    coef = shim.call("_flexgsea_s2n_C", x, shim.logical(ybin))
    assert np.array_equal(coef, oracle.s2n_C(x, ybin))
    o = shim.call("_flexgsea_session_observed_es", sess, 0, 1.0, coef, S, int(off[-1]), True)
    assert list(o) == ["es", "max_es_at", "le_prop", "running_es_at", "running_es_pos", "running_es_neg",
                       "le_first", "le_count"]
    rk = oracle.rank_desc(coef[:, 0])
    for s in (0, 7, S - 1):
        w = oracle.wks_column(coef[:, 0], rk, idx[off[s]:off[s + 1]], extras=True)
        assert abs(o["es"][s, 0] - w["es"]) <= 1e-12 and o["max_es_at"][s, 0] == w["max_es_at"]
        at = o["running_es_at"][off[s]:off[s + 1], 0]
        assert np.array_equal(at, w["running_es_at"])
        le = at[o["le_first"][s, 0]: o["le_first"][s, 0] + o["le_count"][s, 0]]
        assert np.array_equal(le, w["leading_edge"])
    null = shim.call("_flexgsea_session_perm_null", sess, 0, 0, 1.0, False, perm, 0.0, P, S, True)
    assert null.shape == (S, 1, P)
    rc, want = oracle.perm_null(x, oracle.SCORE_S2N, ybin, False, perm, off, idx, oracle.ES_WKS)
    assert rc == 0 and np.abs(null - want).max() <= 1e-10
    sig = shim.call("_flexgsea_session_calc_sig", sess, o["es"][:, 0].copy(), 0, True, True, False)
    assert list(sig) == ["es", "p", "nes", "fdr", "fwer"]         # tibble column order, R/flexgsea.R:628
    osig = oracle.calc_sig(o["es"][:, 0], null[:, 0, :])
    assert np.array_equal(sig["p"], osig["p"]) and np.allclose(sig["nes"], osig["nes"], rtol=0, atol=1e-10)
    assert np.allclose(sig["fdr"], osig["fdr"], rtol=1e-12, atol=0)
    simple = shim.call("_flexgsea_session_calc_sig", sess, o["es"][:, 0].copy(), 0, False, False, False)
    assert list(simple) == ["es", "p.high", "p.low", "p", "fdr", "fwer"]   # :599, :628
    # want_null = FALSE: nothing comes back, the null stays on the GPU for calc_sig
    none = shim.call("_flexgsea_session_perm_null", sess, 0, 0, 1.0, False, perm, 0.0, P, S, False)
    assert none.size == 0
    sig2 = shim.call("_flexgsea_session_calc_sig", sess, o["es"][:, 0].copy(), 0, True, True, False)
    assert np.array_equal(sig2["fdr"], sig["fdr"])
    # a pending Ctrl-C: the null loop stops between blocks and the interrupt is raised from R's side
    shim.call("_flexgsea_session_option", sess, "perm_block", 16.0)
    shim.set_interrupt(True)
    r = shim.call("_flexgsea_session_perm_null", sess, 0, 0, 1.0, False, perm, 0.0, P, S, False)
    assert r is None and shim.interrupt_raised()
    shim.set_interrupt(False)
    # the non-finite guard speaks the reference's words (R/flexgsea.R:400-402)
    xz = x.copy(); xz[:, 3] = 1.0; xz[:, 4] = 2.0
    yb2 = np.zeros((n, 1), dtype=np.int32); yb2[: n // 2] = 1
    shim.call("_flexgsea_session_upload", sess, xz * 0.0, False, off.astype(np.int32), idx.astype(np.int32))
    with pytest.raises(rstandin.RError, match="Gene scoring function returned NA or Inf"):
        shim.call("_flexgsea_session_perm_null", sess, 0, 0, 1.0, False, perm, 0.0, P, S, False)
    shim.call("_flexgsea_session_close", sess)
    with pytest.raises(rstandin.RError, match="released"):
        shim.call("_flexgsea_session_calc_sig", sess, o["es"][:, 0].copy(), 0, True, True, False)


@pytest.mark.gpu
def test_elist_layout_and_user_scores_through_dot_call(shim, oracle):
    """t.x = TRUE (an EList's $E is genes x samples, R/flexgsea.R:159-173) gives the same null; a
    user gene.score.fn's [G, R, P] scores go through _flexgsea_session_es_from_scores."""
    from flexgsea_r_b200 import synth
    G, n, S, P = 800, 16, 40, 60
    off, idx = synth.gene_sets_csr(G, S, (10, 80), "uniform", seed=21)
    x, y = synth.two_class(G, n, off, idx, seed=22)
    perm = synth.perm_matrix(n, P, seed=24)
    res = []
    for t_x in (False, True):
        sess = shim.call("_flexgsea_session_create", 1)
        shim.call("_flexgsea_session_upload", sess, x.T.copy() if t_x else x, t_x, off, idx)
        shim.call("_flexgsea_session_response_s2n", sess, shim.logical(y))
        res.append(shim.call("_flexgsea_session_perm_null", sess, 0, 0, 1.0, False, perm, 0.0, P, S, True))
        if t_x:
            rng = np.random.default_rng(5)
            sc = rng.normal(size=(G, 2, 30))
            got = shim.call("_flexgsea_session_es_from_scores", sess, 0, 1.0, sc, S, True)
            rc, want = oracle.es_from_scores(oracle.ES_WKS, sc, off, idx)
            assert rc == 0 and got.shape == (S, 2, 30) and np.abs(got - want).max() <= 1e-10
            # ... and in tiles of block.size, as flexgsea_perm_native hands them over
            for t0 in range(0, 30, 8):
                shim.call("_flexgsea_session_es_from_scores_tile", sess, 0, 1.0, np.asfortranarray(sc[:, :, t0:t0 + 8]),
                          float(t0), 30.0)
                assert shim.L.standin_preserved() == 1          # the tile in flight is kept from the collector
            tiled = shim.call("_flexgsea_session_null_finish", sess, S, 2, 30, True)
            assert shim.L.standin_preserved() == 0 and np.array_equal(tiled, got)
        shim.call("_flexgsea_session_close", sess)
    assert np.array_equal(res[0], res[1])
