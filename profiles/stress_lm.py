#!/usr/bin/env python
"""Randomised sweep of the lm score path (flexgsea_lm: design with intercept + covariates,
every non-intercept column a response) with weighted KS / maxmean / mean against the oracle:
   python profiles/stress_lm.py [n_cases] [seed]
lm scores come from a tensor-core GEMM, not from lm.fit's QR: scores to 1e-11 relative,
ES to 1e-9 (a pair of genes scored within rounding of each other may swap ranks)."""
import sys, os
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
import pyoracle as oracle


def run(ncase, seed, ctx=None):
    rng = np.random.default_rng(seed)
    ctx = ctx or fg.Context(0)
    oracle.use_all_cores()
    bad = 0
    nflip = 0
    for case in range(ncase):
        G = int(rng.choice([30, 200, 1500]))
        n = int(rng.choice([8, 15, 40, 90]))
        k = int(rng.integers(1, 4))
        S = int(rng.integers(1, 20))
        P = int(rng.integers(1, 40))
        hi = max(1, min(G - 1, int(rng.choice([5, 60, 400]))))
        sets = [rng.choice(G, int(rng.integers(1, hi + 1)), replace=False) + 1 for _ in range(S)]
        off, idx = oracle.csr(sets)
        x = np.asfortranarray(rng.normal(size=(n, G)) * rng.choice([1.0, 30.0]) + rng.choice([0.0, 8.0]))
        design = np.column_stack([np.ones(n), rng.normal(size=(n, k))])
        perm = synth.perm_matrix(n, P, seed=int(rng.integers(1 << 30)))
        es_kind = int(rng.choice([nat.ES_WEIGHTED_KS, nat.ES_MAXMEAN, nat.ES_MEAN]))
        okind = {nat.ES_WEIGHTED_KS: oracle.ES_WKS, nat.ES_MAXMEAN: oracle.ES_MAXMEAN, nat.ES_MEAN: oracle.ES_MEAN}[es_kind]
        tag = dict(case=case, G=G, n=n, k=k, S=S, P=P, es=es_kind)
        ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_lm(design, True)
        rc, want, sc = oracle.perm_null(x, oracle.SCORE_LM, design, True, perm, off, idx, okind, return_scores=True)
        if rc != 0:
            continue
        got = ctx.perm_null(nat.SCORE_LM, es_kind, P, perm=perm)
        gsc, _ = ctx.debug_last(P * k)
        wsc = sc.reshape(G, -1, order="F")
        serr = np.abs(gsc - wsc).max() / max(1e-300, np.abs(wsc).max())
        ok = ~np.isnan(want)
        # es_pos == -es_neg exactly (structural tie, small universes): the sign then hangs on
        # the last bit of the weights, and lm scores are not bit-pinned (a GEMM here, a
        # Householder restatement in the oracle, LINPACK's dqrls in R) -- counted apart
        flip = ok & (np.abs(got + want) <= 1e-9) & (np.abs(got - want) > 1e-9)
        nflip += int(flip.sum())
        ok &= ~flip
        err = np.abs(got[ok] - want[ok]).max() if ok.any() else 0.0
        if got.shape != want.shape or not np.array_equal(np.isnan(got), np.isnan(want)) or serr > 1e-11 or err > 1e-9:
            bad += 1; print("MISMATCH", tag, "score rel err", serr, "ES err", err)
    print('sign differences at exact structural ties (unpinned for lm):', nflip)
    return bad


if __name__ == "__main__":
    n_case = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n_bad = run(n_case, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print("cases", n_case, "mismatches", n_bad)
