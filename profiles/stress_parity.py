#!/usr/bin/env python
"""Randomised parity sweep (GPU vs the CPU oracle) over small problem shapes:
   python profiles/stress_parity.py [n_cases] [seed]
Not part of the test suite (it takes a minute); prints every mismatch."""
import sys, os
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
import pyoracle as oracle

def run(ncase, seed, ctx=None):
    """returns the number of mismatching cases (each is printed)"""
    rng = np.random.default_rng(seed)
    ctx = ctx or fg.Context(0)
    oracle.use_all_cores()
    if os.environ.get('EXACT_SMALL'):   # 0: small problems too go through the production kernels
        ctx.set_option('es_exact_small', int(os.environ['EXACT_SMALL']))
    bad = 0
    for case in range(ncase):
        G = int(rng.choice([7, 16, 33, 100, 257, 1000, 3000]))
        n = int(rng.choice([4, 6, 9, 16, 40]))
        S = int(rng.integers(1, 30))
        P = int(rng.integers(1, 60))
        hi = max(1, min(G, int(rng.choice([3, 20, 200, 1200]))))
        sets = [rng.choice(G, int(rng.integers(1, hi + 1)), replace=bool(rng.random() < 0.1)) + 1 for _ in range(S)]
        off, idx = oracle.csr(sets)
        x = np.asfortranarray(rng.normal(size=(n, G)) * rng.choice([1.0, 1e-3, 50.0]) + rng.choice([0.0, 8.0]))
        if rng.random() < 0.3:
            x = np.asfortranarray(np.round(x, 1))                      # ties in the data
        R = int(rng.choice([1, 2]))
        y = np.zeros((n, R), dtype=np.int32)
        for r in range(R):
            k = int(rng.integers(2, n - 1)); y[rng.choice(n, k, replace=False), r] = 1
        perm = synth.perm_matrix(n, P, seed=int(rng.integers(1 << 30)))
        es_kind = int(rng.choice([nat.ES_WEIGHTED_KS, nat.ES_WEIGHTED_KS, nat.ES_MAXMEAN, nat.ES_MEAN]))
        okind = {nat.ES_WEIGHTED_KS: oracle.ES_WKS, nat.ES_MAXMEAN: oracle.ES_MAXMEAN, nat.ES_MEAN: oracle.ES_MEAN}[es_kind]
        p_exp = float(rng.choice([1.0, 1.0, 0.0, 2.0]))
        abs_flag = bool(rng.random() < 0.3)
        mode = int(rng.choice([0, 1]))
        tag = dict(case=case, G=G, n=n, S=S, P=P, R=R, es=es_kind, p=p_exp, abs=abs_flag, mode=mode)
        if os.environ.get('ONLY') and int(os.environ['ONLY']) != case:
            continue
        try:
            ctx.set_option("s2n_mode", mode)
            ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
            rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, okind, p=p_exp, abs_flag=abs_flag)
            try:
                got = ctx.perm_null(nat.SCORE_S2N, es_kind, P, perm=perm, p=p_exp, abs_flag=abs_flag)
            except nat.NonFiniteScoreError:
                if rc != oracle.ERR_NONFINITE:
                    bad += 1; print("MISMATCH non-finite only on device", tag)
                    if os.environ.get('ONLY'):
                        rc2, w2, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, okind, p=p_exp, abs_flag=abs_flag, return_scores=True)
                        for md in (0, 1):
                            ctx.set_option("s2n_mode", md)
                            for pp in range(P):
                                try:
                                    ctx.perm_null(nat.SCORE_S2N, es_kind, 1, perm=perm[:, pp:pp + 1], p=p_exp, abs_flag=abs_flag)
                                except nat.NonFiniteScoreError:
                                    col = sc[:, :, pp]
                                    print('   mode', md, 'perm', pp, 'raises; oracle column max', np.abs(col).max(), 'argmax gene', np.unravel_index(np.abs(col).argmax(), col.shape), 'x of that gene', x[:, np.unravel_index(np.abs(col).argmax(), col.shape)[0]].tolist(), 'labels', y[perm[:, pp] - 1, :].T.tolist())
                                    break
                        print('  oracle scores finite:', np.isfinite(sc).all(), 'max', np.nanmax(np.abs(sc)), 'n inf', np.isinf(sc).sum(), 'n nan', np.isnan(sc).sum())
                continue
            if rc != 0:
                bad += 1; print("MISMATCH oracle rc", rc, tag); continue
            nan_ok = np.array_equal(np.isnan(got), np.isnan(want))
            ok = ~np.isnan(want)
            err = np.abs(got[ok] - want[ok]).max() if ok.any() else 0.0
            scale = max(1.0, float(np.abs(want[ok]).max())) if ok.any() else 1.0
            if not nan_ok or err > 1e-10 * scale:
                bad += 1; print("MISMATCH", tag, "max err", err, "nan pattern equal", nan_ok)
                if os.environ.get('ONLY'):
                    d = np.argwhere((np.isnan(got) != np.isnan(want)) | (np.abs(got - want) > 1e-10 * scale))
                    for b in d[:4]:
                        s_, r_, p_ = b
                        print('  unit', tuple(b), 'got', got[tuple(b)], 'want', want[tuple(b)], 'set', sets[s_].tolist())
        finally:
            ctx.set_option("s2n_mode", 1)
    return bad


if __name__ == "__main__":
    n_case = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    n_bad = run(n_case, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print("cases", n_case, "mismatches", n_bad)
