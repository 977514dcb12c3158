"""Null loop with few samples (6 = 3 + 3: 5 % of the permutations preserve the labels)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
G, n, S, P = 20000, 6, 8000, 4000
off, idx = synth.gene_sets_csr(G, S, (15, 500), "loguniform", seed=1)
x, y = synth.two_class(G, n, off, idx, seed=2)
ctx = fg.Context(0)
ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
perm = synth.perm_matrix(n, P, seed=3)
for with_obs in (False, True):
    if with_obs:
        sc = ctx.score_observed(nat.SCORE_S2N)
        obs = ctx.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=False)
    for _ in range(2):
        t = time.perf_counter()
        ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm, want_host=False)
        dt = time.perf_counter() - t
    print("observed ES known:", with_obs, "null loop %.1f ms" % (1e3 * dt), "launches", ctx.launch_count(), flush=True)
    if with_obs:
        sig = ctx.calc_sig(obs["es"][:, 0], 0, True, True, False)
        print("min p", sig["p"].min(), "p-count of set 0", sig["p_counts"][0])
