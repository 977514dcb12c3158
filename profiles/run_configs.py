#!/usr/bin/env python
"""Times the null loop of the other BASELINE.json configs on one B200 (not the
bench line; context for DESIGN.md): C1, C3 (lm + covariates), C4 (maxmean, mean),
and a reduced C5 shape.  Prints one JSON line per config."""
import json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth

ctx = fg.Context(0)
ctx.set_profiling(True)


def run(name, G, n, S, sizes, P, score, es_kind, reps=2):
    off, idx = synth.gene_sets_csr(G, S, sizes, "loguniform" if sizes[1] > 60 else "uniform", seed=1)
    x, y = synth.two_class(G, n, off, idx, seed=2)
    ctx.set_x(x); ctx.set_gene_sets(off, idx)
    if score == nat.SCORE_S2N:
        ctx.set_response_s2n(y); R = 1
    else:
        ctx.set_response_lm(synth.continuous_response(n, k=3, seed=3), True); R = 4
    best = None
    for _ in range(reps):
        ctx.perm_null(score, es_kind, P, seed=5, want_host=False)
        ms, ln = ctx.last_timings()
        if best is None or ms["total"] < best["total"]:
            best = ms
    print(json.dumps({"config": name, "G": G, "n": n, "S": S, "P": P, "R": R, "nnz": int(off[-1]),
                      "ms": {k: round(v, 2) for k, v in best.items()},
                      "ES_per_s": S * R * P / (best["total"] / 1e3)}), flush=True)


which = sys.argv[1:] or ["C1", "C3", "C4mm", "C4mean", "C5r"]
if "C1" in which: run("C1 s2n+wKS", 1000, 20, 50, (10, 50), 1000, nat.SCORE_S2N, nat.ES_WEIGHTED_KS)
if "C2" in which: run("C2 s2n+wKS", 20000, 200, 8000, (15, 500), 10000, nat.SCORE_S2N, nat.ES_WEIGHTED_KS)
if "C3" in which: run("C3 lm+wKS (R=4)", 20000, 500, 8000, (15, 500), 10000, nat.SCORE_LM, nat.ES_WEIGHTED_KS)
if "C4mm" in which: run("C4 s2n+maxmean", 20000, 200, 20000, (15, 300), 50000, nat.SCORE_S2N, nat.ES_MAXMEAN, reps=1)
if "C4mean" in which: run("C4 s2n+mean", 20000, 200, 20000, (15, 300), 50000, nat.SCORE_S2N, nat.ES_MEAN, reps=1)
if "C5r" in which: run("C5 shape, 1/8 of one GPU's share (P=1562)", 60000, 1000, 20000, (15, 500), 1562, nat.SCORE_S2N, nat.ES_WEIGHTED_KS, reps=2)
