#!/usr/bin/env python
"""Launches, once each on a realistic shape, the kernels round 1 left without counter evidence
(VERDICT "evidence gaps"): lm_gemm_kernel (C3 shape), es_mean_smem_kernel (C4 shape),
sort_rank_big_kernel + the global-weight ES instantiation (C5 shape), and the significance
kernels sig_partial_kernel / fdr_count_kernel over the full C2 null (8,000 x 10,000).  Meant to
be wrapped by ncu (-k regex:...); prints its own CUDA-event timings when run plain."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth

which = sys.argv[1:] or ["C3", "C4", "C5", "SIG"]
ctx = fg.Context(0)
ctx.set_profiling(True)
out = {}


def shape(name, seed=1):
    c = synth.CONFIGS[name]
    off, idx = synth.gene_sets_csr(c["G"], c["S"], c["sizes"], c["size_dist"], seed=seed)
    x, y = synth.two_class(c["G"], c["n"], off, idx, seed=seed + 1)
    return c, x, y, off, idx


if "C3" in which:
    c, x, y, off, idx = shape("C3")
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_lm(synth.continuous_response(c["n"], 3, 3), True)
    ctx.perm_null(nat.SCORE_LM, nat.ES_WEIGHTED_KS, 296, seed=5, want_host=False)
    out["C3 lm+wKS, 296 permutations x 4 responses"] = ctx.last_timings()[0]
if "C4" in which:
    c, x, y, off, idx = shape("C4")
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    for k, nm in ((nat.ES_MAXMEAN, "maxmean"), (nat.ES_MEAN, "mean")):
        ctx.perm_null(nat.SCORE_S2N, k, 1184, seed=5, want_host=False)
        out["C4 s2n+%s, 1184 permutations" % nm] = ctx.last_timings()[0]
if "C5" in which:
    c, x, y, off, idx = shape("C5")
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 296, seed=5, want_host=False)
    out["C5 s2n+wKS, 296 permutations"] = ctx.last_timings()[0]
if "SIG" in which:
    c, x, y, off, idx = shape("C2")
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    es = ctx.observed_es(nat.ES_WEIGHTED_KS, ctx.score_observed(nat.SCORE_S2N), ragged=False)["es"][:, 0]
    ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, c["P"], seed=5, want_host=False)
    out["C2 null loop, 10,000 permutations"] = ctx.last_timings()[0]
    for _ in range(2):
        ctx.calc_sig(es)
    out["C2 flexgsea_calc_sig over the 8,000 x 10,000 null (640 MB)"] = ctx.last_sig_timings()
print(json.dumps(out))
