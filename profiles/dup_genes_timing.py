#!/usr/bin/env python
"""a C2-shaped dataset in which 5,000 genes are exact duplicates of others.
Every permutation then holds 5,000 exactly tied score pairs, the near-tie list of the tensor-core
score path overflows in every block, and the block falls back to the bit-faithful kernel.  Prints
the null-loop time next to the clean C2 time."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth

c = synth.CONFIGS["C2"]
off, idx = synth.gene_sets_csr(c["G"], c["S"], c["sizes"], c["size_dist"], seed=1)
x, y = synth.two_class(c["G"], c["n"], off, idx, seed=2)
ctx = fg.Context(0)
ctx.set_profiling(True)
out = {}
for name, xx in (("clean", x), ("5000 duplicated genes", None), ("bit-faithful kernel only (s2n_mode=0)", x)):
    if xx is None:
        xx = x.copy(order="F"); xx[:, 15000:] = xx[:, :5000]
    ctx.set_option("s2n_mode", 0 if "s2n_mode" in name else 1)
    ctx.set_x(xx); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    best = None
    for _ in range(3):
        ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, c["P"], seed=5, want_host=False)
        ms = ctx.last_timings()[0]
        if best is None or ms["total"] < best["total"]:
            best = ms
    out[name] = {k: round(v, 2) for k, v in best.items()}
out["ratio duplicated / clean"] = round(out["5000 duplicated genes"]["total"] / out["clean"]["total"], 3)
print(json.dumps(out))
