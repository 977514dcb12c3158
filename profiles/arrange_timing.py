#!/usr/bin/env python
"""C4 (maxmean / mean, 20,000 sets) null-loop ES stage with the bank-aware member arrangement
(arrange_idx16_kernel, option arrange_members) on and off; same nulls within 1e-12 expected."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth

c = synth.CONFIGS["C4"]
off, idx = synth.gene_sets_csr(c["G"], c["S"], c["sizes"], c["size_dist"], seed=1)
x, y = synth.two_class(c["G"], c["n"], off, idx, seed=2)
ctx = fg.Context(0)
ctx.set_profiling(True)
out = {}
ref = {}
for arr in (1, 0, 1):
    ctx.set_option("arrange_members", arr)
    ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
    for k, nm in ((nat.ES_MAXMEAN, "maxmean"), (nat.ES_MEAN, "mean")):
        best = None
        for _ in range(2):
            ctx.perm_null(nat.SCORE_S2N, k, 10000, seed=5, want_host=False)
            ms = ctx.last_timings()[0]
            if best is None or ms["es"] < best["es"]:
                best = ms
        null = ctx.perm_null(nat.SCORE_S2N, k, 64, seed=5)
        key = "%s arrange=%d" % (nm, arr)
        out[key] = {"es_ms_per_10000": round(best["es"], 2), "total": round(best["total"], 2)}
        if nm in ref:
            out[key]["max_abs_diff_vs_first"] = float(np.nanmax(np.abs(null - ref[nm])))
        else:
            ref[nm] = null
print(json.dumps(out))
