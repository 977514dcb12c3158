#!/usr/bin/env python
"""ES-kernel experiments: times the C2 null loop with a variant build of the
library (FLEXGSEA_B200_LIB, see profiles/build_variant.sh) and prints the stage times plus a
checksum of the null, so variants can be compared with the default build: same null expected.
Also times flexgsea_calc_sig for a full and for a one-eighth null (fixed costs of the significance)."""
import hashlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth

c = synth.CONFIGS["C2"]
off, idx = synth.gene_sets_csr(c["G"], c["S"], c["sizes"], c["size_dist"], seed=1)
x, y = synth.two_class(c["G"], c["n"], off, idx, seed=2)
ctx = fg.Context(0)
ctx.set_profiling(True)
ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
best = None
for _ in range(4):
    ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, c["P"], seed=5, want_host=False)
    ms = ctx.last_timings()[0]
    if best is None or ms["es"] < best["es"]:
        best = ms
null = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, 600, seed=5)
es = ctx.observed_es(nat.ES_WEIGHTED_KS, ctx.score_observed(nat.SCORE_S2N), ragged=False)["es"][:, 0]
sig = {}
for P in (c["P"], c["P"] // 8):
    ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, seed=5, want_host=False)
    t = []
    for _ in range(5):
        ctx.calc_sig(es)
        t.append(ctx.last_sig_timings()["sig"])
    sig["P=%d" % P] = round(min(t), 4)
print(json.dumps({"lib": os.path.basename(nat.LIB_PATH), "ms": {k: round(v, 3) for k, v in best.items()},
                  "null_sha1_600perms": hashlib.sha1(np.ascontiguousarray(null).tobytes()).hexdigest()[:16],
                  "calc_sig_ms": sig}))
