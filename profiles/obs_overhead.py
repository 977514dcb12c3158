"""Cost of the observed-ES hand-over in the null loop (C2): stage times without / with it."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat
import bench
c, x, y, off, idx = bench.workload("C2")
ctx = fg.Context(0); ctx.set_profiling(True)
ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y.astype("int32"))
for with_obs in (False, True):
    if with_obs:
        sc = ctx.score_observed(nat.SCORE_S2N); ctx.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=False)
    best = None
    for _ in range(3):
        ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, c["P"], seed=5, want_host=False)
        ms, ln = ctx.last_timings()
        if best is None or ms["total"] < best["total"]: best = ms
    print("observed known:", with_obs, {k: round(v, 2) for k, v in best.items()}, flush=True)
