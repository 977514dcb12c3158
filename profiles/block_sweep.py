"""Null-loop time at C2 as a function of the permutation block size (option perm_block)."""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat
import bench
c, x, y, off, idx = bench.workload("C2")
ctx = fg.Context(0)
ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y.astype("int32"))
for pb in [0, 1036, 2072, 3108, 4144, 5032, 10064]:
    ctx.set_option("perm_block", pb)
    best = None
    for _ in range(3):
        ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, c["P"], seed=5, want_host=False)
        ms, ln = ctx.last_timings()
        if best is None or ms["total"] < best["total"]: best = ms
    print(pb, {k: round(v, 2) for k, v in best.items()}, flush=True)
