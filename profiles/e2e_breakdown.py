import sys, time, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
import bench
c, x, y, _icpt, off, idx = bench.workload("C2")
ctx = fg.Context(0)
perm = synth.perm_matrix(c["n"], c["P"], seed=4)
xp,_1 = bench.pinned(x); yp,_2 = bench.pinned(y.astype(np.int32)); pp,_3 = bench.pinned(perm); offp,_4=bench.pinned(off); idxp,_5=bench.pinned(idx)
def T(name, f):
    t=time.perf_counter(); r=f(); print("%-18s %8.2f ms"%(name, 1e3*(time.perf_counter()-t))); return r
for rep in range(2):
    print("--- rep", rep)
    T("set_x", lambda: ctx.set_x(xp)); T("set_gene_sets", lambda: ctx.set_gene_sets(offp, idxp)); T("set_response", lambda: ctx.set_response_s2n(yp))
    sc = T("score_observed", lambda: ctx.score_observed(nat.SCORE_S2N))
    obs = T("observed_es", lambda: ctx.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=False))
    T("perm_null", lambda: ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, c["P"], perm=pp, want_host=False))
    T("calc_sig", lambda: ctx.calc_sig(obs["es"][:,0], 0, True, True, False))
