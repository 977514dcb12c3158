import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
import bench
c, x, y, off, idx = bench.workload("C2")
ctx = fg.Context(0)
ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y.astype(np.int32))
sc = ctx.score_observed(nat.SCORE_S2N)
for rep in range(3):
    for rg in (True, False):
        t=time.perf_counter(); obs = ctx.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=rg); print("ragged", rg, "%.2f ms"%(1e3*(time.perf_counter()-t)))
