import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
import pyoracle as oracle
G,n,S,P=600,20,40,32
off, idx = synth.gene_sets_csr(G, S, (5,90), "uniform", seed=1)
x, y = synth.two_class(G, n, off, idx, seed=2)
perm = synth.perm_matrix(n, P, seed=4)
ctx = fg.Context(0); ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
for p in (0.0, 0.5):
    got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm, p=p)
    rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS, p=p)
    d = np.abs(got-want); print(p, rc, d.max(), np.isnan(got).sum(), np.isnan(want).sum())
    i = np.unravel_index(np.nanargmax(d), d.shape); print(i, got[i], want[i], np.diff(off)[i[0]])
    ctx.set_option("force_generic_es", 1); gen = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm, p=p); ctx.set_option("force_generic_es", 0)
    print('generic vs oracle', np.abs(gen-want).max())
p=0.0
got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm, p=p)
rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS, p=p)
d=np.abs(got-want); bad=np.argwhere(d>1e-10)
print(len(bad))
for b in bad[:5]: print(tuple(b), got[tuple(b)], want[tuple(b)], np.diff(off)[b[0]])
