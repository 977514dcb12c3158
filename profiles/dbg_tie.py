import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
import pyoracle as oracle
G=11
rng = np.random.default_rng(G)
n, P = 12, 120
sets = [rng.choice(G, m, replace=False) + 1 for m in (1, 1, 1, 2, 2, 3, 4, 1, 2, 3, G - 1, G // 2)]
off, idx = oracle.csr(sets)
x = np.asfortranarray(rng.normal(size=(n, G)))
y = np.zeros((n, 1), dtype=np.int32); y[: n // 2] = 1
perm = synth.perm_matrix(n, P, seed=G + 1)
ctx = fg.Context(0); ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
got = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
gsc, grk = ctx.debug_last(P)
rc, want, sc = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS, return_scores=True)
d=np.abs(got-want); bad=np.argwhere(d>1e-10)
print(len(bad))
for b in bad[:4]:
    s_,_,p_ = b
    mem = idx[off[s_]:off[s_+1]]
    scol = sc[:,0,p_]; r = oracle.rank_desc(scol)
    print('set',s_,'perm',p_,'members',mem,'ranks',r[mem-1],'w',np.abs(scol[mem-1]).tolist(),'got',got[s_,0,p_],'want',want[s_,0,p_])
    print('  device scores equal oracle:', np.array_equal(gsc[:,p_], scol), np.abs(gsc[:,p_]-scol).max())
    w=np.abs(scol[mem-1]); o=np.argsort(r[mem-1]); W=np.sum(w.astype(np.longdouble)); W=np.float64(W)
    wn=w/W; cum=np.cumsum(wn[o].astype(np.longdouble)).astype(np.float64); k=np.arange(1,len(mem)+1)
    dm=(r[mem-1][o]-k)/np.float64(G-len(mem)); pos=cum-dm; neg=pos-wn[o]
    print('  numpy longdouble: pos',pos.tolist(),'neg',neg.tolist())
