"""User-score path (fgs_es_from_scores) at the C2 shape: wall time of the call from HOST score
columns (pageable numpy memory), against the device time of rank + ES alone.
Columns stream in blocks; the upload of block k+1 runs beside the ranking of block k."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat
import bench
c, x, y, off, idx = bench.workload("C2")
G, S = c["G"], c["S"]
P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ctx = fg.Context(0)
ctx.set_gene_sets(off, idx)
rng = np.random.default_rng(3)
sc = np.asfortranarray(rng.standard_normal((G, 1, P)))
print("scores on the host: %.2f GB" % (sc.nbytes / 1e9), flush=True)
for th in (0, 2, 4, 8):
  ctx.set_option("h2d_threads", th)
  print("host threads staging the upload: %d%s" % (th, " (plain copy from pageable memory)" if th == 0 else ""), flush=True)
  for rep in range(3):
    t0 = time.perf_counter()
    ctx.es_from_scores(nat.ES_WEIGHTED_KS, sc, want_host=False)
    wall = time.perf_counter() - t0
    ms, ln = ctx.last_timings()
    print("  rep %d: wall %.1f ms (%.2e ES/s, %.1f GB/s of host scores); device span %.1f ms"
            % (rep, wall * 1e3, S * P / wall, sc.nbytes / wall / 1e9, ms["total"]), flush=True)
# es.null back to the host (640 MB): the permutation loop with and without the host copy
ctx.set_x(x); ctx.set_response_s2n(y.astype("int32"))
out = np.zeros((S, 1, c["P"]), order="F")
for th in (0, 8):
    ctx.set_option("h2d_threads", th)
    for want in (False, True):
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, c["P"], seed=5, want_host=want)
            best = min(best, time.perf_counter() - t0)
        print("perm_null, staging threads %d, es.null to host %s: wall %.1f ms" % (th, want, best * 1e3), flush=True)

