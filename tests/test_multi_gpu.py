"""Permutations sharded over GPUs, on real NCCL: the result of
N = 2 must equal N = 1 bit for bit -- the null (Philox keyed by the global permutation id, or a
shared permutation matrix), p, NES, FDR and every count.  Needs two GPUs (skipped otherwise; run
with `gpurun --gpus 2`).  Replaces flexgsea_perm_parallel + sig.fun on the abind()-ed null
(R/flexgsea.R:422-472, :314-321)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need(n):
    if nat.device_count() < n:
        pytest.skip("needs %d GPUs" % n)


def _problem(seed=0, G=2500, n=30, S=400, P=203):
    off, idx = synth.gene_sets_csr(G, S, (10, 300), "loguniform", seed=seed + 1)
    x, y = synth.two_class(G, n, off, idx, seed=seed + 2)
    perm = synth.perm_matrix(n, P, seed=seed + 4)
    return x, y, perm, off, idx


SIG_KEYS = ("p", "nes", "fdr", "fwer", "p_counts", "fdr_counts", "glob_counts")


def _same_sig(a, b):
    for k in SIG_KEYS:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("es_kind", [nat.ES_WEIGHTED_KS, nat.ES_MAXMEAN])
@pytest.mark.parametrize("philox", [False, True])
def test_group_of_two_equals_single_gpu(es_kind, philox):
    _need(2)
    x, y, perm, off, idx = _problem()
    P = perm.shape[1]
    one = fg.Context(0)
    one.set_x(x); one.set_gene_sets(off, idx); one.set_response_s2n(y)
    sc = one.score_observed(nat.SCORE_S2N)
    obs = one.observed_es(es_kind, sc, ragged=False)
    null1 = one.perm_null(nat.SCORE_S2N, es_kind, P, perm=None if philox else perm, seed=77)
    sig1 = one.calc_sig(obs["es"][:, 0])
    simple1 = one.calc_sig(obs["es"][:, 0], 0, False, False)

    grp = nat.Group(2)
    try:
        assert grp.size == 2
        grp.set_x(x); grp.set_gene_sets(off, idx); grp.set_response_s2n(y)
        sc2 = grp.score_observed(nat.SCORE_S2N)
        assert np.array_equal(sc, sc2)
        obs2 = grp.observed_es(es_kind, sc2, ragged=False)
        assert np.array_equal(obs["es"], obs2["es"])
        null2 = grp.perm_null(nat.SCORE_S2N, es_kind, P, perm=None if philox else perm, seed=77)
        assert np.array_equal(null1, null2, equal_nan=True)
        # each GPU holds its span only
        assert [c.null_device_ptr()[3] for c in grp.ctx] == [P // 2, P - P // 2]
        sig2 = grp.calc_sig(obs2["es"][:, 0])
        _same_sig(sig1, sig2)
        t = grp.last_timings()
        assert t["collectives"] > 0 and t["sig"] >= t["collectives"]
        simple2 = grp.calc_sig(obs2["es"][:, 0], 0, False, False)
        for k in ("p", "p_high", "p_low", "fdr", "fwer", "p_counts"):
            assert np.array_equal(simple1[k], simple2[k], equal_nan=True), k
        # the exchange north_star names: all-gather of the null, every rank computes locally
        sig3 = grp.calc_sig(obs2["es"][:, 0], exchange=1)
        _same_sig(sig1, sig3)
        assert [c.null_device_ptr()[3] for c in grp.ctx] == [P, P]
    finally:
        grp.close()
        one.close()


def test_flexgsea_parallel_two_equals_one():
    """the R-API mirror: flexgsea(parallel = 2) in one process == flexgsea() on one GPU"""
    _need(2)
    G, n, S = 1500, 24, 120
    off, idx = synth.gene_sets_csr(G, S, (10, 120), "loguniform", seed=5)
    x, y = synth.two_class(G, n, off, idx, seed=6)
    names = ["g%d" % i for i in range(G)]
    sets = {"s%d" % k: [names[i - 1] for i in idx[off[k]:off[k + 1]]] for k in range(S)}
    kw = dict(gene_names=names, nperm=301, gs_size_min=1, gs_size_max=500, verbose=False, seed=9,
              return_values=["es_null"])
    a = fg.flexgsea(x, y[:, 0], sets, **kw)
    b = fg.flexgsea(x, y[:, 0], sets, parallel=2, **kw)
    ta, tb = a["table"]["Response 1"], b["table"]["Response 1"]
    for col in ("es", "p", "nes", "fdr", "fwer", "max_es_at", "le_prop"):
        assert np.array_equal(np.asarray(ta[col]), np.asarray(tb[col]), equal_nan=True), col
    assert np.array_equal(a["es_null"], b["es_null"], equal_nan=True)


_WORKER = r'''
import os, sys, time, numpy as np
sys.path.insert(0, %(root)r)
rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import _native as nat, synth
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from test_multi_gpu import _problem, SIG_KEYS
x, y, perm, off, idx = _problem(seed=3, P=517)
P = perm.shape[1]
ctx = fg.Context(rank)
idf = os.path.join(tmp, "nccl_id")
if rank == 0:
    open(idf + ".tmp", "wb").write(nat.comm_unique_id()); os.rename(idf + ".tmp", idf)
while not os.path.exists(idf):
    time.sleep(0.05)
ctx.comm_init(open(idf, "rb").read(), world, rank)
assert ctx.comm_info() == (world, rank)
ctx.set_option("nperm_total", P)
ctx.set_x(x); ctx.set_gene_sets(off, idx); ctx.set_response_s2n(y)
sc = ctx.score_observed(nat.SCORE_S2N)
es = ctx.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=False)["es"][:, 0]
lo, hi = nat.shard_span(P, world, rank)
mine = ctx.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, hi - lo, perm=perm[:, lo:hi], perm_id0=lo)
sig = ctx.calc_sig_sharded(es, P)
assert ctx.last_sig_timings()["n_collectives"] == 2
full = ctx.allgather_null(P)
assert np.array_equal(full[:, :, lo:hi], mine, equal_nan=True)
np.savez(os.path.join(tmp, "r%%d.npz" %% rank), full=full, **{k: sig[k] for k in SIG_KEYS})
ctx.comm_destroy()
print("ok", rank)
'''


def test_one_process_per_gpu_comm_equals_single_gpu(tmp_path):
    """fgs_comm_unique_id / fgs_comm_init (one process per GPU, the id carried through a file
    -- no torch involved), fgs_calc_sig_sharded and fgs_comm_allgather_null against one GPU."""
    _need(2)
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    x, y, perm, off, idx = _problem(seed=3, P=517)
    one = fg.Context(0)
    one.set_x(x); one.set_gene_sets(off, idx); one.set_response_s2n(y)
    es = one.observed_es(nat.ES_WEIGHTED_KS, one.score_observed(nat.SCORE_S2N), ragged=False)["es"][:, 0]
    null = one.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, perm.shape[1], perm=perm)
    sig = one.calc_sig(es)
    one.close()
    for r in range(2):
        got = np.load(os.path.join(tmp_path, "r%d.npz" % r))
        assert np.array_equal(got["full"], null, equal_nan=True)
        for k in SIG_KEYS:
            assert np.array_equal(got[k], sig[k], equal_nan=True), (r, k)


_DIST_WORKER = r'''
import os, sys, numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import synth
G, n, S = 1500, 24, 120
off, idx = synth.gene_sets_csr(G, S, (10, 120), "loguniform", seed=5)
x, y = synth.two_class(G, n, off, idx, seed=6)
names = ["g%%d" %% i for i in range(G)]
sets = {"s%%d" %% k: [names[i - 1] for i in idx[off[k]:off[k + 1]]] for k in range(S)}
res = fg.flexgsea(x, y[:, 0], sets, gene_names=names, nperm=301, gs_size_min=1, gs_size_max=500, verbose=False,
                  seed=9, return_values=["es_null"], device=rank)
assert fg.get_context(rank).comm_info() == (2, rank)      # the library's own communicator was attached
t = res["table"]["Response 1"]
np.savez(os.path.join(sys.argv[1], "d%%d.npz" %% rank), es_null=res["es_null"],
         **{c: np.asarray(t[c]) for c in ("es", "p", "nes", "fdr", "fwer")})
dist.destroy_process_group()
'''


def test_flexgsea_under_torch_distributed_equals_one_gpu(tmp_path):
    """one process per GPU (the torchrun layout): flexgsea() shards the permutations over the ranks of
    torch.distributed, attaches the library's NCCL communicator (id broadcast through torch) and
    returns the single-GPU result on every rank."""
    _need(2)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "d.py"
    script.write_text(_DIST_WORKER % {"root": ROOT})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(tmp_path)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    G, n, S = 1500, 24, 120
    off, idx = synth.gene_sets_csr(G, S, (10, 120), "loguniform", seed=5)
    x, y = synth.two_class(G, n, off, idx, seed=6)
    names = ["g%d" % i for i in range(G)]
    sets = {"s%d" % k: [names[i - 1] for i in idx[off[k]:off[k + 1]]] for k in range(S)}
    one = fg.flexgsea(x, y[:, 0], sets, gene_names=names, nperm=301, gs_size_min=1, gs_size_max=500, verbose=False,
                      seed=9, return_values=["es_null"])
    t = one["table"]["Response 1"]
    for r in range(2):
        got = np.load(os.path.join(tmp_path, "d%d.npz" % r))
        assert np.array_equal(got["es_null"], one["es_null"], equal_nan=True)
        for c in ("es", "p", "nes", "fdr", "fwer"):
            assert np.array_equal(got[c], np.asarray(t[c]), equal_nan=True), (r, c)


def test_few_samples_dedup_and_user_score_tiles_over_two_gpus(oracle):
    """(1) six samples: every GPU deduplicates the labellings of ITS share, routed by the total
    permutation count -- same tables as one GPU, counts identical to the oracle; (2) tiles of a
    user gene.score.fn go to the GPU whose span they fall into (a tile may straddle two)."""
    _need(2)
    G, n, S, P = 3000, 6, 200, 400
    off, idx = synth.gene_sets_csr(G, S, (10, 200), "loguniform", seed=31)
    x, y = synth.two_class(G, n, off, idx, seed=32)
    perm = synth.perm_matrix(n, P, seed=34)
    grp = nat.Group(2)
    one = fg.Context(0)
    try:
        for eng in (one, grp):
            eng.set_x(x); eng.set_gene_sets(off, idx); eng.set_response_s2n(y)
        res = []
        for eng in (one, grp):
            sc = eng.score_observed(nat.SCORE_S2N)
            es = eng.observed_es(nat.ES_WEIGHTED_KS, sc, ragged=False)["es"][:, 0]
            null = eng.perm_null(nat.SCORE_S2N, nat.ES_WEIGHTED_KS, P, perm=perm)
            res.append((null, eng.calc_sig(es)))
        assert one.get_stat("dedup_labellings") == 20 and grp.ctx[1].get_stat("dedup_labellings") == 20
        assert np.array_equal(res[0][0], res[1][0], equal_nan=True)
        _same_sig(res[0][1], res[1][1])
        rc, want = oracle.perm_null(x, oracle.SCORE_S2N, y, False, perm, off, idx, oracle.ES_WKS)
        osig = oracle.calc_sig(es, want[:, 0, :])
        assert rc == 0 and np.array_equal(res[1][1]["p_counts"], osig["p_counts"])
        assert np.array_equal(res[1][1]["fdr_counts"], osig["fdr_counts"])
        # user-score tiles over the group: 90 permutations in tiles of 25 (the 2nd tile straddles the spans 45 | 45)
        rng = np.random.default_rng(35)
        scu = rng.normal(size=(G, 2, 90))
        ref = one.es_from_scores(nat.ES_WEIGHTED_KS, scu)
        for t0 in range(0, 90, 25):
            grp.es_from_scores_tile(nat.ES_WEIGHTED_KS, np.asfortranarray(scu[:, :, t0:t0 + 25]), t0, 90)
        got = grp.null_finish()
        assert np.array_equal(got, ref, equal_nan=True)
        assert [c.null_device_ptr()[3] for c in grp.ctx] == [45, 45]
        es2 = rng.normal(size=S) * 0.3
        a = grp.calc_sig(es2, 1)
        one.set_null(ref[:, 1, :])
        b = one.calc_sig(es2, 0)
        for k in ("p_counts", "fdr_counts", "glob_counts", "nes", "fdr"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    finally:
        grp.close(); one.close()
