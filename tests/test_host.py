"""CPU tests of the host side: the C-ABI library loads and exports every symbol
include/flexgsea_b200.h declares (no compute without a GPU), and the host logic
that mirrors the reference's R API (GMT ingest, gene-set filtering, CSR packing,
response encoding, R's RNG stream, permutation sharding)."""
import os
import time
import re
import subprocess
import sys

import numpy as np
import pytest

import flexgsea_r_b200 as fg
from flexgsea_r_b200 import api, _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "flexgsea_b200.h")).read()
    declared = set(re.findall(r"\b(fgs_[a-z0-9_]+)\s*\(", hdr)) - {"fgs_ctx"}
    lib = fg._native.load()
    assert declared == set(fg._native.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fgs_abi_version() == 2


def test_no_cpu_fallback_without_device():
    if fg._native.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(fg.FlexgseaError):
        fg.Context(0)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "flexgsea-r_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_read_gmt_and_filter(tmp_path):
    p = tmp_path / "a.gmt"
    p.write_text("S1\thttp://x\tg1\tg2\tg3\nS2\tna\tg2\tzz\nS3\tna\tg9\n")
    gs = fg.read_gmt(str(p))
    assert list(gs) == ["S1", "S2", "S3"] and gs["S1"] == ["g1", "g2", "g3"] and gs["S2"] == ["g2", "zz"]
    names = ["g1", "g2", "g3", "g4"]
    f = fg.filter_gene_sets(gs, names, gs_size_min=1, gs_size_max=2, verbose=False)
    assert list(f) == ["S2"] and f["S2"] == ["g2"]        # R/flexgsea.R:941-952
    f = fg.filter_gene_sets(gs, names, gs_size_min=1, gs_size_max=300, verbose=False)
    off, idx = fg.gene_sets_to_csr(f, names)
    assert list(off) == [0, 3, 4] and list(idx) == [1, 2, 3, 2]
    # duplicated gene names: match() takes the first (:279)
    off, idx = fg.gene_sets_to_csr({"a": ["g2", "g2"]}, ["g1", "g2", "g2"])
    assert list(idx) == [2, 2]


def test_native_gmt_ingest_matches_host_path(tmp_path):
    """fgs_gmt_parse / fgs_gmt_read (GMT -> CSR, native) against read_gmt +
    filter_gene_sets + gene_sets_to_csr on a random collection: CRLF lines, an
    empty line, a set without genes, unknown genes, a gene repeated inside a set,
    a duplicated gene name in the data (first match wins), duplicated set names."""
    rng = np.random.default_rng(5)
    genes = ["G%d" % i for i in range(400)]
    names = genes[:300] + ["G7"] + genes[300:350]          # G7 twice: match() returns the first
    lines = []
    for k in range(120):
        m = int(rng.integers(0, 40))
        mem = [genes[i] for i in rng.integers(0, 400, size=m)]  # with repeats, some absent from `names`
        lines.append("\t".join(["SET%d" % (k % 100), "desc %d" % k] + mem))
    lines.insert(17, "")
    lines.append("LONELY\tno genes")
    text = "\r\n".join(lines) + "\n"
    for smin, smax in ((1, 300), (5, 20), (10, 300)):
        # host path (lists keep duplicated set names apart: parse by hand like read_gmt does per line)
        sets = []
        for line in text.splitlines():
            if not line:
                continue
            parts = line.split("\t", 2)
            sets.append((parts[0], parts[2].split("\t") if len(parts) > 2 else []))
        present = set(names)
        first = {}
        for i, g in enumerate(names):
            first.setdefault(g, i + 1)
        w_names, w_off, w_idx = [], [0], []
        for nm, gs in sets:
            kept = [g for g in gs if g in present]
            if smin <= len(kept) <= smax:
                w_names.append(nm); w_idx += [first[g] for g in kept]; w_off.append(len(w_idx))
        got_names, off, idx, stats = nat.gmt_to_csr(text, names, smin, smax)
        assert got_names == w_names and list(off) == w_off and list(idx) == w_idx
        assert stats["sets_before"] == len(sets)
        assert stats["genes_before"] == sum(len(gs) for _, gs in sets)
        assert stats["genes_after"] == sum(sum(g in present for g in gs) for _, gs in sets)
    p = tmp_path / "c.gmt"
    p.write_bytes(text.encode())
    a = nat.gmt_to_csr(str(p), names, 5, 20)
    b = nat.gmt_to_csr(text, names, 5, 20)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # and the same answer as the documented Python helpers when set names are unique
    uniq = "A\tx\tG1\tG2\tG999\nB\tx\tG3\n"
    gs = fg.filter_gene_sets(fg.read_gmt(uniq), names, 1, 300, verbose=False)
    o2, i2 = fg.gene_sets_to_csr(gs, names)
    n3, o3, i3, _ = nat.gmt_to_csr(uniq, names, 1, 300)
    assert n3 == list(gs) and np.array_equal(o2, o3) and np.array_equal(i2, i3)


def test_binarise_like_flexgsea_s2n():
    yb, names = api._binarise(np.array(["b", "a", "a", "b"]))
    assert list(yb[:, 0]) == [0, 1, 1, 0] and names is None          # classes[1] = 'a'
    yb, _ = api._binarise(np.array([[1, 1, 0, 0]]).T)
    assert list(yb[:, 0]) == [0, 0, 1, 1]                            # class 1 = sort(unique)[1] = 0
    yb, _ = api._binarise(np.array([True, False, True]))
    assert list(yb[:, 0]) == [1, 0, 1]
    yb, _ = api._binarise(np.array([1.0, np.nan, 2.0]))
    assert yb[1, 0] == api.NA_LOGICAL
    with pytest.raises(ValueError):
        api._binarise(np.array([1, 2, 3]))
    import pandas as pd
    fac = pd.Series(pd.Categorical(["lo", "hi", "hi", "lo"], categories=["hi", "lo"]))
    yb, _ = api._binarise(fac)
    assert list(yb[:, 0]) == [0, 1, 1, 0]                            # a factor: class 1 = levels(y)[1] = 'hi'
    fac2 = pd.Series(pd.Categorical(["lo", "hi", "hi", "lo"], categories=["lo", "hi"]))
    assert list(api._binarise(fac2)[0][:, 0]) == [1, 0, 0, 1]


def test_design_like_flexgsea_lm():
    d, names = api._design(np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 7.0]]))
    assert names == ["(Intercept)", "Response1", "Response2"] and d.shape == (3, 3) and (d[:, 0] == 1).all()
    d, names = api._design(np.array([1.0, 2.0, 4.0]))
    assert names == ["(Intercept)", "Response"]


def test_rrng_sample_kinds():
    r = fg.RRNG(1, "Rejection")
    assert sorted(r.sample_int(10)) == list(range(1, 11))
    # R >= 3.6: set.seed(1); sample.int(10) -> 9 4 7 1 2 5 3 10 6 8
    assert list(fg.RRNG(1, "Rejection").sample_int(10)) == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    # R 3.x (Rounding): set.seed(1); sample.int(10) -> 3 4 5 7 2 8 9 6 10 1
    assert list(fg.RRNG(1, "Rounding").sample_int(10)) == [3, 4, 5, 7, 2, 8, 9, 6, 10, 1]
    pm = fg.RRNG(3).perm_matrix(7, 5)
    assert pm.shape == (5, 7) and all(sorted(row) == list(range(1, 8)) for row in pm)


def test_shard_spans_cover_all_permutations():
    for world in (1, 2, 3, 8):
        spans = []
        for rank in range(world):
            s = api._Shard(False); s.rank, s.world = rank, world
            spans.append(s.span(1001))
        assert spans[0][0] == 0 and spans[-1][1] == 1001
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


_GLOO = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
import torch.distributed as dist
import flexgsea_r_b200 as fg
from flexgsea_r_b200 import api, _native as nat
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(sys.argv[2]), rank=int(sys.argv[1]), world_size=2)
sh = api._Shard(None)
assert (sh.rank, sh.world) == (int(sys.argv[1]), 2)
lo, hi = sh.span(101)
g = sh.all_gather(np.full((3, 4), float(sh.rank)))
assert g.shape == (2, 3, 4) and (g[0] == 0).all() and (g[1] == 1).all()   # rank order
r = sh.all_reduce_sum(np.arange(5, dtype=np.int64) * (sh.rank + 1))
assert list(r) == [0, 3, 6, 9, 12]
tot = sh.all_reduce_sum(np.array([hi - lo], dtype=np.int64))
assert tot[0] == 101
# the NCCL unique id of the library's communicator travels from rank 0 like this
b = sh.broadcast_bytes(bytes(range(128)) if sh.rank == 0 else b"", 128)
assert b == bytes(range(128))
seeds = sh.all_gather(np.array([sh.random_seed()], dtype=np.int64))
assert seeds[0, 0] == seeds[1, 0] and seeds[0, 0] >= 0                      # one seed for all ranks
assert nat.shard_span(101, 2, sh.rank) == (lo, hi)                          # library and host agree on the spans
dist.destroy_process_group()
print("ok")
'''


def test_sharding_collectives_gloo_world2(tmp_path):
    """N>1 host path on CPU: two gloo ranks, all-gather in rank order and
    int64 all-reduce as sharded_sig uses them."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO % ROOT)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_limma_style_client_on_the_host():
    """flexgsea_limma (R/flexgsea_limma.R:15-31) in the plugin position: drops the intercept column,
    takes |t| with abs, needs a model matrix, and shrinks towards the ordinary t statistic."""
    rng = np.random.default_rng(3)
    G, n = 400, 14
    E = rng.normal(size=(G, n))
    D = np.column_stack([np.ones(n), rng.normal(size=n)])
    mm = fg.ModelMatrix(D, ["(Intercept)", "dose"])
    t = fg.flexgsea_limma(fg.EList(E, genes=["g%d" % i for i in range(G)]), mm)
    assert list(t.columns) == ["dose"] and t.shape == (G, 1) and np.isfinite(t.to_numpy()).all()
    assert np.array_equal(fg.flexgsea_limma(E, mm, abs=True).to_numpy(), np.abs(t.to_numpy()))
    beta = np.linalg.lstsq(D, E.T, rcond=None)[0].T
    res = E - beta @ D.T
    tord = beta[:, 1] / np.sqrt(np.linalg.inv(D.T @ D)[1, 1] * (res ** 2).sum(1) / (n - 2))
    assert np.corrcoef(tord, t.to_numpy()[:, 0])[0, 1] > 0.95
    assert np.abs(t.to_numpy()[:, 0]).max() <= np.abs(tord).max() * 1.5   # moderated: extremes pulled in
    with pytest.raises(ValueError):
        fg.flexgsea_limma(E, D)
    assert api._call_score(lambda x, y, t_x=False, abs=False: t_x, None, None, True, False) is True
    assert api._call_score(lambda x, y, abs=False: "no t_x", None, None, True, False) == "no t_x"


def test_bench_clock_sampler_keeps_the_timed_window():
    """bench.py's nvidia-smi sampler runs from before the warm-up; only the lines that arrived inside
    the timed region count, and a region shorter than one sampling period reports the warm-up's
    samples and says so"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)

    class Proc:
        def terminate(self): pass
        def wait(self, timeout=None): pass
    s = bench.ClockSampler(0)
    s.proc = Proc()
    s.lines = [(time.perf_counter(), "1200, 1965, 300.0, Not Active, Not Active, Not Active, Active")]
    time.sleep(0.002)
    s.begin()
    r = s.stop()
    assert r["window"] == "warmup" and r["samples"] == 1 and r["reasons"] == ["sw_power_cap"]
    s.lines.append((time.perf_counter(), "1965, 1965, 900.0, Not Active, Not Active, Not Active, Not Active"))
    r = s.stop()
    assert r == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 1, "window": "timed"}
