"""ctypes front-end of the CPU oracle (oracle/oracle.c) and of the verbatim
compile of the reference's src/ggsea.cpp (oracle/_ref/libggsea_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline /
``--impl reference`` legs and __graft_entry__.smoke(); never by the product
package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

SCORE_S2N, SCORE_LM = 0, 1
ES_WKS, ES_MAXMEAN, ES_MEAN = 0, 1, 2
ERR_NONFINITE = 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_wks_column.restype = C.c_double
        _LIB.orc_maxmean_column.restype = C.c_double
        _LIB.orc_mean_column.restype = C.c_double
    return _LIB


def ref_lib():
    """The reference's own s2n_C, compiled verbatim (None when not built)."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libggsea_ref.so")
        if not os.path.exists(path):
            return None
        _REF = C.CDLL(path)
    return _REF


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def num_threads():
    return lib().orc_num_threads()


def use_all_cores():
    """OpenMP threads = the cores this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which would otherwise throttle the CPU baseline)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_threads(int(n))
    return num_threads()


# Arrays are passed the way R holds them: a Fortran-ordered [rows x cols]
# matrix is handed over as its flat column-major buffer.
def _colmajor(a, dtype):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def s2n_C(x, y, use_ref=False):
    """x [n x G], y [n x R] int logical (NA = INT_MIN) -> coef [G x R]."""
    x = _colmajor(x, np.float64); y = _colmajor(y, np.int32)
    n, G = x.shape; ny, R = y.shape
    out = np.zeros((G, R), dtype=np.float64, order="F")
    L = ref_lib() if use_ref else lib()
    fn = L.ref_s2n_C if use_ref else L.orc_s2n_C
    fn(_p(x, _dp), n, G, _p(y, _ip), ny, R, _p(out, _dp))
    return out


def s2n_post(coef, abs_flag=False):
    c = np.asfortranarray(np.array(coef, dtype=np.float64))
    lib().orc_s2n_post(_p(c, _dp), C.c_long(c.size), int(abs_flag))
    return c


def gene_sd(x):
    x = _colmajor(x, np.float64); n, G = x.shape
    out = np.zeros(G)
    lib().orc_gene_sd(_p(x, _dp), n, G, _p(out, _dp))
    return out


def lm_scores(x, design, has_intercept, abs_flag=False):
    x = _colmajor(x, np.float64); d = _colmajor(design, np.float64)
    n, G = x.shape; q = d.shape[1]
    R = q - (1 if has_intercept else 0)
    sd = gene_sd(x)
    out = np.zeros((G, R), order="F")
    rc = lib().orc_lm_scores(_p(x, _dp), n, G, _p(d, _dp), q, int(has_intercept), _p(sd, _dp),
                             int(abs_flag), _p(out, _dp))
    return rc, out


def rank_desc(s):
    s = _f64(s); G = s.size
    out = np.zeros(G, dtype=np.int32)
    lib().orc_rank_desc(_p(s, _dp), G, _p(out, _ip))
    return out


def wks_column(score, rank, gene_set, p=1.0, extras=False):
    score = _f64(score); rank = _i32(rank); gs = _i32(gene_set)
    G, m = score.size, gs.size
    if not extras:
        return lib().orc_wks_column(_p(score, _dp), _p(rank, _ip), G, _p(gs, _ip), m, C.c_double(p),
                                    None, None, None, None, None, None, None)
    mea = C.c_double(); lep = C.c_double(); lf = C.c_int(); lc = C.c_int()
    order = np.zeros(m, dtype=np.int32); rp = np.zeros(m); rn = np.zeros(m)
    es = lib().orc_wks_column(_p(score, _dp), _p(rank, _ip), G, _p(gs, _ip), m, C.c_double(p),
                              C.byref(mea), C.byref(lep), _p(order, _ip), _p(rp, _dp), _p(rn, _dp),
                              C.byref(lf), C.byref(lc))
    return dict(es=es, max_es_at=mea.value, le_prop=lep.value, running_es_at=order,
                running_es_pos=rp, running_es_neg=rn,
                leading_edge=order[lf.value:lf.value + lc.value].copy())


def csr(gene_sets_idx):
    """list of 1-based index arrays -> (offsets[S+1], index[nnz]) int32."""
    off = np.zeros(len(gene_sets_idx) + 1, dtype=np.int32)
    for i, g in enumerate(gene_sets_idx):
        off[i + 1] = off[i] + len(g)
    idx = np.concatenate([np.asarray(g, dtype=np.int32) for g in gene_sets_idx]) if len(gene_sets_idx) else np.zeros(0, np.int32)
    return off, _i32(idx)


def es_from_scores(es_kind, scores, offsets, index, p=1.0):
    """scores [G, R, P] -> es [S, R, P]."""
    sc = np.asfortranarray(np.asarray(scores, dtype=np.float64))
    G, R, P = sc.shape
    offsets = _i32(offsets); index = _i32(index); S = offsets.size - 1
    out = np.zeros((S, R, P), order="F")
    rc = lib().orc_es_from_scores(es_kind, C.c_double(p), _p(sc, _dp), G, R, P, _p(offsets, _ip),
                                  _p(index, _ip), S, _p(out, _dp))
    return rc, out


def philox_perm(seed, perm_id, n):
    out = np.zeros(n, dtype=np.int32)
    lib().orc_philox_perm(C.c_uint64(seed), C.c_uint64(perm_id), n, _p(out, _ip))
    return out


def perm_null(x, score_kind, resp, has_intercept, perm, offsets, index, es_kind, p=1.0,
              abs_flag=False, return_scores=False):
    """x [n x G]; resp: int y_bin [n x R] (S2N) or double design [n x q] (LM);
    perm [n x P] 1-based.  Returns (rc, es_null [S, R, P][, scores [G,R,P]])."""
    x = _colmajor(x, np.float64); n, G = x.shape
    if score_kind == SCORE_S2N:
        r = _colmajor(resp, np.int32); ncol = r.shape[1]; R = ncol
    else:
        r = _colmajor(resp, np.float64); ncol = r.shape[1]; R = ncol - (1 if has_intercept else 0)
    perm = _colmajor(perm, np.int32); P = perm.shape[1]
    assert perm.shape[0] == n
    offsets = _i32(offsets); index = _i32(index); S = offsets.size - 1
    out = np.zeros((S, R, P), order="F")
    sc = np.zeros((G, R, P), order="F") if return_scores else None
    rc = lib().orc_perm_null(_p(x, _dp), n, G, score_kind, r.ctypes.data_as(C.c_void_p), ncol,
                             int(has_intercept), _p(perm, _ip), P, _p(offsets, _ip), _p(index, _ip), S,
                             es_kind, C.c_double(p), int(abs_flag), _p(out, _dp), _p(sc, _dp))
    return (rc, out, sc) if return_scores else (rc, out)


def calc_sig(es, es_null, split_p=True, calc_nes=True, abs_flag=False, want_nes_null=False):
    """flexgsea_calc_sig (split_p=calc_nes=True) / flexgsea_calc_sig_simple
    (both False).  es [S], es_null [S x P].  Returns dict of columns + counts."""
    es = _f64(es); en = _colmajor(es_null, np.float64)
    S, P = en.shape
    res = {k: np.full(S, np.nan) for k in ("p", "p_high", "p_low", "nes", "fdr", "fwer")}
    pc = np.zeros((S, 2), dtype=np.int64); fc = np.zeros((S, 2), dtype=np.int64)
    gc = np.zeros(4, dtype=np.int64)
    nn = np.zeros((S, P), order="F") if want_nes_null else None
    lib().orc_calc_sig(_p(es, _dp), _p(en, _dp), S, C.c_long(P), int(split_p), int(calc_nes),
                       int(abs_flag), _p(res["p"], _dp), _p(res["p_high"], _dp), _p(res["p_low"], _dp),
                       _p(res["nes"], _dp), _p(res["fdr"], _dp), _p(res["fwer"], _dp), _p(pc, _lp),
                       _p(fc, _lp), _p(gc, _lp), _p(nn, _dp))
    res["p_counts"] = pc; res["fdr_counts"] = fc; res["glob_counts"] = gc
    if want_nes_null:
        res["nes_null"] = nn
    return res
