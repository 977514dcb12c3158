"""ctypes binding of the C ABI in include/flexgsea_b200.h.

This is the Python stand-in for the Rcpp shim a maintainer of the reference
would add (r-package/src/flexgsea_native.cpp): every method is
one C-ABI call with host numpy buffers laid out the way R lays its arrays out
(column-major).  There is no CPU fallback: if the library is missing or no GPU
is visible, loading / creating a context raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (FLEXGSEA_B200_LIB: a differently built copy of the library, for the kernel experiments of profiles/)
LIB_PATH = os.environ.get("FLEXGSEA_B200_LIB") or os.path.join(_HERE, "libflexgsea_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NONFINITE, ERR_OOM, ERR_STATE, ERR_UNSUPPORTED, ERR_CANCELLED, ERR_COMM = range(9)
COMM_ID_BYTES = 128
SCORE_S2N, SCORE_LM = 0, 1
ES_WEIGHTED_KS, ES_MAXMEAN, ES_MEAN = 0, 1, 2

# every symbol include/flexgsea_b200.h declares (checked by the CPU test-suite)
SYMBOLS = [
    "fgs_abi_version", "fgs_last_error", "fgs_device_count", "fgs_create", "fgs_destroy",
    "fgs_s2n", "fgs_set_x", "fgs_set_gene_sets", "fgs_set_response_s2n", "fgs_set_response_lm",
    "fgs_n_response", "fgs_score_observed", "fgs_observed_es", "fgs_set_observed_es", "fgs_perm_null",
    "fgs_es_from_scores", "fgs_es_from_scores_tile", "fgs_null_finish", "fgs_debug_last_scores", "fgs_null_device_ptr", "fgs_set_null",
    "fgs_calc_sig", "fgs_sig_stage1", "fgs_sig_stage2", "fgs_sig_finish", "fgs_launch_count",
    "fgs_last_timings", "fgs_set_profiling", "fgs_set_option", "fgs_get_stat", "fgs_set_progress",
    "fgs_gmt_parse", "fgs_gmt_read", "fgs_gmt_n_sets", "fgs_gmt_nnz", "fgs_gmt_offsets", "fgs_gmt_index",
    "fgs_gmt_set_name", "fgs_gmt_stats", "fgs_gmt_free",
    "fgs_comm_unique_id", "fgs_comm_init", "fgs_comm_destroy", "fgs_comm_info", "fgs_shard_span",
    "fgs_calc_sig_sharded", "fgs_comm_allgather_null", "fgs_last_sig_timings",
    "fgs_group_create", "fgs_group_destroy", "fgs_group_size", "fgs_group_ctx", "fgs_group_set_x",
    "fgs_group_set_gene_sets", "fgs_group_set_response_s2n", "fgs_group_set_response_lm",
    "fgs_group_set_option", "fgs_group_score_observed", "fgs_group_observed_es", "fgs_group_perm_null",
    "fgs_group_es_from_scores", "fgs_group_es_from_scores_tile", "fgs_group_null_finish", "fgs_group_calc_sig", "fgs_group_last_timings",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)
_lib = None


class FlexgseaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class NonFiniteScoreError(FlexgseaError):
    """The reference's stop("Gene scoring function returned NA or Inf.")."""


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.fgs_last_error.restype = C.c_char_p
        _lib.fgs_group_ctx.restype = C.c_void_p
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f(a):
    """column-major float64 view/copy of a matrix (R layout)"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _i(a):
    return np.asfortranarray(np.asarray(a, dtype=np.int32))


def gmt_to_csr(gmt, gene_names, gs_size_min=10, gs_size_max=300):
    """read_gmt + filter_gene_sets + match(), natively and once (fgs_gmt_parse /
    fgs_gmt_read).  `gmt` is a path or literal GMT text.  Returns
    (set names, offsets[S+1], index[nnz] 1-based, stats dict)."""
    L = load()
    L.fgs_gmt_set_name.restype = C.c_char_p
    L.fgs_gmt_offsets.restype = _ip
    L.fgs_gmt_index.restype = _ip
    L.fgs_gmt_nnz.restype = C.c_longlong
    names = [str(g).encode() for g in gene_names]
    arr = (C.c_char_p * len(names))(*names)
    h = C.c_void_p()
    if isinstance(gmt, (str, os.PathLike)) and os.path.exists(gmt):
        rc = L.fgs_gmt_read(str(gmt).encode(), arr, len(names), int(gs_size_min), int(gs_size_max), C.byref(h))
    else:
        text = gmt if isinstance(gmt, bytes) else str(gmt).encode()
        rc = L.fgs_gmt_parse(text, C.c_size_t(len(text)), arr, len(names), int(gs_size_min),
                             int(gs_size_max), C.byref(h))
    if rc != OK:
        raise FlexgseaError(rc, L.fgs_last_error().decode())
    try:
        S, nnz = L.fgs_gmt_n_sets(h), L.fgs_gmt_nnz(h)
        offsets = np.ctypeslib.as_array(L.fgs_gmt_offsets(h), shape=(S + 1,)).copy()
        index = (np.ctypeslib.as_array(L.fgs_gmt_index(h), shape=(nnz,)).copy() if nnz > 0
                 else np.zeros(0, dtype=np.int32))
        set_names = [L.fgs_gmt_set_name(h, k).decode() for k in range(S)]
        gb, ga, sb = C.c_longlong(), C.c_longlong(), C.c_int()
        L.fgs_gmt_stats(h, C.byref(gb), C.byref(ga), C.byref(sb))
        stats = {"genes_before": gb.value, "genes_after": ga.value, "sets_before": sb.value}
    finally:
        L.fgs_gmt_free(h)
    return set_names, offsets.astype(np.int32), index.astype(np.int32), stats


def device_count():
    n = C.c_int(0)
    load().fgs_device_count(C.byref(n))
    return n.value


def comm_unique_id():
    """NCCL unique id (bytes) made by the library; rank 0 makes it, the host
    program carries it to the other ranks (fgs_comm_unique_id)."""
    L = load()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = L.fgs_comm_unique_id(buf)
    if rc != OK:
        raise FlexgseaError(rc, L.fgs_last_error().decode())
    return buf.raw


def shard_span(n_perm, n_ranks, rank):
    """[lo, hi) of the permutations rank `rank` of `n_ranks` takes (fgs_shard_span)."""
    lo, hi = C.c_longlong(), C.c_longlong()
    rc = load().fgs_shard_span(C.c_longlong(n_perm), int(n_ranks), int(rank), C.byref(lo), C.byref(hi))
    if rc != OK:
        raise FlexgseaError(rc, "fgs_shard_span: bad argument")
    return lo.value, hi.value


class Context:
    """One per GPU; owns the device-resident problem and null distribution."""

    def __init__(self, device=0, _handle=None):
        self._lib = load()
        self._owned = _handle is None
        if _handle is None:
            self._h = C.c_void_p()
            self._check(self._lib.fgs_create(int(device), C.byref(self._h)))
        else:  # a context that belongs to a Group
            self._h = C.c_void_p(_handle)
        self.device = device

    def _check(self, rc):
        if rc != OK:
            msg = self._lib.fgs_last_error().decode()
            if rc == ERR_NONFINITE:
                raise NonFiniteScoreError(rc, msg)
            raise FlexgseaError(rc, "flexgsea_b200 error %d: %s" % (rc, msg))

    def close(self):
        if self._h and self._owned:
            self._lib.fgs_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads ---------------------------------------------------------
    def set_x(self, x):
        x = _f(x)
        self.n, self.G = x.shape
        self._check(self._lib.fgs_set_x(self._h, _p(x, _dp), self.n, self.G))

    def set_gene_sets(self, offsets, index):
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        index = np.ascontiguousarray(index, dtype=np.int32)
        self.S = offsets.size - 1
        self.nnz = int(offsets[-1])
        self.offsets = offsets
        self._check(self._lib.fgs_set_gene_sets(self._h, _p(offsets, _ip), _p(index, _ip), self.S))

    def set_response_s2n(self, y_bin):
        y = _i(y_bin)
        if y.ndim == 1:
            y = y.reshape(-1, 1, order="F")
        self._check(self._lib.fgs_set_response_s2n(self._h, _p(y, _ip), y.shape[0], y.shape[1]))

    def set_response_lm(self, design, has_intercept):
        d = _f(design)
        self._check(self._lib.fgs_set_response_lm(self._h, _p(d, _dp), d.shape[0], d.shape[1],
                                                  int(bool(has_intercept))))

    def n_response(self, score_kind):
        r = C.c_int(0)
        self._check(self._lib.fgs_n_response(self._h, score_kind, C.byref(r)))
        return r.value

    # ---- a4 --------------------------------------------------------------
    def s2n(self, x, y):
        """s2n_C(x, y): x [n x G] double, y [n' x R] int logical -> [G x R]."""
        x = _f(x); y = _i(y)
        out = np.zeros((x.shape[1], y.shape[1]), order="F")
        self._check(self._lib.fgs_s2n(self._h, _p(x, _dp), x.shape[0], x.shape[1], _p(y, _ip),
                                      y.shape[0], y.shape[1], _p(out, _dp)))
        return out

    # ---- observed --------------------------------------------------------
    def score_observed(self, score_kind, abs_flag=False):
        R = self.n_response(score_kind)
        out = np.zeros((self.G, R), order="F")
        self._check(self._lib.fgs_score_observed(self._h, score_kind, int(abs_flag), _p(out, _dp)))
        return out

    def observed_es(self, es_kind, scores, p=1.0, ragged=True):
        """ragged=False skips the per-member arrays (running_es_*), which are
        only needed when the caller asked for them in return_values."""
        sc = _f(scores)
        G, R = sc.shape
        S, nnz = self.S, self.nnz
        es = np.zeros((S, R), order="F")
        res = {"es": es}
        if es_kind == ES_WEIGHTED_KS:
            mea = np.zeros((S, R), order="F"); lep = np.zeros((S, R), order="F")
            at = rp = rn = None
            if ragged:
                at = np.zeros((max(nnz, 1), R), dtype=np.int32, order="F")
                rp = np.zeros((max(nnz, 1), R), order="F"); rn = np.zeros((max(nnz, 1), R), order="F")
            lf = np.zeros((S, R), dtype=np.int32, order="F"); lc = np.zeros((S, R), dtype=np.int32, order="F")
            self._check(self._lib.fgs_observed_es(self._h, es_kind, C.c_double(p), _p(sc, _dp), G, R,
                                                  _p(es, _dp), _p(mea, _dp), _p(lep, _dp), _p(at, _ip),
                                                  _p(rp, _dp), _p(rn, _dp), _p(lf, _ip), _p(lc, _ip)))
            res.update(max_es_at=mea, le_prop=lep, running_es_at=at, running_es_pos=rp,
                       running_es_neg=rn, le_first=lf, le_count=lc)
        else:
            self._check(self._lib.fgs_observed_es(self._h, es_kind, C.c_double(p), _p(sc, _dp), G, R,
                                                  _p(es, _dp), None, None, None, None, None, None, None))
        return res

    def set_observed_es(self, es_kind, es):
        """observed ES [S x R] computed elsewhere: handed to the null loop like
        observed_es() does with its own result"""
        e = _f(es if np.ndim(es) == 2 else np.asarray(es).reshape(-1, 1))
        self._check(self._lib.fgs_set_observed_es(self._h, int(es_kind), _p(e, _dp), e.shape[0], e.shape[1]))

    # ---- null ------------------------------------------------------------
    def perm_null(self, score_kind, es_kind, n_perm, perm=None, seed=0, perm_id0=0, p=1.0,
                  abs_flag=False, want_host=True):
        """es_null [S, R, n_perm].  perm: [n x n_perm] int32 1-based (column
        per permutation) or None for device-drawn Philox permutations."""
        R = self.n_response(score_kind)
        pp = None
        if perm is not None:
            pm = _i(perm)
            assert pm.shape == (self.n, n_perm), (pm.shape, (self.n, n_perm))
            pp = _p(pm, _ip)
        out = np.zeros((self.S, R, n_perm), order="F") if want_host else None
        self._check(self._lib.fgs_perm_null(self._h, score_kind, es_kind, C.c_double(p), int(abs_flag),
                                            pp, C.c_ulonglong(seed), C.c_longlong(perm_id0), int(n_perm),
                                            _p(out, _dp)))
        return out

    def es_from_scores(self, es_kind, scores, p=1.0, want_host=True):
        sc = _f(scores)
        G, R, P = sc.shape
        out = np.zeros((self.S, R, P), order="F") if want_host else None
        self._check(self._lib.fgs_es_from_scores(self._h, es_kind, C.c_double(p), _p(sc, _dp), G, R, P,
                                                 _p(out, _dp)))
        return out

    def es_from_scores_tile(self, es_kind, scores, perm_offset, n_perm_total, p=1.0):
        """queue one tile [G, R, B] of user scores (permutations perm_offset ..); returns at once"""
        sc = _f(scores)
        G, R, B = sc.shape
        self._check(self._lib.fgs_es_from_scores_tile(self._h, es_kind, C.c_double(p), _p(sc, _dp), G, R, B,
                                                      C.c_longlong(perm_offset), C.c_longlong(n_perm_total)))
        self._tile_keep = sc   # the helper thread reads it until the next call on this context

    def null_finish(self, want_host=True):
        _, S, R, P = self.null_device_ptr()
        out = np.zeros((S, R, P), order="F") if want_host else None
        self._check(self._lib.fgs_null_finish(self._h, _p(out, _dp)))
        return out

    def debug_last(self, n_col, G=None):
        G = G or self.G
        sc = np.zeros((G, n_col), order="F"); rk = np.zeros((G, n_col), dtype=np.int32, order="F")
        self._check(self._lib.fgs_debug_last_scores(self._h, _p(sc, _dp), _p(rk, _ip), C.c_longlong(n_col)))
        return sc, rk

    def set_null(self, es_null):
        a = _f(es_null)
        if a.ndim == 2:
            a = a.reshape(a.shape[0], 1, a.shape[1], order="F")
        self._check(self._lib.fgs_set_null(self._h, _p(a, _dp), a.shape[0], a.shape[1], a.shape[2]))

    def null_device_ptr(self):
        ptr = C.c_void_p(); s = C.c_int(); r = C.c_int(); p = C.c_int()
        self._check(self._lib.fgs_null_device_ptr(self._h, C.byref(ptr), C.byref(s), C.byref(r), C.byref(p)))
        return ptr.value, s.value, r.value, p.value

    # ---- significance ----------------------------------------------------
    def calc_sig(self, es, response=0, split_p=True, calc_nes=True, abs_flag=False):
        es = np.ascontiguousarray(es, dtype=np.float64)
        S = es.size
        res = {k: np.full(S, np.nan) for k in ("p", "p_high", "p_low", "nes", "fdr", "fwer")}
        pc = np.zeros((S, 2), dtype=np.int64); fc = np.zeros((S, 2), dtype=np.int64)
        gc = np.zeros(4, dtype=np.int64)
        self._check(self._lib.fgs_calc_sig(self._h, _p(es, _dp), int(response), int(split_p), int(calc_nes),
                                           int(abs_flag), _p(res["p"], _dp), _p(res["p_high"], _dp),
                                           _p(res["p_low"], _dp), _p(res["nes"], _dp), _p(res["fdr"], _dp),
                                           _p(res["fwer"], _dp), _p(pc, _lp), _p(fc, _lp), _p(gc, _lp)))
        res.update(p_counts=pc, fdr_counts=fc, glob_counts=gc)
        return res

    def sig_stage1(self, es, response=0, split_p=True, abs_flag=False):
        es = np.ascontiguousarray(es, dtype=np.float64)
        S = es.size
        sums = np.zeros((S, 4)); counts = np.zeros((S, 6), dtype=np.int64)
        self._check(self._lib.fgs_sig_stage1(self._h, _p(es, _dp), int(response), int(split_p),
                                             int(abs_flag), _p(sums, _dp), _p(counts, _lp)))
        return sums, counts

    def sig_stage2(self, es, sums_parts, counts_total, n_perm_total, response=0, split_p=True,
                   calc_nes=True, abs_flag=False):
        es = np.ascontiguousarray(es, dtype=np.float64)
        S = es.size
        parts = np.ascontiguousarray(sums_parts, dtype=np.float64).reshape(-1, S, 4)
        cnt = np.ascontiguousarray(counts_total, dtype=np.int64)
        res = {k: np.full(S, np.nan) for k in ("p", "p_high", "p_low", "nes", "fwer", "fdr")}
        nc = np.zeros(S, dtype=np.int64); gl = np.zeros(2, dtype=np.int64)
        self._check(self._lib.fgs_sig_stage2(self._h, _p(es, _dp), int(response), int(split_p), int(calc_nes),
                                             int(abs_flag), _p(parts, _dp), parts.shape[0], _p(cnt, _lp),
                                             C.c_longlong(n_perm_total), _p(res["p"], _dp),
                                             _p(res["p_high"], _dp), _p(res["p_low"], _dp), _p(res["nes"], _dp),
                                             _p(res["fwer"], _dp), _p(res["fdr"], _dp), _p(nc, _lp),
                                             _p(gl, _lp)))
        res.update(null_counts=nc, glob=gl)
        return res

    def sig_finish(self, nes, null_counts, glob, abs_flag=False):
        nes = np.ascontiguousarray(nes, dtype=np.float64)
        S = nes.size
        nc = np.ascontiguousarray(null_counts, dtype=np.int64); gl = np.ascontiguousarray(glob, dtype=np.int64)
        fdr = np.zeros(S); fc = np.zeros((S, 2), dtype=np.int64); gc = np.zeros(4, dtype=np.int64)
        self._check(self._lib.fgs_sig_finish(self._h, _p(nes, _dp), _p(nc, _lp), _p(gl, _lp), S, int(abs_flag),
                                             _p(fdr, _dp), _p(fc, _lp), _p(gc, _lp)))
        return dict(fdr=fdr, fdr_counts=fc, glob_counts=gc)

    # ---- permutations sharded over GPUs (one process per GPU) ---------------
    def comm_init(self, unique_id, n_ranks, rank):
        assert len(unique_id) == COMM_ID_BYTES
        self._check(self._lib.fgs_comm_init(self._h, unique_id, int(n_ranks), int(rank)))

    def comm_destroy(self):
        self._check(self._lib.fgs_comm_destroy(self._h))

    def comm_info(self):
        n, r = C.c_int(), C.c_int()
        self._check(self._lib.fgs_comm_info(self._h, C.byref(n), C.byref(r)))
        return n.value, r.value

    def calc_sig_sharded(self, es, n_perm_total, response=0, split_p=True, calc_nes=True, abs_flag=False,
                         exchange=0):
        """flexgsea_calc_sig over the null sharded over the ranks of the attached
        communicator (NCCL inside the library); same result dict as calc_sig."""
        es = np.ascontiguousarray(es, dtype=np.float64)
        S = es.size
        res = {k: np.full(S, np.nan) for k in ("p", "p_high", "p_low", "nes", "fdr", "fwer")}
        pc = np.zeros((S, 2), dtype=np.int64); fc = np.zeros((S, 2), dtype=np.int64)
        gc = np.zeros(4, dtype=np.int64)
        self._check(self._lib.fgs_calc_sig_sharded(
            self._h, _p(es, _dp), int(response), int(split_p), int(calc_nes), int(abs_flag),
            C.c_longlong(n_perm_total), int(exchange), _p(res["p"], _dp), _p(res["p_high"], _dp),
            _p(res["p_low"], _dp), _p(res["nes"], _dp), _p(res["fdr"], _dp), _p(res["fwer"], _dp), _p(pc, _lp),
            _p(fc, _lp), _p(gc, _lp)))
        res.update(p_counts=pc, fdr_counts=fc, glob_counts=gc)
        return res

    def allgather_null(self, n_perm_total, want_host=True):
        """every rank's resident null becomes the full [S, R, n_perm_total]"""
        _, S, R, _ = self.null_device_ptr()
        out = np.zeros((S, R, n_perm_total), order="F") if want_host else None
        self._check(self._lib.fgs_comm_allgather_null(self._h, C.c_longlong(n_perm_total), _p(out, _dp)))
        return out

    def last_sig_timings(self):
        ms = (C.c_float * 4)()
        self._check(self._lib.fgs_last_sig_timings(self._h, ms))
        return {"sig": ms[0], "collectives": ms[1], "step": ms[2], "n_collectives": int(ms[3])}

    # ---- instrumentation -------------------------------------------------
    def launch_count(self):
        n = C.c_longlong(0)
        self._check(self._lib.fgs_launch_count(self._h, C.byref(n)))
        return n.value

    def set_profiling(self, on):
        self._check(self._lib.fgs_set_profiling(self._h, int(on)))

    def last_timings(self):
        ms = (C.c_float * 5)(); ln = (C.c_longlong * 5)()
        self._check(self._lib.fgs_last_timings(self._h, ms, ln))
        keys = ("perm", "score", "rank", "es", "total")
        return {k: ms[i] for i, k in enumerate(keys)}, {k: ln[i] for i, k in enumerate(keys)}

    def set_progress(self, fn):
        """fn(columns_done, columns_total) -> truthy to stop the running null loop
        (FlexgseaError with code 7); None removes the hook."""
        if fn is None:
            self._progress_cb = None
            self._check(self._lib.fgs_set_progress(self._h, None, None))
            return
        proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_longlong)
        self._progress_cb = proto(lambda _u, done, total: 1 if fn(int(done), int(total)) else 0)
        self._check(self._lib.fgs_set_progress(self._h, self._progress_cb, None))

    def set_option(self, name, value):
        self._check(self._lib.fgs_set_option(self._h, name.encode(), C.c_longlong(int(value))))

    def get_stat(self, name):
        v = C.c_longlong(0)
        self._check(self._lib.fgs_get_stat(self._h, name.encode(), C.byref(v)))
        return v.value


class Group:
    """One process, several GPUs: what flexgsea(parallel = N) binds from R
    (fgs_group_*).  Uploads are replicated, the null loop takes the whole
    permutation range and shards it, significance exchanges over NCCL."""

    def __init__(self, n_gpu, devices=None):
        self._lib = load()
        self._h = C.c_void_p()
        dv = None
        if devices is not None:
            dv = (C.c_int * n_gpu)(*[int(d) for d in devices])
        rc = self._lib.fgs_group_create(int(n_gpu), dv, C.byref(self._h))
        self._check(rc)
        self.size = self._lib.fgs_group_size(self._h)
        self.ctx = [Context(i, _handle=self._lib.fgs_group_ctx(self._h, i)) for i in range(self.size)]

    def _check(self, rc):
        if rc != OK:
            msg = self._lib.fgs_last_error().decode()
            if rc == ERR_NONFINITE:
                raise NonFiniteScoreError(rc, msg)
            raise FlexgseaError(rc, "flexgsea_b200 error %d: %s" % (rc, msg))

    def close(self):
        if self._h:
            for c in self.ctx:
                c._h = C.c_void_p()
            self._lib.fgs_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_x(self, x):
        x = _f(x)
        self.n, self.G = x.shape
        self._check(self._lib.fgs_group_set_x(self._h, _p(x, _dp), self.n, self.G))
        for c in self.ctx:
            c.n, c.G = self.n, self.G

    def set_gene_sets(self, offsets, index):
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        index = np.ascontiguousarray(index, dtype=np.int32)
        self.S, self.nnz, self.offsets = offsets.size - 1, int(offsets[-1]), offsets
        self._check(self._lib.fgs_group_set_gene_sets(self._h, _p(offsets, _ip), _p(index, _ip), self.S))
        for c in self.ctx:
            c.S, c.nnz, c.offsets = self.S, self.nnz, offsets

    def set_response_s2n(self, y_bin):
        y = _i(y_bin)
        if y.ndim == 1:
            y = y.reshape(-1, 1, order="F")
        self._check(self._lib.fgs_group_set_response_s2n(self._h, _p(y, _ip), y.shape[0], y.shape[1]))

    def set_response_lm(self, design, has_intercept):
        d = _f(design)
        self._check(self._lib.fgs_group_set_response_lm(self._h, _p(d, _dp), d.shape[0], d.shape[1],
                                                        int(bool(has_intercept))))

    def set_option(self, name, value):
        self._check(self._lib.fgs_group_set_option(self._h, name.encode(), C.c_longlong(int(value))))

    def n_response(self, score_kind):
        return self.ctx[0].n_response(score_kind)

    def score_observed(self, score_kind, abs_flag=False):
        R = self.n_response(score_kind)
        out = np.zeros((self.G, R), order="F")
        self._check(self._lib.fgs_group_score_observed(self._h, score_kind, int(abs_flag), _p(out, _dp)))
        return out

    def observed_es(self, es_kind, scores, p=1.0, ragged=True):
        sc = _f(scores)
        G, R = sc.shape
        S, nnz = self.S, self.nnz
        es = np.zeros((S, R), order="F")
        res = {"es": es}
        mea = lep = at = rp = rn = lf = lc = None
        if es_kind == ES_WEIGHTED_KS:
            mea = np.zeros((S, R), order="F"); lep = np.zeros((S, R), order="F")
            if ragged:
                at = np.zeros((max(nnz, 1), R), dtype=np.int32, order="F")
                rp = np.zeros((max(nnz, 1), R), order="F"); rn = np.zeros((max(nnz, 1), R), order="F")
            lf = np.zeros((S, R), dtype=np.int32, order="F"); lc = np.zeros((S, R), dtype=np.int32, order="F")
        self._check(self._lib.fgs_group_observed_es(self._h, es_kind, C.c_double(p), _p(sc, _dp), G, R,
                                                    _p(es, _dp), _p(mea, _dp), _p(lep, _dp), _p(at, _ip),
                                                    _p(rp, _dp), _p(rn, _dp), _p(lf, _ip), _p(lc, _ip)))
        if es_kind == ES_WEIGHTED_KS:
            res.update(max_es_at=mea, le_prop=lep, running_es_at=at, running_es_pos=rp, running_es_neg=rn,
                       le_first=lf, le_count=lc)
        return res

    def perm_null(self, score_kind, es_kind, n_perm, perm=None, seed=0, p=1.0, abs_flag=False,
                  want_host=True):
        R = self.n_response(score_kind)
        pp = None
        if perm is not None:
            pm = _i(perm)
            assert pm.shape == (self.n, n_perm), (pm.shape, (self.n, n_perm))
            pp = _p(pm, _ip)
        out = np.zeros((self.S, R, n_perm), order="F") if want_host else None
        self._check(self._lib.fgs_group_perm_null(self._h, score_kind, es_kind, C.c_double(p), int(abs_flag),
                                                  pp, C.c_ulonglong(seed), int(n_perm), _p(out, _dp)))
        return out

    def es_from_scores(self, es_kind, scores, p=1.0, want_host=True):
        sc = _f(scores)
        G, R, P = sc.shape
        out = np.zeros((self.S, R, P), order="F") if want_host else None
        self._check(self._lib.fgs_group_es_from_scores(self._h, es_kind, C.c_double(p), _p(sc, _dp), G, R, P,
                                                       _p(out, _dp)))
        return out

    def es_from_scores_tile(self, es_kind, scores, perm_offset, n_perm_total, p=1.0):
        sc = _f(scores)
        G, R, B = sc.shape
        self._tile_shape = (R, int(n_perm_total))
        self._check(self._lib.fgs_group_es_from_scores_tile(self._h, es_kind, C.c_double(p), _p(sc, _dp), G, R, B,
                                                            C.c_longlong(perm_offset), C.c_longlong(n_perm_total)))
        self._tile_keep = getattr(self, "_tile_keep2", None), sc   # tiles may straddle two GPUs' spans
        self._tile_keep2 = sc

    def null_finish(self, want_host=True):
        R, P = self._tile_shape
        out = np.zeros((self.S, R, P), order="F") if want_host else None
        self._check(self._lib.fgs_group_null_finish(self._h, _p(out, _dp)))
        return out

    def calc_sig(self, es, response=0, split_p=True, calc_nes=True, abs_flag=False, exchange=0):
        es = np.ascontiguousarray(es, dtype=np.float64)
        S = es.size
        res = {k: np.full(S, np.nan) for k in ("p", "p_high", "p_low", "nes", "fdr", "fwer")}
        pc = np.zeros((S, 2), dtype=np.int64); fc = np.zeros((S, 2), dtype=np.int64)
        gc = np.zeros(4, dtype=np.int64)
        self._check(self._lib.fgs_group_calc_sig(self._h, _p(es, _dp), int(response), int(split_p),
                                                 int(calc_nes), int(abs_flag), int(exchange), _p(res["p"], _dp),
                                                 _p(res["p_high"], _dp), _p(res["p_low"], _dp),
                                                 _p(res["nes"], _dp), _p(res["fdr"], _dp), _p(res["fwer"], _dp),
                                                 _p(pc, _lp), _p(fc, _lp), _p(gc, _lp)))
        res.update(p_counts=pc, fdr_counts=fc, glob_counts=gc)
        return res

    def last_timings(self):
        ms = (C.c_float * 4)()
        self._check(self._lib.fgs_group_last_timings(self._h, ms))
        return {"null": ms[0], "sig": ms[1], "collectives": ms[2], "step": ms[3]}
