"""Host-side mirror of the reference's R API for the hot path.

Same names, argument meaning and error behaviour as R/flexgsea.R (dots become
underscores): ``flexgsea()`` with ``gene_score_fn``, ``es_fn``, ``sig_fun`` and
``nperm``; the built-ins ``flexgsea_s2n``, ``flexgsea_lm``,
``flexgsea_weighted_ks``, ``flexgsea_maxmean``, ``flexgsea_mean``,
``flexgsea_calc_sig``, ``flexgsea_calc_sig_simple``; and ``read_gmt``.

What changed underneath: the per-permutation closure of
flexgsea_perm_sequential/_parallel (R/flexgsea.R:356-472), the ES functions
(:725-937) and flexgsea_calc_sig (:474-629) are one call each into the sm_100a
library (include/flexgsea_b200.h).  A user-defined ``gene_score_fn`` still runs
on the host per permutation (:393-394) and its [G, R, P] score array is shipped
once; a user-defined ``es_fn`` / ``sig_fun`` is arbitrary host code and keeps
the reference's plugin loop.  The built-ins never run on the CPU.
"""
import inspect
import os

import numpy as np

from . import _native as nat
from .rrng import RRNG

try:  # result tables are data frames when pandas is around (it is in this image)
    import pandas as _pd
except Exception:  # pragma: no cover
    _pd = None

NA_LOGICAL = np.iinfo(np.int32).min

_contexts = {}
_groups = {}


def get_context(device=None):
    """Process-wide context of a GPU (created on first use)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else 0
        if device >= max(nat.device_count(), 1):
            device = 0
    if device not in _contexts:
        _contexts[device] = nat.Context(device)
    return _contexts[device]


def get_group(n_gpu):
    """Process-wide group of the first `n_gpu` GPUs (flexgsea(parallel = N) from a
    single process: one context and one host thread per GPU, NCCL inside the library)."""
    if n_gpu not in _groups:
        _groups[n_gpu] = nat.Group(n_gpu)
    return _groups[n_gpu]


# --------------------------------------------------------------------------
# plugin records
# --------------------------------------------------------------------------
class ModelMatrix:
    """A design matrix with column names, the analogue of an R matrix carrying
    the `assign` attribute (is.model.matrix, R/flexgsea.R:631-633)."""

    def __init__(self, values, colnames):
        self.values = np.asarray(values, dtype=np.float64)
        self.colnames = list(colnames)
        assert self.values.ndim == 2 and self.values.shape[1] == len(self.colnames)


class EList:
    """limma's EList as far as flexgsea() looks at it (R/flexgsea.R:159-173, t.x = TRUE): ``E`` is
    [genes x samples] (log-expression, e.g. from voom), ``genes`` its row names, ``weights`` the
    optional [genes x samples] precision weights limma::lmFit uses."""

    def __init__(self, E, genes=None, weights=None):
        self.E = np.asarray(E, dtype=np.float64)
        assert self.E.ndim == 2
        self.genes = None if genes is None else [str(g) for g in genes]
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        assert self.weights is None or self.weights.shape == self.E.shape


def _y_matrix(y):
    """-> (2-D object array / float array, colnames or None)"""
    if isinstance(y, ModelMatrix):
        return y.values, y.colnames
    if _pd is not None and isinstance(y, _pd.DataFrame):
        return y.to_numpy(), [str(c) for c in y.columns]
    if _pd is not None and isinstance(y, _pd.Series):
        return y.to_numpy().reshape(-1, 1), None
    a = np.asarray(y)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return a, None


def _binarise(y):
    """flexgsea_s2n's y_bin (R/flexgsea.R:695-710) -> int32 [n x R], NA kept.  A logical y has
    classes c(TRUE, FALSE), a factor (pandas Categorical) its levels, anything else
    sort(unique()) per response; class 1 -> TRUE."""
    levels = None
    if _pd is not None and isinstance(y, _pd.Series) and isinstance(y.dtype, _pd.CategoricalDtype):
        levels = list(y.cat.categories)   # is.factor(y): classes = levels(y), :699-701
        y = y.astype(object)
    vals, names = _y_matrix(y)
    n, R = vals.shape
    out = np.zeros((n, R), dtype=np.int32, order="F")
    for r in range(R):
        col = vals[:, r]
        if col.dtype == np.bool_:
            out[:, r] = col.astype(np.int32)  # classes = c(TRUE, FALSE)
            continue
        isna = np.array([v is None or (isinstance(v, float) and np.isnan(v)) for v in col.tolist()])
        classes = levels if levels is not None else sorted(set(np.asarray(col)[~isna].tolist()))  # sort(unique(phenotype))
        if len(classes) != 2:
            raise ValueError("length(classes) == 2 is not TRUE")  # stopifnot, :708
        b = np.array([v == classes[0] for v in col.tolist()], dtype=np.int32)
        b[isna] = NA_LOGICAL
        out[:, r] = b
    return out, names


def _design(y):
    """flexgsea_lm's model matrix (R/flexgsea.R:648-663) -> (design, colnames)."""
    if isinstance(y, ModelMatrix):
        return y.values, y.colnames
    vals, names = _y_matrix(y)
    try:
        v = np.asarray(vals, dtype=np.float64)
    except (TypeError, ValueError):
        raise NotImplementedError("flexgsea_lm: factor/character columns need an explicit ModelMatrix")
    if names is None:
        names = ["Response"] if v.shape[1] == 1 else ["Response%d" % (i + 1) for i in range(v.shape[1])]
    return np.column_stack([np.ones(v.shape[0]), v]), ["(Intercept)"] + list(names)


class ScoreFn:
    """Built-in gene.score.fn (flexgsea_s2n R/flexgsea.R:689-723, flexgsea_lm
    :647-676).  Callable like the R function: fn(x, y, abs=False) -> [G x R]."""

    def __init__(self, name, kind):
        self.name, self.kind = name, kind

    def __repr__(self):
        return "<flexgsea_%s>" % self.name

    def upload_response(self, ctx, y):
        if self.kind == nat.SCORE_S2N:
            yb, names = _binarise(y)
            ctx.set_response_s2n(yb)
            return names
        d, names = _design(y)
        icpt = names[0] == "(Intercept)"
        ctx.set_response_lm(d, icpt)
        return names[1:] if icpt else names

    def __call__(self, x, y, abs=False, device=None):
        ctx = get_context(device)
        ctx.set_x(_x_values(x)[0])
        self.upload_response(ctx, y)
        return ctx.score_observed(self.kind, abs)


class EsFn:
    """Built-in es.fn record: list(run, prepare, extra_stats, extra)
    (R/flexgsea.R:749-753, :775-779, :925-937)."""

    def __init__(self, name, kind, extra_stats=(), extra=()):
        self.name, self.kind = name, kind
        self.extra_stats, self.extra = tuple(extra_stats), tuple(extra)

    def __repr__(self):
        return "<flexgsea_%s>" % self.name

    def prepare(self, gene_score):
        return {}

    def run(self, gene_score, gene_set, prep=None, return_stats=(), return_values=(), p=1.0,
            device=None):
        """es.fn$run contract (R/flexgsea.R:60-78): gene_score [G, R, P],
        gene_set 1-based indices -> dict(es=[R, P], ...)."""
        ctx = get_context(device)
        gs = np.asarray(gene_set, dtype=np.int32)
        sc = np.asarray(gene_score, dtype=np.float64)
        if sc.ndim == 2:
            sc = sc[:, :, None]
        ctx.set_gene_sets(np.array([0, gs.size], dtype=np.int32), gs)
        want_extra = bool(set(return_stats) & set(self.extra_stats)) or bool(set(return_values) & set(self.extra))
        res = {}
        if want_extra and self.kind == nat.ES_WEIGHTED_KS:
            G, R, P = sc.shape
            res["es"] = np.zeros((R, P))
            for k in list(return_stats):
                res[k] = np.full((R, P), np.nan)
            for k in return_values:
                if k in self.extra:
                    res[k] = [[None] * P for _ in range(R)]
            for j in range(P):
                o = _observed(ctx, self, sc[:, :, j], p)
                res["es"][:, j] = o["es"][0]
                for k in return_stats:
                    res[k][:, j] = o[k][0]
                for k in return_values:
                    if k in self.extra:
                        for i in range(R):
                            res[k][i][j] = o[k][i][0]
            return res
        es = ctx.es_from_scores(self.kind, sc, p)  # [1, R, P]
        res["es"] = es[0]
        return res


flexgsea_s2n = ScoreFn("s2n", nat.SCORE_S2N)
flexgsea_lm = ScoreFn("lm", nat.SCORE_LM)
flexgsea_weighted_ks = EsFn("weighted_ks", nat.ES_WEIGHTED_KS, extra_stats=("max_es_at", "le_prop"),
                            extra=("running_es_pos", "running_es_neg", "running_es_at", "leading_edge"))
flexgsea_maxmean = EsFn("maxmean", nat.ES_MAXMEAN)
flexgsea_mean = EsFn("mean", nat.ES_MEAN)


def _trigamma_inverse(x):
    """limma::trigammaInverse: Newton iterations on 1 / trigamma (x > 0)."""
    from scipy.special import polygamma
    x = float(x)
    if x > 1e7:
        return 1.0 / np.sqrt(x)
    if x < 1e-6:
        return 1.0 / x
    y = 0.5 + 1.0 / x
    for _ in range(50):
        tri = float(polygamma(1, y))
        dif = tri * (1.0 - tri / x) / float(polygamma(2, y))
        y += dif
        if -dif / y < 1e-8:
            break
    return y


def flexgsea_limma(x, y, abs=False):
    """The reference's limma adaptor (R/flexgsea_limma.R:15-31) as a client of the user-score
    path: moderated t of every non-intercept design column, ``fit <- eBayes(lmFit(x, y)); fit$t``.
    ``x`` is an EList (E [genes x samples], optional precision weights, as voom makes it) or a
    [genes x samples] matrix, ``y`` a ModelMatrix.  limma itself is third-party R code; this is
    a numpy restatement of lmFit (weighted least squares per gene) and eBayes' variance shrinkage
    (squeezeVar / fitFDist, Smyth 2004), host code in the plugin position like any user-defined
    gene.score.fn -- its scores are shipped to the device in tiles, ranking onward runs there.
    Numeric parity with limma is not pinned (limma is absent here)."""
    from scipy.special import digamma, polygamma
    if not isinstance(y, ModelMatrix):
        raise ValueError("is.model.matrix(y) is not TRUE")  # :16
    E = x.E if isinstance(x, EList) else np.asarray(x, dtype=np.float64)
    W = x.weights if isinstance(x, EList) else None
    D = y.values
    G, n = E.shape
    q = D.shape[1]
    if W is None:
        A = np.broadcast_to(D.T @ D, (G, q, q))
        b = E @ D
    else:
        A = np.einsum("ni,gn,nj->gij", D, W, D)
        b = np.einsum("ni,gn,gn->gi", D, W, E)
    Ainv = np.linalg.inv(A)
    beta = np.einsum("gij,gj->gi", Ainv, b)
    res = E - beta @ D.T
    rss = (res * res * (1.0 if W is None else W)).sum(axis=1)
    df = float(n - q)
    s2 = rss / df
    unscaled = np.sqrt(np.einsum("gii->gi", Ainv))
    # eBayes: s2 ~ s0^2 * F(df, df0); moment estimates on log s2 (limma::fitFDist)
    ok = s2 > 0
    z = np.log(s2[ok])
    e = z - digamma(df / 2.0) + np.log(df / 2.0)
    emean = e.mean()
    evar = ((e - emean) ** 2).sum() / max(e.size - 1, 1) - float(polygamma(1, df / 2.0))
    if evar > 0:
        df0 = 2.0 * _trigamma_inverse(evar)
        s02 = np.exp(emean + digamma(df0 / 2.0) - np.log(df0 / 2.0))
        s2_post = (df0 * s02 + df * s2) / (df0 + df)
    else:  # df0 = Inf: every gene gets the prior variance
        s2_post = np.full(G, np.exp(emean))
    t = beta / (unscaled * np.sqrt(s2_post)[:, None])
    names = list(y.colnames)
    if names[0] == "(Intercept)":  # :22-24
        t, names = t[:, 1:], names[1:]
    t = np.abs(t) if abs else t
    if _pd is not None:
        return _pd.DataFrame(t, columns=names)
    return t


def _call_score(fn, x, y, t_x, abs):
    """gene.score.fn(x, y, [t.x = t.x,] abs = abs): t.x is passed only to functions that take it
    (R/flexgsea.R:246-250, :393-394)."""
    try:
        takes = "t_x" in inspect.signature(fn).parameters
    except (TypeError, ValueError):
        takes = False
    return fn(x, y, t_x=t_x, abs=abs) if takes else fn(x, y, abs=abs)


def _table(cols):
    if _pd is not None:
        return _pd.DataFrame(cols)
    return cols


class SigFn:
    """flexgsea_calc_sig (R/flexgsea.R:564-629) / flexgsea_calc_sig_simple
    (:544-547).  Callable like the R function: fn(es, es_null, verbose, abs)
    -> data frame with columns es, p, nes, fdr, fwer (or es, p.high, [p.low,]
    p, fdr, fwer)."""

    def __init__(self, name, split_p, calc_nes):
        self.name, self.split_p, self.calc_nes = name, split_p, calc_nes

    def __repr__(self):
        return "<%s>" % self.name

    def columns(self, es, r, abs):
        cols = {"es": np.asarray(es, dtype=np.float64)}
        if self.split_p:
            cols["p"] = r["p"]
        else:
            cols["p.high"] = r["p_high"]
            if not abs:
                cols["p.low"] = r["p_low"]
            cols["p"] = r["p"]
        if self.calc_nes:
            cols["nes"] = r["nes"]
        cols["fdr"] = r["fdr"]
        cols["fwer"] = r["fwer"]
        return cols

    def __call__(self, es, es_null, verbose=False, abs=False, device=None):
        es = np.asarray(es, dtype=np.float64)
        en = np.asarray(es_null, dtype=np.float64)
        assert es.ndim == 1 and en.ndim == 2 and en.shape[0] == es.size  # stopifnot :566-571
        if en.shape[1] == 0:
            return _table({"es": es})  # :575-577
        ctx = get_context(device)
        ctx.set_null(en)
        r = ctx.calc_sig(es, 0, self.split_p, self.calc_nes, abs)
        return _table(self.columns(es, r, abs))


flexgsea_calc_sig = SigFn("flexgsea_calc_sig", True, True)
flexgsea_calc_sig_simple = SigFn("flexgsea_calc_sig_simple", False, False)


# --------------------------------------------------------------------------
# gene-set ingest (R/flexgsea.R:939-973)
# --------------------------------------------------------------------------
def read_gmt(file, progress=False):
    """GMT: name <tab> description <tab> gene <tab> gene ... per line
    (read_gmt, R/flexgsea.R:966-973).  Returns {name: [genes]} in file order."""
    if isinstance(file, (str, os.PathLike)) and os.path.exists(file):
        if str(file).endswith(".gz"):
            import gzip
            text = gzip.open(file, "rt").read()
        else:
            text = open(file, "r").read()
    else:
        text = str(file)  # literal data, like readr
    out = {}
    for line in text.splitlines():
        if not line:
            continue
        parts = line.split("\t", 2)
        out[parts[0]] = parts[2].split("\t") if len(parts) > 2 else []
    return out


def filter_gene_sets(gene_sets, gene_names, gs_size_min=10, gs_size_max=300, verbose=True):
    """filter_gene_sets (R/flexgsea.R:939-958): keep genes present in the data,
    then sets with gs_size_min <= size <= gs_size_max; order and names kept."""
    present = set(gene_names)
    out = {}
    n_before = n_after = 0
    for name, gs in gene_sets.items():
        kept = [g for g in gs if g in present]
        n_before += len(gs); n_after += len(kept)
        if gs_size_min <= len(kept) <= gs_size_max:
            out[name] = kept
    if verbose:
        print("Filtered %d of %d genes out" % (n_before - n_after, n_before))
        print("Filtered %d of %d gene sets out" % (len(gene_sets) - len(out), len(gene_sets)))
    return out


def gene_sets_to_csr(gene_sets, gene_names):
    """match(set, gene.names) for every set, packed once as CSR (replaces the
    per-set, per-block match() of R/flexgsea.R:279 and :412).  First match wins
    for duplicated gene names; duplicates inside a set are kept."""
    first = {}
    for i, g in enumerate(gene_names):
        first.setdefault(g, i + 1)
    offsets = np.zeros(len(gene_sets) + 1, dtype=np.int32)
    chunks = []
    for k, gs in enumerate(gene_sets.values()):
        idx = np.array([first[g] for g in gs], dtype=np.int32)
        chunks.append(idx)
        offsets[k + 1] = offsets[k] + idx.size
    index = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int32)
    return offsets, index


# --------------------------------------------------------------------------
def _x_values(x):
    if isinstance(x, EList):  # genes x samples; the device wants a gene's samples contiguous
        return np.asfortranarray(x.E.T), x.genes
    if _pd is not None and isinstance(x, _pd.DataFrame):
        return np.asfortranarray(x.to_numpy(dtype=np.float64)), [str(c) for c in x.columns]
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError("x should be a matrix or limma EList")  # :164
    return np.asfortranarray(a, dtype=np.float64), None


def _observed(ctx, es_fn, scores, p=1.0, ragged=True):
    """observed ES + extras for scores [G x R] -> dict of [S x R] arrays and
    ragged lists indexed [response][set]."""
    o = ctx.observed_es(es_fn.kind, scores, p, ragged=ragged)
    res = {"es": o["es"]}
    if es_fn.kind == nat.ES_WEIGHTED_KS and not ragged:
        res["max_es_at"] = o["max_es_at"]
        res["le_prop"] = o["le_prop"]
    elif es_fn.kind == nat.ES_WEIGHTED_KS:
        S, R = o["es"].shape
        off = ctx.offsets
        res["max_es_at"] = o["max_es_at"]
        res["le_prop"] = o["le_prop"]
        for k in ("running_es_pos", "running_es_neg", "running_es_at"):
            res[k] = [[o[k][off[s]:off[s + 1], r].copy() for s in range(S)] for r in range(R)]
        res["leading_edge"] = [[o["running_es_at"][off[s] + o["le_first"][s, r]:
                                                   off[s] + o["le_first"][s, r] + o["le_count"][s, r], r].copy()
                                for s in range(S)] for r in range(R)]
    return res


class _Shard:
    """Permutation sharding over the ranks of torch.distributed (one process
    per GPU); a single process is rank 0 of 1."""

    def __init__(self, parallel):
        self.rank, self.world, self.dist = 0, 1, None
        if parallel is False:
            return
        try:
            import torch.distributed as dist
        except Exception:  # pragma: no cover
            return
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.dist, self.rank, self.world = dist, dist.get_rank(), dist.get_world_size()

    def span(self, nperm):
        lo = (nperm * self.rank) // self.world
        hi = (nperm * (self.rank + 1)) // self.world
        return lo, hi

    def broadcast_bytes(self, b, n):
        """`n` bytes from rank 0 to every rank (the NCCL unique id of the library's communicator)"""
        if self.dist is None:
            return b
        import torch
        t = torch.zeros(n, dtype=torch.uint8)
        if self.rank == 0:
            t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
        if self.dist.get_backend() == "nccl":
            t = t.cuda()
        self.dist.broadcast(t, 0)
        return t.cpu().numpy().tobytes()

    def attach(self, ctx):
        """Give `ctx` the library's own NCCL communicator over these ranks (once), so that the
        significance exchange runs device to device inside the C ABI (fgs_calc_sig_sharded)."""
        if self.dist is None:
            return False
        if ctx.comm_info() == (self.world, self.rank):
            return True
        uid = nat.comm_unique_id() if self.rank == 0 else b""
        ctx.comm_init(self.broadcast_bytes(uid, nat.COMM_ID_BYTES), self.world, self.rank)
        return True

    def random_seed(self):
        """a fresh seed, the same on every rank (R draws from the session RNG when none is set)"""
        sd = np.frombuffer(os.urandom(8), dtype=np.int64) & np.int64(0x7fffffffffffffff)
        if self.dist is None:
            return int(sd[0])
        import torch
        t = self._tensor(sd.copy())
        self.dist.broadcast(t, 0)
        return int(t.cpu().numpy()[0])

    def _tensor(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a))
        if self.dist.get_backend() == "nccl":
            t = t.cuda()
        return t

    def all_gather(self, a):
        """[world, *a.shape] in rank order"""
        if self.dist is None:
            return a[None]
        import torch
        t = self._tensor(a)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def all_reduce_sum(self, a):
        if self.dist is None:
            return a
        t = self._tensor(a)
        self.dist.all_reduce(t)
        return t.cpu().numpy()


def _gather_null(ctx, shard, null_local, nperm_total):
    """the whole es.null [S, R, P] on every rank (return_values = 'es_null', or a user sig.fun):
    device to device through the library's communicator when it is attached
    (fgs_comm_allgather_null; spans may differ by one permutation), else through the host"""
    if shard.world == 1:
        return null_local
    lo, hi = shard.span(nperm_total)
    if ctx.comm_info() == (shard.world, shard.rank) and ctx.null_device_ptr()[3] == hi - lo:
        return ctx.allgather_null(nperm_total)     # (the resident null still is this rank's span)
    S, R, Pl = null_local.shape
    width = -(-nperm_total // shard.world)             # equal shapes for the host all-gather: pad to the widest span
    padded = np.zeros((S, R, width), order="F")
    padded[:, :, :Pl] = null_local
    parts = shard.all_gather(np.ascontiguousarray(padded))
    out = []
    for r in range(shard.world):
        lo, hi = (nperm_total * r) // shard.world, (nperm_total * (r + 1)) // shard.world
        out.append(parts[r][:, :, :hi - lo])
    return np.asfortranarray(np.concatenate(out, axis=2))


def sharded_sig(ctx, shard, sig_fun, es, response, abs, nperm_total):
    """flexgsea_calc_sig over a null whose permutations are sharded over ranks:
    all-gather of the per-set double-double null sums (combined in rank order),
    all-reduce of the integer counts, then the pooled-FDR exceedance counts
    all-reduced as int64 -- the exact equivalent of all-gathering the null NES
    (SURVEY.md section 8e).  With the library's communicator attached
    (_Shard.attach) the whole exchange is one C-ABI call, device to device; the
    staged form below goes through the host's torch.distributed instead."""
    if shard.dist is not None and ctx.comm_info() == (shard.world, shard.rank):
        return ctx.calc_sig_sharded(es, nperm_total, response, sig_fun.split_p, sig_fun.calc_nes, abs)
    sums, counts = ctx.sig_stage1(es, response, sig_fun.split_p, abs)
    parts = shard.all_gather(sums)
    total = shard.all_reduce_sum(counts)
    r = ctx.sig_stage2(es, parts, total, nperm_total, response, sig_fun.split_p, sig_fun.calc_nes, abs)
    if sig_fun.calc_nes:
        nc = shard.all_reduce_sum(r["null_counts"])
        gl = shard.all_reduce_sum(r["glob"])
        r.update(ctx.sig_finish(r["nes"], nc, gl, abs))
    return r


def flexgsea(x, y, gene_sets, gene_score_fn=flexgsea_s2n, es_fn=flexgsea_weighted_ks,
             sig_fun=flexgsea_calc_sig, gene_names=None, nperm=1000, gs_size_min=10,
             gs_size_max=300, verbose=True, block_size=100, parallel=None, abs=False,
             return_values=(), *, perm=None, seed=None, rng="philox", device=None):
    """Drop-in for flexgsea() (R/flexgsea.R:146-354).

    Extra keyword-only arguments (not in the reference): ``perm`` -- an
    [n_samples x nperm] 1-based permutation matrix, e.g. made by R itself, so
    the null is the one the reference would compute; ``seed`` + ``rng`` --
    ``rng='R'`` replays R's set.seed(seed); sample.int() stream on the host,
    ``rng='philox'`` draws on the device.  ``block_size`` is accepted for
    compatibility; the device path sizes its own blocks.  ``parallel`` never
    means forked workers: under torch.distributed (one process per GPU) the
    permutations are sharded over its ranks; in a single process ``parallel=N``
    (an integer > 1) shards them over N GPUs through the library's group of
    contexts, which is what the R binding does.
    """
    # ---- prepare and check input (:152-239) -----------------------------
    t_x = isinstance(x, EList)   # :159-173
    xv, xnames = _x_values(x)
    n_samples, n_genes = xv.shape
    yv, _ = _y_matrix(y)
    if yv.shape[0] != n_samples:
        raise ValueError("n.samples == nrow(y) is not TRUE")  # :174
    if not (np.isscalar(block_size) and block_size > 0):
        raise ValueError("block.size > 0 is not TRUE")  # :176-178
    return_values = [return_values] if isinstance(return_values, str) else list(return_values)
    gmt_file = None
    if isinstance(gene_sets, (str, os.PathLike)):
        if verbose:
            print("Reading gene sets from file")
        if str(gene_sets).endswith(".gz") or not os.path.exists(gene_sets):
            gene_sets = read_gmt(gene_sets)      # compressed / literal data: host reader
        else:
            gmt_file = gene_sets                 # plain file: native GMT -> CSR ingest below
    elif not isinstance(gene_sets, dict):
        raise TypeError("is.list(gene.sets) is not TRUE")
    if gene_names is not None:
        if len(gene_names) != n_genes:
            raise ValueError("ncol(x) == length(gene.names) is not TRUE")
        gene_names = [str(g) for g in gene_names]
    elif xnames is not None:
        gene_names = xnames
    else:
        raise ValueError("Gene names should be given in gene.name or as colnames of x")  # :225-226
    if verbose:
        print("Filtering gene sets on size in dataset")
    if gmt_file is not None:
        # read_gmt + filter_gene_sets + match(), once and natively (fgs_gmt_read)
        set_names, offsets, index, st = nat.gmt_to_csr(gmt_file, gene_names, gs_size_min, gs_size_max)
        if verbose:
            print("Filtered %d of %d genes out" % (st["genes_before"] - st["genes_after"], st["genes_before"]))
            print("Filtered %d of %d gene sets out" % (st["sets_before"] - len(set_names), st["sets_before"]))
        if len(set_names) == 0:
            raise ValueError("No valid gene sets after filtering for size.")  # :238
        gene_sets = {nm: [gene_names[i - 1] for i in index[offsets[k]:offsets[k + 1]]]
                     for k, nm in enumerate(set_names)}
    else:
        gene_sets = filter_gene_sets(gene_sets, gene_names, gs_size_min, gs_size_max, verbose)
        if len(gene_sets) == 0:
            raise ValueError("No valid gene sets after filtering for size.")  # :238
        set_names = list(gene_sets.keys())
        offsets, index = gene_sets_to_csr(gene_sets, gene_names)

    shard = _Shard(parallel)
    n_group = parallel if (isinstance(parallel, int) and not isinstance(parallel, bool)) else 0
    if shard.world == 1 and n_group > 1:
        ctx = get_group(n_group)     # same calls; every one fans out over the group's GPUs
    else:
        ctx = get_context(device)
        if shard.world > 1 and shard.dist.get_backend() == "nccl":
            shard.attach(ctx)
    grouped = isinstance(ctx, nat.Group)
    ctx.set_option("nperm_total", nperm)
    ctx.set_x(xv)
    ctx.set_gene_sets(offsets, index)
    builtin_score = isinstance(gene_score_fn, ScoreFn)
    builtin_es = isinstance(es_fn, EsFn)
    builtin_sig = isinstance(sig_fun, SigFn)

    # ---- observed (:241-293) ---------------------------------------------
    if verbose:
        print("Scoring Genes (Observed)")
    if builtin_score:
        responses = gene_score_fn.upload_response(ctx, y)
        try:
            scores = ctx.score_observed(gene_score_fn.kind, abs)
        except nat.NonFiniteScoreError:
            raise ValueError("Gene scoring function returned NA or Inf. Check that all genes in x "
                             "have a positive variance.")  # :252
    else:
        scores = _call_score(gene_score_fn, x, y, t_x, abs)
        responses = None
        if _pd is not None and isinstance(scores, _pd.DataFrame):
            responses = [str(c) for c in scores.columns]
        scores = np.asarray(scores, dtype=np.float64)
        if not np.isfinite(scores).all():
            raise ValueError("Gene scoring function returned NA or Inf. Check that all genes in x "
                             "have a positive variance.")
    if scores.ndim != 2 or scores.shape[0] != n_genes:
        raise ValueError("dim(gene.scores)[1] == n.genes is not TRUE")  # :260
    R = scores.shape[1]
    if responses is None:
        responses = ["Response %d" % (i + 1) for i in range(R)]  # :257
    if verbose:
        print("Calculating ES (Observed)")
    if builtin_es:
        obs = _observed(ctx, es_fn, scores, ragged=bool(set(es_fn.extra) & set(return_values)))
        es = obs["es"]
        extra_stats = {k: obs[k] for k in es_fn.extra_stats}
        extra = {k: obs[k] for k in es_fn.extra if k in return_values}
    else:
        es, extra_stats, extra = _plugin_observed(es_fn, scores, offsets, index, return_values)

    # ---- permutation test (:295-307) --------------------------------------
    if verbose:
        print("Performing Permutation test")
    lo, hi = shard.span(nperm)
    local = hi - lo
    if perm is not None:
        pm = np.asarray(perm, dtype=np.int32)
        if pm.shape != (n_samples, nperm):
            raise ValueError("perm must be [n_samples x nperm], 1-based")
        pm_local = np.asfortranarray(pm[:, lo:hi])
    else:
        if seed is None:   # every run its own permutations, the same on every rank
            seed = shard.random_seed()
        if rng == "R" or not (builtin_score and builtin_es):
            # the host loop below needs the permutations themselves
            gen = RRNG(seed % (2 ** 31))
            pm_local = np.asfortranarray(gen.perm_matrix(n_samples, nperm).T[:, lo:hi])
        else:
            pm_local = None
    want_null_host = ("es_null" in return_values) or not builtin_sig or not builtin_es
    if builtin_score and builtin_es:
        failed = None
        try:
            if grouped:
                null_local = ctx.perm_null(gene_score_fn.kind, es_fn.kind, nperm, perm=pm_local, seed=seed or 0,
                                           abs_flag=abs, want_host=want_null_host)
            else:
                null_local = ctx.perm_null(gene_score_fn.kind, es_fn.kind, local, perm=pm_local, seed=seed or 0,
                                           perm_id0=lo, abs_flag=abs, want_host=want_null_host)
        except nat.NonFiniteScoreError:
            failed = ValueError("Gene scoring function returned NA or Inf.")  # :401
        except nat.FlexgseaError as e:
            failed = e
        # a rank whose shard failed must not leave the others waiting in the exchange: everybody learns of it first
        if shard.world > 1 and int(shard.all_reduce_sum(np.array([0 if failed is None else 1], dtype=np.int64))[0]):
            raise failed if failed is not None else ValueError(
                "Gene scoring function returned NA or Inf. (on another rank's share of the permutations)")
        if failed is not None:
            raise failed
    elif builtin_es and not builtin_score:
        # A user-defined gene.score.fn (limma's moderated t, say) runs on the host per permutation
        # (:383-394); its scores go to the device in tiles of block.size permutations, and the
        # call that hands a tile over only queues the work, so tile k is uploaded, ranked and
        # scored while the host computes tile k + 1.
        tile = max(1, int(block_size))
        try:
            in_flight = None   # the tile the library's helper thread is still reading
            for t0 in range(0, local, tile):
                B = min(tile, local - t0)
                sc_tile = np.zeros((n_genes, R, B), order="F")
                for j in range(B):
                    yp = _permute_rows(y, pm_local[:, t0 + j] - 1)
                    sc_tile[:, :, j] = np.asarray(_call_score(gene_score_fn, x, yp, t_x, abs), dtype=np.float64)
                if not np.isfinite(sc_tile).all():
                    raise ValueError("Gene scoring function returned NA or Inf.")  # :400-402
                ctx.es_from_scores_tile(es_fn.kind, sc_tile, t0, local)
                in_flight = sc_tile
            null_local = ctx.null_finish(want_host=want_null_host) if local > 0 else np.zeros((len(set_names), R, 0), order="F")
        except nat.NonFiniteScoreError:
            raise ValueError("Gene scoring function returned NA or Inf.")  # :401
    else:
        sc_null = np.zeros((n_genes, R, local), order="F")
        for j in range(local):  # user closure per permutation (:383-394)
            yp = _permute_rows(y, pm_local[:, j] - 1)
            if builtin_score:
                sc_null[:, :, j] = gene_score_fn(x, yp, abs=abs, device=device)
            else:
                sc_null[:, :, j] = np.asarray(_call_score(gene_score_fn, x, yp, t_x, abs), dtype=np.float64)
        if not np.isfinite(sc_null).all():
            raise ValueError("Gene scoring function returned NA or Inf.")
        if builtin_es:
            if builtin_score:  # the calls above replaced the resident problem
                ctx.set_x(xv); ctx.set_gene_sets(offsets, index)
            null_local = ctx.es_from_scores(es_fn.kind, sc_null, want_host=want_null_host)
        else:
            null_local = _plugin_null(es_fn, sc_null, offsets, index)

    # ---- significance (:309-330) ------------------------------------------
    if verbose:
        print("Calculating Significance")
    sig = []
    for r in range(R):
        if builtin_sig and builtin_es:
            if shard.world == 1:   # (a group exchanges over NCCL inside this call)
                res = ctx.calc_sig(es[:, r], r, sig_fun.split_p, sig_fun.calc_nes, abs)
            else:
                res = sharded_sig(ctx, shard, sig_fun, es[:, r], r, abs, nperm)
            cols = sig_fun.columns(es[:, r], res, abs)
        else:
            full = _gather_null(ctx, shard, null_local, nperm)
            tab = sig_fun(es[:, r], np.ascontiguousarray(full[:, r, :]), verbose=verbose, abs=abs)
            cols = {k: np.asarray(tab[k]) for k in (tab.columns if hasattr(tab, "columns") else tab)}
        for k, v in extra_stats.items():  # :334-338
            cols[k] = v[:, r]
        cols["GeneSet"] = set_names
        sig.append(_table(cols))

    res = {"table": dict(zip(responses, sig))}
    if "es_null" in return_values:  # :342-344
        res["es_null"] = _gather_null(ctx, shard, null_local, nperm)
        res["es_null_dimnames"] = (set_names, list(responses))
    if "gene_names" in return_values:
        res["gene_names"] = gene_names
    for k, v in extra.items():  # :348-352
        if k not in res:
            res[k] = {resp: dict(zip(set_names, v[i])) for i, resp in enumerate(responses)}
    return res


def _permute_rows(y, idx0):
    if isinstance(y, ModelMatrix):
        return ModelMatrix(y.values[idx0], y.colnames)  # attributes kept (:385-392)
    if _pd is not None and isinstance(y, (_pd.DataFrame, _pd.Series)):
        return y.iloc[idx0].reset_index(drop=True)
    a = np.asarray(y)
    return a[idx0]


# ---- user-defined es.fn: the reference's plugin loop, on the host ----------
def _plugin_observed(es_fn, scores, offsets, index, return_values):
    S = offsets.size - 1
    R = scores.shape[1]
    gs3 = scores[:, :, None]
    prep = es_fn["prepare"](gs3)
    stats = list(es_fn.get("extra_stats", ()))
    es = np.full((S, R), np.nan)
    extra_stats = {k: np.full((S, R), np.nan) for k in stats}
    extra = {k: [[None] * S for _ in range(R)] for k in es_fn.get("extra", ()) if k in return_values}
    for s in range(S):
        r = es_fn["run"](gs3, index[offsets[s]:offsets[s + 1]], prep, return_values=return_values,
                         return_stats=stats)
        es[s] = np.asarray(r["es"])[:, 0]
        for k in stats:
            extra_stats[k][s] = np.asarray(r[k])[:, 0]
        for k in extra:
            for i in range(R):
                extra[k][i][s] = r[k][i][0]
    return es, extra_stats, extra


def _plugin_null(es_fn, sc_null, offsets, index):
    S = offsets.size - 1
    G, R, P = sc_null.shape
    out = np.zeros((S, R, P), order="F")
    prep = es_fn["prepare"](sc_null)
    for s in range(S):
        out[s] = np.asarray(es_fn["run"](sc_null, index[offsets[s]:offsets[s + 1]], prep)["es"])
    return out
