"""Synthetic workloads of the named shapes (BASELINE.md
section 3): expression x [n x G] ~ N(mu_g, 1), mu_g ~ U(4, 12); two-class y;
MSigDB-sized gene sets with log-uniform sizes; 1 % of the sets spiked with a +1
shift in class a.  NumPy default_rng(seed) throughout; seeds are reported in
the bench JSON."""
import numpy as np

CONFIGS = {
    # name: G, n, S, (lo, hi), P
    "C1": dict(G=1000, n=20, S=50, sizes=(10, 50), P=1000, size_dist="uniform"),
    "C2": dict(G=20000, n=200, S=8000, sizes=(15, 500), P=10000, size_dist="loguniform"),
    "C3": dict(G=20000, n=500, S=8000, sizes=(15, 500), P=10000, size_dist="loguniform"),
    "C4": dict(G=20000, n=200, S=20000, sizes=(15, 300), P=50000, size_dist="loguniform"),
    "C5": dict(G=60000, n=1000, S=20000, sizes=(15, 500), P=100000, size_dist="loguniform"),
}


def gene_sets_csr(G, S, sizes, size_dist="loguniform", seed=1):
    rng = np.random.default_rng(seed)
    lo, hi = sizes
    if size_dist == "uniform":
        m = rng.integers(lo, hi + 1, size=S)
    else:
        m = np.floor(np.exp(rng.uniform(np.log(lo), np.log(hi + 1), size=S))).astype(np.int64)
        m = np.clip(m, lo, hi)
    m = np.minimum(m, G)
    offsets = np.zeros(S + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(m)
    index = np.empty(int(offsets[-1]), dtype=np.int32)
    for s in range(S):  # members without replacement, 1-based
        index[offsets[s]:offsets[s + 1]] = rng.choice(G, size=int(m[s]), replace=False) + 1
    return offsets, index


def two_class(G, n, offsets=None, index=None, seed=2, spike_frac=0.01):
    rng = np.random.default_rng(seed)
    mu = rng.uniform(4, 12, size=G)
    x = np.asfortranarray(rng.normal(size=(n, G)) + mu[None, :])
    y = np.zeros(n, dtype=np.int32)
    y[: n // 2] = 1  # class a == classes[1] -> TRUE
    rng.shuffle(y)
    if offsets is not None and spike_frac > 0:
        S = offsets.size - 1
        for s in rng.choice(S, size=max(1, int(S * spike_frac)), replace=False):
            genes = index[offsets[s]:offsets[s + 1]] - 1
            x[np.ix_(np.where(y == 1)[0], genes)] += 1.0
    return x, y.reshape(-1, 1)


def continuous_response(n, k=3, seed=3):
    rng = np.random.default_rng(seed)
    y = rng.normal(size=(n, 1 + k))
    return np.column_stack([np.ones(n), y])  # model matrix with intercept


def perm_matrix(n, P, seed=4):
    """[n x P] int32 1-based permutations (host-made; parity runs share it
    between the oracle and the device)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, P), dtype=np.int32, order="F")
    for p in range(P):
        out[:, p] = rng.permutation(n) + 1
    return out
