"""Emulation of base R's default RNG stream (Mersenne-Twister + Inversion).

flexgsea draws every sample permutation with ``sample.int(nrow(y))``
(reference R/flexgsea.R:384, :443).  The native path takes the permutation
matrix as an input, so a host that wants the *same* null distribution as an R
session started with ``set.seed(seed)`` needs R's stream.  This module restates
base R's ``set.seed`` scrambling, ``unif_rand`` (MT19937), ``norm_rand``
(Inversion, Wichura AS241) and ``sample.int`` without replacement for both
``sample.kind`` values ("Rounding", R < 3.6 / ``RNGversion('3.4')``, used by the
reference's tests, tests/testthat/test_weighted_ks.R:6-7; and "Rejection", the
default since R 3.6).  Pure host code; no device work.
"""
import math
import numpy as np

_I2_32M1 = 2.328306437080797e-10


class RRNG:
    def __init__(self, seed, sample_kind="Rejection"):
        assert sample_kind in ("Rounding", "Rejection")
        self.sample_kind = sample_kind
        s = np.uint32(seed & 0xFFFFFFFF)
        mul = np.uint32(69069)
        one = np.uint32(1)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = mul * s + one
            words = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = mul * s + one
                words[j] = s
        bg = np.random.MT19937()
        st = bg.state
        st["state"]["key"] = words[1:].copy()  # words[0] is the position slot (forced to 624)
        st["state"]["pos"] = 624
        bg.state = st
        self._bg = bg
        self._buf = np.empty(0, dtype=np.float64)
        self._pos = 0

    def _refill(self, n):
        raw = self._bg.random_raw(max(n, 4096)).astype(np.uint64)
        u = raw.astype(np.float64) * 2.3283064365386963e-10
        u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
        self._buf = np.concatenate([self._buf[self._pos:], u])
        self._pos = 0

    def unif(self, n=None):
        k = 1 if n is None else int(n)
        if self._pos + k > self._buf.size:
            self._refill(k)
        out = self._buf[self._pos:self._pos + k]
        self._pos += k
        return float(out[0]) if n is None else out.copy()

    # ---- rnorm, INVERSION ------------------------------------------------
    def rnorm(self, n):
        u = self.unif(2 * n)
        big = 134217728.0
        v = np.floor(big * u[0::2]) + u[1::2]
        return np.array([_qnorm(p / big) for p in v])

    # ---- sample.int ------------------------------------------------------
    def _unif_index(self, dn):
        if self.sample_kind == "Rounding":
            return int(math.floor(dn * self.unif()))
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            v = 0
            nn = 0
            while nn <= bits:
                v1 = int(math.floor(self.unif() * 65536))
                v = 65536 * v + v1
                nn += 16
            if bits < 64:
                v &= (1 << bits) - 1
            if v < dn:
                return v

    def sample_int(self, n, k=None):
        """sample.int(n, k) without replacement, 1-based."""
        k = n if k is None else k
        x = list(range(n))
        out = np.empty(k, dtype=np.int32)
        nn = n
        for i in range(k):
            j = self._unif_index(float(nn))
            out[i] = x[j] + 1
            nn -= 1
            x[j] = x[nn]
        return out

    def perm_matrix(self, n, nperm):
        """[n x nperm] int32 (column-major friendly: returned as (nperm, n)
        C-array == R's n x nperm column-major), one sample.int(n) per
        permutation in order, as flexgsea_perm_sequential draws them."""
        out = np.empty((nperm, n), dtype=np.int32)
        for p in range(nperm):
            out[p] = self.sample_int(n)
        return out


def _horner(r, cs):
    v = cs[0]
    for c in cs[1:]:
        v = v * r + c
    return v


_A = (2509.0809287301226727, 33430.575583588128105, 67265.770927008700853,
      45921.953931549871457, 13731.693765509461125, 1971.5909503065514427,
      133.14166789178437745, 3.387132872796366608)
_B = (5226.495278852545925, 28729.085735721942674, 39307.89580009271061,
      21213.794301586595867, 5394.1960214247511077, 687.1870074920579083,
      42.313330701600911252, 1.0)
_C = (7.7454501427834140764e-4, .0227238449892691845833, .24178072517745061177,
      1.27045825245236838258, 3.64784832476320460504, 5.7694972214606914055,
      4.6303378461565452959, 1.42343711074968357734)
_D = (1.05075007164441684324e-9, 5.475938084995344946e-4, .0151986665636164571966,
      .14810397642748007459, .68976733498510000455, 1.6763848301838038494,
      2.05319162663775882187, 1.0)
_E = (2.01033439929228813265e-7, 2.71155556874348757815e-5, .0012426609473880784386,
      .026532189526576123093, .29656057182850489123, 1.7848265399172913358,
      5.4637849111641143699, 6.6579046435011037772)
_F = (2.04426310338993978564e-15, 1.4215117583164458887e-7, 1.8463183175100546818e-5,
      7.868691311456132591e-4, .0148753612908506148525, .13692988092273580531,
      .59983224388401682147, 1.0)


def _qnorm(p):
    q = p - 0.5
    if abs(q) <= 0.425:
        r = 0.180625 - q * q
        return q * _horner(r, _A) / _horner(r, _B)
    r = p if q < 0 else 1.0 - p
    r = math.sqrt(-math.log(r))
    if r <= 5.0:
        r -= 1.6
        val = _horner(r, _C) / _horner(r, _D)
    else:
        r -= 5.0
        val = _horner(r, _E) / _horner(r, _F)
    return -val if q < 0 else val
