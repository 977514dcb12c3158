"""Plays R for the .Call entry points of r-package/src/flexgsea_native.cpp: the shim is compiled
against the functional stand-in of the R / Rcpp API in r-package/standin/ and its registered
routines are called by name with SEXP arguments, arity checked like R's .Call does.  Test
infrastructure only (R is absent from this image)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STANDIN = os.path.join(ROOT, "r-package", "standin")
NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP = 0, 10, 13, 14, 16, 19, 22


class RError(RuntimeError):
    pass


class Shim:
    def __init__(self):
        subprocess.check_call(["make", "-C", STANDIN, "-s"], stdout=subprocess.DEVNULL)
        L = self.L = C.CDLL(os.path.join(STANDIN, "libflexgsea_rshim.so"))
        for f in ("standin_call", "standin_nil", "standin_real", "standin_int", "standin_strings", "standin_elt"):
            getattr(L, f).restype = C.c_void_p
        for f in ("standin_condition", "standin_string", "standin_name", "standin_routine_name"):
            getattr(L, f).restype = C.c_char_p
        L.standin_real_ptr.restype = C.POINTER(C.c_double)
        L.standin_int_ptr.restype = C.POINTER(C.c_int)
        L.standin_length.restype = C.c_long
        L.standin_load()

    def routines(self):
        out = {}
        for k in range(self.L.standin_n_routines()):
            nm = self.L.standin_routine_name(k).decode()
            out[nm] = self.L.standin_arity(nm.encode())
        return out

    # ---- R values -------------------------------------------------------------------------
    def to_sexp(self, v):
        L = self.L
        if v is None:
            return C.c_void_p(L.standin_nil())
        if isinstance(v, C.c_void_p):
            return v
        if isinstance(v, (bool, np.bool_)):
            a = np.array([int(v)], dtype=np.int32)
            return C.c_void_p(L.standin_int(a.ctypes.data_as(C.POINTER(C.c_int)), 1, (C.c_int * 1)(1), 1))
        if isinstance(v, (int, np.integer)):
            a = np.array([int(v)], dtype=np.int32)
            return C.c_void_p(L.standin_int(a.ctypes.data_as(C.POINTER(C.c_int)), 1, (C.c_int * 1)(1), 0))
        if isinstance(v, float):
            a = np.array([v])
            return C.c_void_p(L.standin_real(a.ctypes.data_as(C.POINTER(C.c_double)), 1, (C.c_int * 1)(1)))
        if isinstance(v, str):
            v = [v]
        if isinstance(v, (list, tuple)) and all(isinstance(s, str) for s in v):
            arr = (C.c_char_p * len(v))(*[s.encode() for s in v])
            return C.c_void_p(L.standin_strings(len(v), arr))
        a = np.asarray(v)
        d = (C.c_int * a.ndim)(*a.shape)
        if a.dtype == np.bool_:
            f = np.asfortranarray(a.astype(np.int32))
            return C.c_void_p(L.standin_int(f.ctypes.data_as(C.POINTER(C.c_int)), a.ndim, d, 1))
        if np.issubdtype(a.dtype, np.integer):
            f = np.asfortranarray(a.astype(np.int32))
            return C.c_void_p(L.standin_int(f.ctypes.data_as(C.POINTER(C.c_int)), a.ndim, d, 0))
        f = np.asfortranarray(a.astype(np.float64))
        return C.c_void_p(L.standin_real(f.ctypes.data_as(C.POINTER(C.c_double)), a.ndim, d))

    def logical(self, a):
        """an R logical matrix (NA_LOGICAL = INT_MIN kept)"""
        a = np.asfortranarray(np.asarray(a, dtype=np.int32))
        d = (C.c_int * a.ndim)(*a.shape)
        return C.c_void_p(self.L.standin_int(a.ctypes.data_as(C.POINTER(C.c_int)), a.ndim, d, 1))

    def from_sexp(self, s):
        L = self.L
        if not s:
            return None
        s = C.c_void_p(s) if not isinstance(s, C.c_void_p) else s
        t = L.standin_type(s)
        if t == NILSXP:
            return None
        if t == EXTPTRSXP:
            return s
        n = L.standin_length(s)
        if t == VECSXP:
            return {L.standin_name(s, i).decode(): self.from_sexp(L.standin_elt(s, i)) for i in range(n)}
        if t == STRSXP:
            return [L.standin_string(s, i).decode() for i in range(n)]
        nd = L.standin_ndim(s)
        shape = tuple(L.standin_dim(s, k) for k in range(nd)) if nd else (n,)
        if t == REALSXP:
            a = np.ctypeslib.as_array(L.standin_real_ptr(s), shape=(n,)).copy() if n else np.zeros(0)
        else:
            a = np.ctypeslib.as_array(L.standin_int_ptr(s), shape=(n,)).copy() if n else np.zeros(0, np.int32)
        return a.reshape(shape, order="F")

    def call(self, name, *args):
        """.Call(name, ...)"""
        sx = [self.to_sexp(a) for a in args]
        arr = (C.c_void_p * max(len(sx), 1))(*[s.value for s in sx])
        r = self.L.standin_call(name.encode(), len(sx), arr)
        cond = self.L.standin_condition()
        if cond is not None:
            raise RError(cond.decode())
        return self.from_sexp(r)

    def set_interrupt(self, pending):
        self.L.standin_set_interrupt(int(pending))

    def interrupt_raised(self):
        return bool(self.L.standin_interrupt_raised())


def r_call_sites(path):
    """(.Call name, number of arguments after the name) for every .fgs_call(...) in an R file"""
    txt = open(path).read()
    out = []
    for m in re.finditer(r"\.fgs_call\('([A-Za-z0-9_]+)'", txt):
        i, depth, nargs = m.end(), 1, 0
        in_str = None
        while depth:
            ch = txt[i]
            if in_str:
                if ch == in_str:
                    in_str = None
            elif ch in "'\"":
                in_str = ch
            elif ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            i += 1
        out.append((m.group(1), nargs))
    return out


def r_formals(path, fn):
    """argument names of `fn <- function(...)` in an R file"""
    txt = open(path).read()
    m = re.search(re.escape(fn) + r"\s*<-\s*function\s*\(", txt)
    i, depth, cur, out = m.end(), 1, "", []
    while depth:
        ch = txt[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
            if depth == 0:
                break
        if ch == "," and depth == 1:
            out.append(cur); cur = ""
        else:
            cur += ch
        i += 1
    out.append(cur)
    return [a.split("=")[0].strip() for a in out]
