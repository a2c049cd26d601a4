"""ctypes binding of oracle/libexa_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libexa_oracle.so")
    src = os.path.join(_HERE, "exa_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libexa_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ora_create.restype = C.c_void_p
        L.ora_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int]
        L.ora_error.restype = C.c_char_p
        L.ora_error.argtypes = [C.c_void_p]
        L.ora_obj.restype = C.c_double
        for name in ("ora_destroy", "ora_dims", "ora_set_params"):
            getattr(L, name).restype = None
        L.ora_emit_source.restype = C.c_void_p
        L.ora_emit_source.argtypes = [C.c_void_p]
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_free.restype = None
        _LIB = L
    return _LIB


class CompiledOracle:
    """The per-pattern straight-line port emitted by the oracle itself (exa_oracle.cpp `ora_emit_source`) and compiled
    with the interpreter's own flags: the stand-in for the native code Julia generates for src/hessian.jl:681-712.
    Values are bit-identical to the interpreter's.  TEST / BASELINE INFRASTRUCTURE ONLY."""

    def __init__(self, ora):
        import hashlib
        L = lib()
        ptr = L.ora_emit_source(ora.h)
        src = C.string_at(ptr).decode()
        L.ora_free(ptr)
        self.source = src
        gen = os.path.join(_HERE, "_gen")
        os.makedirs(gen, exist_ok=True)
        with open(os.path.join(_HERE, "ora_tables.hpp"), "rb") as f:
            tag = hashlib.sha1(src.encode() + f.read()).hexdigest()[:16]
        so, cpp = os.path.join(gen, f"ora_{tag}.so"), os.path.join(gen, f"ora_{tag}.cpp")
        if not os.path.exists(so):
            with open(cpp, "w") as f:
                f.write(src)
            tmp = so + f".tmp{os.getpid()}"
            subprocess.check_call(["/usr/bin/g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-pthread", "-fPIC", "-std=c++17",
                                   "-fext-numeric-literals", "-w", "-I", _HERE, "-shared", "-o", tmp, cpp])
            os.replace(tmp, so)
        self.so = so
        self.lib = C.CDLL(so)
        self.lib.cmp_hess.restype = None
        self.ora = ora
        self._data = (C.c_void_p * max(1, len(ora._bufs)))(*[b.ctypes.data for b in ora._bufs])

    def hess_coord(self, x, y=None, obj_weight=1.0, out=None):
        o = self.ora
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        out = np.empty(o.nnzh, dtype=np.float64) if out is None else out
        th = o._theta if getattr(o, "_theta", None) is not None else np.zeros(max(1, o.npar))
        self.lib.cmp_hess(_p(x), _p(y), _p(th), C.c_double(obj_weight), _p(out), self._data,
                          C.c_int(getattr(o, "_threads", 1)), C.c_int(getattr(o, "_rank", 0)), C.c_int(getattr(o, "_world", 1)))
        return out


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """CPU restatement of the reference callbacks over a pattern IR."""

    def __init__(self, ir: bytes, bufs, theta=None):
        L = lib()
        self._ir = ir
        self._bufs = [np.ascontiguousarray(b) for b in bufs]
        arr = (C.c_void_p * max(1, len(self._bufs)))(*[b.ctypes.data for b in self._bufs])
        self.h = C.c_void_p(L.ora_create(ir, len(ir), arr, len(self._bufs)))
        err = L.ora_error(self.h).decode()
        if err:
            raise RuntimeError("oracle: " + err)
        d = np.zeros(8, dtype=np.int64)
        L.ora_dims(self.h, _p(d))
        (self.nvar, self.ncon, self.nnzj, self.nnzh, self.nobj, self.nnzg, self.nconaug,
         self.npar) = (int(v) for v in d)
        if theta is not None and self.npar:
            self.set_params(theta)

    @classmethod
    def from_core(cls, core):
        ir, bufs = core.to_ir()
        return cls(ir, bufs, core.meta()["theta"])

    def __del__(self):
        try:
            lib().ora_destroy(self.h)
        except Exception:
            pass

    def compile(self):
        """g++-compiled straight-line form of this model's hess_coord! (see CompiledOracle)."""
        return CompiledOracle(self)

    def set_threads(self, n):
        self._threads = int(n)
        lib().ora_set_threads(self.h, int(n))

    def set_shard(self, rank, world):
        """Test aid: evaluate only shard `rank` of `world` of every pattern's iterator."""
        self._rank, self._world = int(rank), int(world)
        lib().ora_set_shard(self.h, int(rank), int(world))

    @staticmethod
    def max_threads():
        return int(lib().ora_max_threads())

    def set_params(self, theta):
        t = np.ascontiguousarray(theta, dtype=np.float64)
        assert t.size == self.npar
        self._theta = t
        lib().ora_set_params(self.h, _p(t))

    def npatterns(self):
        return int(lib().ora_npatterns(self.h))

    def pattern_info(self, k):
        o = np.zeros(9, dtype=np.int64)
        lib().ora_pattern_info(self.h, k, _p(o))
        keys = ("kind", "nitr", "o0", "o1", "o2", "o1step", "o2step", "ncomp1", "ncomp2")
        return dict(zip(keys, (int(v) for v in o)))

    def comp(self, k, which):
        info = self.pattern_info(k)
        o = np.zeros(info["ncomp1" if which == 1 else "ncomp2"], dtype=np.int64)
        lib().ora_pattern_comp(self.h, k, which, _p(o))
        return o

    # -- callbacks ---------------------------------------------------------
    def obj(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return float(lib().ora_obj(self.h, _p(x)))

    def _call(self, name, n, x, *extra):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(n, dtype=np.float64)
        getattr(lib(), name)(self.h, _p(x), *extra, _p(out))
        return out

    def cons(self, x):
        return self._call("ora_cons", self.ncon, x)

    def grad(self, x):
        return self._call("ora_grad", self.nvar, x)

    def sgrad(self, x):
        return self._call("ora_sgrad", self.nnzg, x)

    def jac_coord(self, x):
        return self._call("ora_jac", self.nnzj, x)

    def hess_coord(self, x, y=None, obj_weight=1.0, out=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        out = np.empty(self.nnzh, dtype=np.float64) if out is None else out
        lib().ora_hess(self.h, _p(x), _p(y), C.c_double(obj_weight), _p(out))
        return out

    def jac_structure(self):
        r = np.zeros(self.nnzj, dtype=np.int64)
        c = np.zeros(self.nnzj, dtype=np.int64)
        lib().ora_jac_structure(self.h, _p(r), _p(c))
        return r, c

    def hess_structure(self):
        r = np.zeros(self.nnzh, dtype=np.int64)
        c = np.zeros(self.nnzh, dtype=np.int64)
        lib().ora_hess_structure(self.h, _p(r), _p(c))
        return r, c

    def jprod(self, x, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        return self._call("ora_jprod", self.ncon, x, _p(v))

    def jtprod(self, x, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        return self._call("ora_jtprod", self.nvar, x, _p(v))

    def hprod(self, x, y, v, obj_weight=1.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty(self.nvar, dtype=np.float64)
        lib().ora_hprod(self.h, _p(x), _p(y), _p(v), C.c_double(obj_weight), _p(out))
        return out


def uni(op: int, x: float):
    o = np.zeros(3)
    lib().ora_uni(op, C.c_double(x), _p(o))
    return o


def bi(op: int, x1: float, x2: float, e2_is_int=False):
    o = np.zeros(6)
    lib().ora_bi(op, C.c_double(x1), C.c_double(x2), int(e2_is_int), _p(o))
    return o
