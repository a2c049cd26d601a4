"""Host-side mirror of `ExaModel` + the NLPModels callback surface over the C ABI.

Mirrors (reference file:line):
  * `ExaModel(c; prod)` and `build_extension`          src/nlp.jl:765-798, 898;
                                                      ext/ExaModelsKernelAbstractions.jl:33-191
  * `obj, cons_nln!, grad!, jac_structure!, jac_coord!, hess_structure!, hess_coord!`
                                                      src/nlp.jl:1798-1940 | ext:212-547
  * meta (`NLPModelMeta`): nvar, ncon, nnzj, nnzh, x0, lvar, uvar, y0, lcon, ucon, minimize

Every callback is ONE call into `libexa_b200.so` (include/exa_b200.h).  Arguments may be
torch CUDA tensors (device pointers are passed through, work is enqueued on torch's current
stream and not synchronised -- the KA callbacks never synchronise either) or numpy arrays
(the `exb_host_*` shims copy through pinned staging, like `WrapperNLPModel`,
src/utils.jl:16-267).  Mutating callbacks return their output argument, as in the reference.

There is no CPU evaluation path here: without the CUDA library / a B200 every callback raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIBPATH = os.path.join(_CSRC, "libexa_b200.so")
_LIB = None

EXB_FLAG_NO_COMPILE = 1
EXB_FLAG_SORTED_PRODUCTS = 2
EXB_FLAG_TUNE_AT_CREATE = 4
_ERRS = {1: "invalid handle", 2: "internal error", 3: "malformed IR", 4: "kernel module compile/load failed",
         5: "CUDA error / no device", 6: "bad argument"}

_SYMBOLS = """exb_plan_create exb_plan_destroy exb_plan_dims exb_plan_npatterns exb_plan_pattern exb_plan_comp
exb_plan_source exb_plan_module_path exb_plan_compile exb_create exb_destroy exb_dims exb_set_params exb_obj
exb_obj_async exb_grad exb_cons exb_jac_structure64 exb_jac_structure32 exb_jac exb_hess_structure64
exb_hess_structure32 exb_hess exb_host_obj exb_host_grad exb_host_cons exb_host_jac exb_host_hess
exb_host_jac_structure64 exb_host_hess_structure64 exb_shard exb_stats exb_last_error exb_abi_version
exb_jprod exb_jtprod exb_hprod exb_compressed_dims exb_jac_structure_compressed64 exb_hess_structure_compressed64
exb_jac_compressed exb_hess_compressed exb_set_timing exb_timings exb_kernel_choice exb_host_bytes
exb_comm_unique_id exb_comm_init exb_comm_attach exb_comm_destroy exb_comm_set_mode exb_comm_gather_coo exb_owned
exb_comm_stats exb_compressed_shard exb_jac_structure_compressed32 exb_hess_structure_compressed32
exb_host_jac_compressed exb_host_hess_compressed exb_plan_tile exb_eval exb_plan_create_data exb_tune exb_build_info exb_host_jprod exb_host_jtprod exb_host_hprod
exb_host_jac_structure32 exb_host_hess_structure32""".split()


class ExbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libexa_b200: {_ERRS.get(code, code)}: {msg}")
        self.code = code


def build_library(force=False):
    """Compile libexa_b200.so in-tree (nvcc cross-compiles for sm_100a without a GPU)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)
            if f.endswith((".cpp", ".cu", ".cuh", ".hpp", ".h")) and f != "exb_embed.cpp"]
    srcs.append(os.path.join(_HERE, "..", "include", "exa_b200.h"))
    def stale():
        return (not os.path.exists(_LIBPATH)
                or any(os.path.getmtime(s) > os.path.getmtime(_LIBPATH) for s in srcs))
    if force or stale():
        # the ranks of one job (torchrun) may all get here at once on a fresh checkout: one builds, the others wait
        import fcntl
        with open(os.path.join(_CSRC, ".build.lock"), "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.check_call(["make", "-s", "-C", _CSRC, "libexa_b200.so"])
            finally:
                fcntl.flock(lk, fcntl.LOCK_UN)
    return _LIBPATH


def lib():
    """Load the C-ABI library.  Raises if it is missing: there is no fallback path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIBPATH):
            raise ExbError(4, f"{_LIBPATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_LIBPATH)
        L.exb_last_error.restype = C.c_char_p
        for s in _SYMBOLS:
            getattr(L, s)  # every declared symbol must be exported
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise ExbError(rc, lib().exb_last_error().decode(errors="replace"))


class _Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("flags", C.c_int32),
                ("fuse_below", C.c_int64), ("tune_x0", C.c_void_p)]


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class Plan:
    """Host-only analysis of a core (no GPU needed): dims, per-pattern layout, generated source."""

    def __init__(self, core):
        self.ir, self.bufs = core.to_ir()
        self.h = C.c_void_p()
        arr = (C.c_void_p * max(1, len(self.bufs)))(*[b.ctypes.data for b in self.bufs])
        _check(lib().exb_plan_create_data(self.ir, C.c_size_t(len(self.ir)), arr, len(self.bufs), None, C.byref(self.h)))
        d = np.zeros(8, dtype=np.int64)
        _check(lib().exb_plan_dims(self.h, _np_ptr(d)))
        (self.nvar, self.ncon, self.nnzj, self.nnzh, self.nobj, self.nnzg, self.nconaug,
         self.npar) = (int(v) for v in d)

    def __del__(self):
        try:
            if self.h:
                lib().exb_plan_destroy(self.h)
        except Exception:
            pass

    def npatterns(self):
        return int(lib().exb_plan_npatterns(self.h))

    def pattern_info(self, k):
        o = np.zeros(9, dtype=np.int64)
        _check(lib().exb_plan_pattern(self.h, k, _np_ptr(o)))
        keys = ("kind", "nitr", "o0", "o1", "o2", "o1step", "o2step", "ncomp1", "ncomp2")
        return dict(zip(keys, (int(v) for v in o)))

    def comp(self, k, which):
        info = self.pattern_info(k)
        o = np.zeros(max(1, info["ncomp1" if which == 1 else "ncomp2"]), dtype=np.int64)
        _check(lib().exb_plan_comp(self.h, k, which, _np_ptr(o)))
        return o[: info["ncomp1" if which == 1 else "ncomp2"]]

    def tile_info(self):
        """Fused duplicate-free Hessian (column-tile kernel): applicability and closed-form unique count."""
        o = np.zeros(4, dtype=np.int64)
        _check(lib().exb_plan_tile(self.h, _np_ptr(o)))
        return {"fused": bool(o[0]), "nnzh_unique": int(o[1]), "distances": int(o[2]), "halo": int(o[3])}

    def source(self):
        src, n = C.c_char_p(), C.c_size_t()
        _check(lib().exb_plan_source(self.h, C.byref(src), C.byref(n)))
        return src.value.decode()

    def module_path(self):
        buf = C.create_string_buffer(4096)
        _check(lib().exb_plan_module_path(self.h, buf, C.c_size_t(4096)))
        return buf.value.decode()

    def compile(self):
        """Run nvcc for this model's kernel module if it is not cached (works without a GPU)."""
        _check(lib().exb_plan_compile(self.h))
        return self.module_path()


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class ExaModel:
    """`ExaModel(core)` on one B200 (or on shard `rank` of `world` when sharded)."""

    def __init__(self, core, device=None, rank=0, world=1, allow_compile=True, sorted_products=False, tune_at_create=False):
        import torch  # device memory + streams only

        self._torch = torch
        if not torch.cuda.is_available():
            raise ExbError(5, "no CUDA device: this evaluator has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        self.rank, self.world = int(rank), int(world)
        self.core = core
        meta = core.meta()
        self.meta = meta
        self.minimize = meta["minimize"]
        self.x0, self.lvar, self.uvar = meta["x0"], meta["lvar"], meta["uvar"]
        self.y0, self.lcon, self.ucon = meta["y0"], meta["lcon"], meta["ucon"]
        ir, bufs = core.to_ir()
        self._ir, self._bufs = ir, bufs
        arr = (C.c_void_p * max(1, len(bufs)))(*[b.ctypes.data for b in bufs])
        flags = ((0 if allow_compile else EXB_FLAG_NO_COMPILE) | (EXB_FLAG_SORTED_PRODUCTS if sorted_products else 0)
                 | (EXB_FLAG_TUNE_AT_CREATE if tune_at_create else 0))
        x0 = np.ascontiguousarray(meta["x0"], dtype=np.float64)
        opt = _Options(self.device.index, self.rank, self.world, flags, 0, x0.ctypes.data if tune_at_create else None)
        self.h = C.c_void_p()
        _check(lib().exb_create(ir, C.c_size_t(len(ir)), arr, len(bufs), C.byref(opt), C.byref(self.h)))
        d = np.zeros(8, dtype=np.int64)
        _check(lib().exb_dims(self.h, _np_ptr(d)))
        (self.nvar, self.ncon, self.nnzj, self.nnzh, self.nobj, self.nnzg, self.nconaug,
         self.npar) = (int(v) for v in d)
        self.npatterns = len(core.patterns)
        self.has_comm = False
        if self.npar:
            self.set_params(meta["theta"])

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().exb_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t, n, dtype=None):
        torch = self._torch
        dtype = torch.float64 if dtype is None else dtype
        assert t.is_cuda and t.dtype == dtype and t.is_contiguous() and t.numel() == n, \
            f"expected a contiguous CUDA {dtype} tensor of length {n}"
        return C.c_void_p(t.data_ptr())

    def _host(self, a, n, dtype=np.float64):
        assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and a.size == n, \
            f"expected a contiguous {dtype} array of length {n}"
        return _np_ptr(a)

    def new(self, n, dtype=None):
        torch = self._torch
        return torch.empty(n, dtype=torch.float64 if dtype is None else dtype, device=self.device)

    def set_params(self, theta):
        """`set_value!` on parameters (src/nlp.jl:1217-1287): upload θ."""
        t = np.ascontiguousarray(theta, dtype=np.float64)
        assert t.size == self.npar
        _check(lib().exb_set_params(self.h, _np_ptr(t), self._stream()))
        self._torch.cuda.current_stream(self.device).synchronize()

    # -- callbacks (src/nlp.jl:1798-1940) --------------------------------------------
    def obj(self, x):
        out = C.c_double()
        if _is_torch(x):
            _check(lib().exb_obj(self.h, self._dev(x, self.nvar), C.byref(out), self._stream()))
        else:
            _check(lib().exb_host_obj(self.h, self._host(x, self.nvar), C.byref(out)))
        return out.value

    def obj_async(self, x, out):
        """Objective left on the device (no synchronisation): `out` is a 1-element CUDA tensor."""
        _check(lib().exb_obj_async(self.h, self._dev(x, self.nvar), self._dev(out, 1), self._stream()))
        return out

    def grad(self, x, g):
        if _is_torch(x):
            _check(lib().exb_grad(self.h, self._dev(x, self.nvar), self._dev(g, self.nvar), self._stream()))
        else:
            _check(lib().exb_host_grad(self.h, self._host(x, self.nvar), self._host(g, self.nvar)))
        return g

    def cons_nln(self, x, c):
        if _is_torch(x):
            _check(lib().exb_cons(self.h, self._dev(x, self.nvar), self._dev(c, self.ncon), self._stream()))
        else:
            _check(lib().exb_host_cons(self.h, self._host(x, self.nvar), self._host(c, self.ncon)))
        return c

    cons = cons_nln

    def jac_coord(self, x, vals):
        if _is_torch(x):
            _check(lib().exb_jac(self.h, self._dev(x, self.nvar), self._dev(vals, self.nnzj), self._stream()))
        else:
            _check(lib().exb_host_jac(self.h, self._host(x, self.nvar), self._host(vals, self.nnzj)))
        return vals

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        """`hess_coord!(m, x, y, hess; obj_weight)`; `y=None` is the objective-only form."""
        w = C.c_double(float(obj_weight))
        if _is_torch(x):
            yp = None if y is None else self._dev(y, self.ncon)
            _check(lib().exb_hess(self.h, self._dev(x, self.nvar), yp, w, self._dev(vals, self.nnzh), self._stream()))
        else:
            yp = None if y is None else self._host(y, self.ncon)
            _check(lib().exb_host_hess(self.h, self._host(x, self.nvar), yp, w, self._host(vals, self.nnzh)))
        return vals

    def eval_all(self, x, y, obj_out, g, c, jac, hess, obj_weight=1.0, mask=31):
        """obj + grad! + cons! + jac_coord! + hess_coord! at the same x from ONE sweep (exb_eval): every data point is
        evaluated once.  `obj_out` is a 1-element CUDA tensor (no synchronisation).  `mask` selects callbacks (bit 0 obj,
        1 grad, 2 cons, 3 jac, 4 hess); outputs that are not requested may be None."""
        def p(t, n):
            return None if t is None else self._dev(t, n)
        _check(lib().exb_eval(self.h, C.c_uint(mask), self._dev(x, self.nvar), p(y, self.ncon), C.c_double(float(obj_weight)),
                              p(obj_out, 1), p(g, self.nvar), p(c, self.ncon), p(jac, self.nnzj), p(hess, self.nnzh), self._stream()))
        return obj_out, g, c, jac, hess

    def _structure(self, which, n, rows, cols):
        torch = self._torch
        if _is_torch(rows):
            assert rows.dtype == cols.dtype and rows.dtype in (torch.int64, torch.int32)
            f = getattr(lib(), f"exb_{which}_structure{64 if rows.dtype == torch.int64 else 32}")
            _check(f(self.h, self._dev(rows, n, rows.dtype), self._dev(cols, n, cols.dtype), self._stream()))
        else:
            assert rows.dtype == cols.dtype and rows.dtype in (np.int64, np.int32)
            f = getattr(lib(), f"exb_host_{which}_structure{64 if rows.dtype == np.int64 else 32}")
            _check(f(self.h, self._host(rows, n, rows.dtype), self._host(cols, n, cols.dtype)))
        return rows, cols

    def jac_structure(self, rows, cols):
        return self._structure("jac", self.nnzj, rows, cols)

    def hess_structure(self, rows, cols):
        return self._structure("hess", self.nnzh, rows, cols)

    # -- matrix-free products (src/nlp.jl:1882-1978 | ext:353-511) ---------------------------
    def jprod_nln(self, x, v, Jv):
        if not _is_torch(x):
            _check(lib().exb_host_jprod(self.h, self._host(x, self.nvar), self._host(v, self.nvar), self._host(Jv, self.ncon)))
            return Jv
        _check(lib().exb_jprod(self.h, self._dev(x, self.nvar), self._dev(v, self.nvar), self._dev(Jv, self.ncon), self._stream()))
        return Jv

    def jtprod_nln(self, x, v, Jtv):
        if not _is_torch(x):
            _check(lib().exb_host_jtprod(self.h, self._host(x, self.nvar), self._host(v, self.ncon), self._host(Jtv, self.nvar)))
            return Jtv
        _check(lib().exb_jtprod(self.h, self._dev(x, self.nvar), self._dev(v, self.ncon), self._dev(Jtv, self.nvar), self._stream()))
        return Jtv

    def hprod(self, x, y, v, Hv, obj_weight=1.0):
        if not _is_torch(x):
            yp = None if y is None else self._host(y, self.ncon)
            _check(lib().exb_host_hprod(self.h, self._host(x, self.nvar), yp, self._host(v, self.nvar), C.c_double(float(obj_weight)),
                                        self._host(Hv, self.nvar)))
            return Hv
        yp = None if y is None else self._dev(y, self.ncon)
        _check(lib().exb_hprod(self.h, self._dev(x, self.nvar), yp, self._dev(v, self.nvar), C.c_double(float(obj_weight)),
                               self._dev(Hv, self.nvar), self._stream()))
        return Hv

    def capture_fused_eval(self, x, y, obj_out, g, c, jac, hess, obj_weight=1.0):
        """`eval_all` (one sweep + its finishing steps) at fixed buffers as ONE CUDA graph: for latency-bound models the 3-4
        launches of the fused evaluation replay without per-launch submission cost.  Returns the `torch.cuda.CUDAGraph`."""
        torch = self._torch
        for _ in range(2):         # first call ranks the launch-shape variants (synchronises): not capturable
            self.eval_all(x, y, obj_out, g, c, jac, hess, obj_weight=obj_weight)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.eval_all(x, y, obj_out, g, c, jac, hess, obj_weight=obj_weight)
        return graph

    def capture_full_eval(self, x, y, obj_out, g, c, jac, hess, obj_weight=1.0):
        """Capture obj + grad! + cons! + jac_coord! + hess_coord! at fixed buffers into ONE CUDA graph.
        For the many-small-patterns regime (AC-OPF: ~0.6 M nonzeros) a full evaluation is launch-latency
        bound; replaying the graph costs one submission instead of 7-8 launches.  Returns the
        `torch.cuda.CUDAGraph`; update `x` / `y` in place and call `.replay()`."""
        torch = self._torch

        calls = [lambda: self.obj_async(x, obj_out), lambda: self.grad(x, g), lambda: self.cons_nln(x, c),
                 lambda: self.jac_coord(x, jac), lambda: self.hess_coord(x, y, hess, obj_weight=obj_weight)]
        for f in calls:            # first calls pick the launch-shape variants (and synchronise); not capturable
            f()
        torch.cuda.synchronize(self.device)
        side = [torch.cuda.Stream(self.device) for _ in calls[1:]]

        def run():                 # the five callbacks are independent: fork them onto parallel branches
            cur = torch.cuda.current_stream(self.device)
            for st, f in zip(side, calls[1:]):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    f()
            calls[0]()
            for st in side:
                cur.wait_stream(st)
        warm = torch.cuda.Stream(self.device)
        warm.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(warm):
            run()
        torch.cuda.current_stream(self.device).wait_stream(warm)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        return graph

    def compressed(self):
        """`CompressedNLPModel(m)` (src/utils.jl:425-579): the same model with duplicate COO entries summed."""
        return CompressedExaModel(self)

    def host_hess_compressed(self, x, y, vals, obj_weight=1.0):
        """Duplicate-free `hess_coord!` with HOST buffers (exb_host_hess_compressed): the D2H copy carries the unique entries only."""
        yp = None if y is None else self._host(y, self.ncon)
        _check(lib().exb_host_hess_compressed(self.h, self._host(x, self.nvar), yp, C.c_double(float(obj_weight)), _np_ptr(vals)))
        return vals

    def host_jac_compressed(self, x, vals):
        _check(lib().exb_host_jac_compressed(self.h, self._host(x, self.nvar), _np_ptr(vals)))
        return vals

    def compressed_shard(self):
        """(lo, hi, fused): range of the duplicate-free Hessian values this handle writes; fused = emitted by one launch."""
        o = np.zeros(3, dtype=np.int64)
        _check(lib().exb_compressed_shard(self.h, _np_ptr(o)))
        return int(o[0]), int(o[1]), bool(o[2])

    # -- multi-GPU: the library's own communicator (include/exa_b200.h "multi-GPU") ---------------------
    def comm_init(self, group=None, mode="replicate"):
        """Create the NCCL communicator of this sharded model inside the library.  The 128-byte ncclUniqueId made by rank 0
        travels through `torch.distributed` (any backend) -- a Julia host would send it over MPI.  Collective."""
        import torch.distributed as dist
        ids = [None]
        if dist.get_rank(group) == 0:
            buf = C.create_string_buffer(128)
            _check(lib().exb_comm_unique_id(buf))
            ids[0] = buf.raw
        dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        _check(lib().exb_comm_init(self.h, C.c_char_p(ids[0])))
        self.has_comm = True
        self.comm_set_mode(mode)
        return self

    def comm_set_mode(self, mode):
        _check(lib().exb_comm_set_mode(self.h, {"replicate": 0, "owner": 1}[mode]))
        self.comm_mode = mode

    def comm_destroy(self):
        _check(lib().exb_comm_destroy(self.h))
        self.has_comm = False

    def gather_coo(self, which, vals):
        """Replicate the sharded COO values in place (which = 1: jac, 2: hess)."""
        n = self.nnzj if which == 1 else self.nnzh
        _check(lib().exb_comm_gather_coo(self.h, which, self._dev(vals, n), self._stream()))
        return vals

    def owned(self):
        """0-based half-open range of the variables this handle owns."""
        o = np.zeros(2, dtype=np.int64)
        _check(lib().exb_owned(self.h, _np_ptr(o)))
        return int(o[0]), int(o[1])

    def comm_stats(self):
        o = np.zeros(4, dtype=np.int64)
        _check(lib().exb_comm_stats(self.h, _np_ptr(o)))
        return dict(zip(("collectives", "last_collectives", "attached", "mode"), (int(v) for v in o)))

    # -- sharding / introspection ------------------------------------------------------
    def shard(self, k):
        o = np.zeros(6, dtype=np.int64)
        _check(lib().exb_shard(self.h, k, _np_ptr(o)))
        return dict(zip(("lo", "hi", "jac_lo", "jac_hi", "hess_lo", "hess_hi"), (int(v) for v in o)))

    def set_timing(self, on=True):
        """Per-callback device timing (the `TimedNLPModel` role, src/utils.jl:271-408)."""
        _check(lib().exb_set_timing(self.h, 1 if on else 0))

    def timings(self, reset=False):
        ms, calls = np.zeros(8), np.zeros(8, dtype=np.int64)
        _check(lib().exb_timings(self.h, _np_ptr(ms), _np_ptr(calls), 1 if reset else 0))
        names = ("obj", "grad", "cons", "jac", "hess", "jprod", "jtprod", "hprod")
        return {n: {"ms": float(m), "calls": int(c)} for n, m, c in zip(names, ms, calls)}

    def kernel_choice(self, callback):
        """Which generated kernel `callback` ('obj' | 'grad' | 'cons' | 'jac' | 'hess') launches (first-call tuner's verdict)."""
        o = np.zeros(4, dtype=np.int64)
        _check(lib().exb_kernel_choice(self.h, ("obj", "grad", "cons", "jac", "hess").index(callback), _np_ptr(o)))
        return {"min_blocks": int(o[0]), "persistent": bool(o[1]), "grid": int(o[2]), "owner_computes": bool(o[3])}

    def host_bytes(self):
        """(H2D, D2H) bytes moved by the last host-buffer callback on this handle."""
        o = np.zeros(2, dtype=np.int64)
        _check(lib().exb_host_bytes(self.h, _np_ptr(o)))
        return int(o[0]), int(o[1])

    def tune(self, x, y=None):
        """Rank the launch-shape variants of every kernel now (exb_tune), at the CUDA tensors x / y."""
        _check(lib().exb_tune(self.h, self._dev(x, self.nvar), None if y is None else self._dev(y, self.ncon), self._stream()))

    def build_info(self):
        o = np.zeros(5)
        _check(lib().exb_build_info(self.h, _np_ptr(o)))
        return dict(zip(("plan_s", "nvcc_s", "load_s", "tune_s", "create_s"), (float(v) for v in o)))

    def stats(self):
        o = np.zeros(4, dtype=np.int64)
        _check(lib().exb_stats(self.h, _np_ptr(o)))
        return dict(zip(("launches", "last_launches", "device_bytes", "module_cached"), (int(v) for v in o)))


class CompressedExaModel:
    """Duplicate-free COO view of an `ExaModel`: unique (row, col) coordinates in the reference's order
    (sorted by (col, row), src/utils.jl:478-487,509-510), duplicate values summed (`_compress!`, :564-571).
    The Hessian of a shift-indexed model is emitted duplicate-free by ONE launch (`fused_hess`); see include/exa_b200.h."""

    def __init__(self, inner):
        self.inner = inner
        nh = C.c_int64()
        _check(lib().exb_compressed_dims(inner.h, None, C.byref(nh)))
        self.nvar, self.ncon, self.nnzh = inner.nvar, inner.ncon, int(nh.value)
        self._nnzj = None
        self.hess_lo, self.hess_hi, self.fused_hess = inner.compressed_shard()
        for a in ("obj", "grad", "cons_nln", "cons", "new"):
            setattr(self, a, getattr(inner, a))

    @property
    def nnzj(self):
        if self._nnzj is None:
            nj = C.c_int64()
            _check(lib().exb_compressed_dims(self.inner.h, C.byref(nj), None))
            self._nnzj = int(nj.value)
        return self._nnzj

    def _structure(self, which, n, rows, cols):
        i = self.inner
        torch = i._torch
        assert rows.dtype == cols.dtype and rows.dtype in (torch.int64, torch.int32), "structures are int64 or int32 tensors"
        f = getattr(lib(), f"exb_{which}_structure_compressed{64 if rows.dtype == torch.int64 else 32}")
        _check(f(i.h, i._dev(rows, n, rows.dtype), i._dev(cols, n, cols.dtype), i._stream()))
        return rows, cols

    def jac_structure(self, rows, cols):
        return self._structure("jac", self.nnzj, rows, cols)

    def hess_structure(self, rows, cols):
        return self._structure("hess", self.nnzh, rows, cols)

    def jac_coord(self, x, vals):
        i = self.inner
        if _is_torch(x):
            _check(lib().exb_jac_compressed(i.h, i._dev(x, self.nvar), i._dev(vals, self.nnzj), i._stream()))
        else:
            i.host_jac_compressed(x, i._host(vals, self.nnzj) and vals)
        return vals

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        i = self.inner
        if _is_torch(x):
            yp = None if y is None else i._dev(y, self.ncon)
            _check(lib().exb_hess_compressed(i.h, i._dev(x, self.nvar), yp, C.c_double(float(obj_weight)), i._dev(vals, self.nnzh), i._stream()))
        else:
            i.host_hess_compressed(x, y, i._host(vals, self.nnzh) and vals, obj_weight=obj_weight)
        return vals
