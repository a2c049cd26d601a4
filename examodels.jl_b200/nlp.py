"""Model builder: the host-side mirror of `ExaCore` and `add_var / add_par / add_obj /
add_con / add_con!` (/root/reference/src/nlp.jl:328-1796), plus the emitter of the
language-neutral pattern IR consumed by the C ABI (`include/exa_b200.h`).

What is mirrored (and why it matters for bit-exact structure):

* variable / parameter indexing builds `Var(Node2(+, i, o))` with the block
  offset as a plain Int child, even when it is 0 (nlp.jl:900-926, 954-962);
  multi-dimensional indexing goes through the column-major `idxx`
  (nlp.jl:2012-2015);
* patterns are kept in ADD ORDER; their offsets (`o0, o1, o2`) are running
  counters at add time (nlp.jl:1448-1482, 1551-1611, 1680-1738).  The counters
  need each pattern's `o1step/o2step`, which come from the sparsity probe
  (simdfunction.jl:78-100); that probe lives behind the C ABI, so offsets are
  assigned there by the same rules and reported back through `exb_plan_*`;
* iterators: `range` stays a range, product iterators are collected in
  column-major order (nlp.jl:2165-2184), arrays of isbits elements are AoS.

Nothing here touches the GPU.
"""
from __future__ import annotations

import struct

import numpy as np

from . import graph as G

IR_MAGIC = 0x0031425845  # "EXB1"
IR_VERSION = 1

KIND_OBJ, KIND_CON, KIND_AUG = 0, 1, 2
ITR_RANGE, ITR_AOS = 0, 1
FT_I64, FT_F64, FT_I32, FT_F32 = 0, 1, 2, 3
(T_CONST_I, T_CONST_F, T_DATA_SELF, T_DATA_FIELD, T_VAR, T_PAR, T_NULL, T_OP1, T_OP2,
 T_VAL) = range(10)

_FT_OF = {np.dtype("int64"): FT_I64, np.dtype("float64"): FT_F64,
          np.dtype("int32"): FT_I32, np.dtype("float32"): FT_F32}


def _start(s):
    return s.start if isinstance(s, range) else 1


def _length(s):
    return len(s) if isinstance(s, range) else int(s)


def idxx(coord, si):
    """nlp.jl:2012-2015 — column-major linear index; works on ints and on nodes."""

    def _idxx(coord, si, a):
        if not coord:
            return 0
        return a * (coord[0] - 1) + _idxx(coord[1:], si[1:], a * si[0])

    return _idxx(tuple(coord), tuple(si), 1) + 1


class Variable:
    """nlp.jl `Variable(size, length, offset, name, tag)`; indexing per nlp.jl:900-926."""

    def __init__(self, size, length, offset, name="x"):
        self.size, self.length, self.offset, self.name = tuple(size), length, offset, name

    def _check(self, dim, i):
        if G._is_int(i):
            s = self.size[dim]
            ok = (i in s) if isinstance(s, range) else (1 <= i <= s)
            assert ok, "Variable index bound error"

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            self._check(0, idx)
            o = self.offset - _start(self.size[0]) + 1
            if isinstance(idx, G.AbstractNode):
                return G.Var(G.Node2("+", idx, o))  # nlp.jl:908
            return G.Var(idx + o)  # nlp.jl:922
        assert len(idx) == len(self.size), "Variable index dimension error"
        for d, i in enumerate(idx):
            self._check(d, i)
        shifted = tuple(i - (_start(s) - 1) for i, s in zip(idx, self.size))
        return G.Var(self.offset + idxx(shifted, tuple(_length(s) for s in self.size)))


class Parameter:
    """nlp.jl:954-962."""

    def __init__(self, size, length, offset):
        self.size, self.length, self.offset = tuple(size), length, offset

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            return G.ParameterNode(idx + (self.offset - _start(self.size[0]) + 1))
        assert len(idx) == len(self.size), "Parameter index dimension error"
        shifted = tuple(i - (_start(s) - 1) for i, s in zip(idx, self.size))
        return G.ParameterNode(self.offset + idxx(shifted, tuple(_length(s) for s in self.size)))


class _Pattern:
    """One `Objective` / `Constraint` / `ConstraintAugmentation` (nlp.jl:107-177)."""

    def __init__(self, kind, tree, itr, index=None, base=None, idx=(), dims=(), size=()):
        self.kind, self.tree, self.itr = kind, tree, itr
        self.index = index  # position in add order
        self.base, self.idx, self.dims = base, tuple(idx), tuple(dims)
        self.size = tuple(size)  # Constraint.size (ints or ranges), nlp.jl:137-143
        self.nitr = len(itr)

    # `g[i] += expr` sugar: returns the adjusted slot index (nlp.jl:232-245)
    def __getitem__(self, idx):
        if self.kind == KIND_AUG:
            return _Slot(self, idx if isinstance(idx, tuple) else (idx,))
        if not isinstance(idx, tuple):
            return _Slot(self, (idx - (_start(self.size[0]) - 1),))
        return _Slot(self, tuple(_con_adjust(i, _start(s)) for i, s in zip(idx, self.size)))


def _con_adjust(idx, start):
    """nlp.jl:262-263."""
    if G._is_int(idx):
        return G.Constant(idx - start + 1)
    return idx - (start - 1)


class _Slot:
    """nlp.jl `ConstraintSlot`; `slot + expr` -> (con, idx => expr) (nlp.jl:243-248)."""

    def __init__(self, con, idx):
        self.con, self.idx = con, idx

    def _pair(self, expr):
        idx = self.idx[0] if len(self.idx) == 1 else self.idx
        return _AugPair(self.con, idx, expr)

    def __add__(self, e):
        return self._pair(e if isinstance(e, G.AbstractNode) else G.Null(e))

    __radd__ = __add__

    def __sub__(self, e):
        return self._pair(-e if isinstance(e, G.AbstractNode) else G.Null(-e))


class _AugPair:
    def __init__(self, con, idx, expr):
        self.con, self.idx, self.expr = con, idx, expr


class Iterator:
    """An adapted iterator (nlp.jl:2165-2184): a unit range or an AoS numpy array."""

    def __init__(self, src):
        self.range = None
        self.array = None
        self.shape = None
        self.ranges = None  # product iterators remember their ranges (nlp.jl:1743)
        if isinstance(src, Iterator):
            self.__dict__.update(src.__dict__)
            return
        if isinstance(src, range):
            assert src.step == 1, "only unit ranges are kept as ranges"
            self.range = src
            self.shape = (len(src),)
            return
        if isinstance(src, ProductArray):
            self.ranges = src.ranges
            src = src.array
        if isinstance(src, np.ndarray):
            arr = src
            self.shape = arr.shape
            arr = arr.reshape(-1, order="F")  # Julia arrays are column-major
        else:
            arr = _to_struct_array(list(src))
            self.shape = (len(arr),)
        if arr.dtype.names is None:
            if np.issubdtype(arr.dtype, np.integer):
                arr = arr.astype(np.int64)
            else:
                arr = arr.astype(np.float64)
        self.array = np.ascontiguousarray(arr)

    def __len__(self):
        return len(self.range) if self.range is not None else len(self.array)

    @property
    def sizes(self):
        """`_infer_subexpr_dims` (nlp.jl:1741-1744): ranges for ranges/products, else size."""
        if self.range is not None:
            return (self.range,)
        return tuple(self.ranges) if self.ranges is not None else tuple(self.shape)

    @property
    def layout(self):
        if self.range is not None or self.array.dtype.names is None:
            return None
        return self.array.dtype


def _to_struct_array(items):
    """list of ints / floats / tuples / dicts / namedtuples -> numpy (structured) array."""
    if not items:
        return np.zeros(0, dtype=np.int64)
    first = items[0]
    if G._is_real(first):
        return np.array(items, dtype=np.int64 if all(G._is_int(v) for v in items) else np.float64)
    if isinstance(first, dict):
        names = list(first.keys())
        rows = [tuple(d[n] for n in names) for d in items]
    elif hasattr(first, "_fields"):
        names = list(first._fields)
        rows = [tuple(d) for d in items]
    else:
        names = [f"f{k + 1}" for k in range(len(first))]
        rows = [tuple(d) for d in items]

    def ftype(col):
        return np.int64 if all(G._is_int(v) for v in col) else np.float64

    cols = list(zip(*rows))
    dt = np.dtype([(n, ftype(c)) for n, c in zip(names, cols)])
    return np.array(rows, dtype=dt)


class ProductArray:
    """A collected `Iterators.product` that remembers its ranges."""

    def __init__(self, array, ranges):
        self.array, self.ranges = array, tuple(ranges)


def product(*ranges):
    """`Iterators.product(r1, r2, ...)` collected (nlp.jl:2178-2181): first index fastest."""
    grids = np.meshgrid(*[np.arange(r.start, r.stop, dtype=np.int64) for r in ranges],
                        indexing="ij")
    dt = np.dtype([(f"f{k + 1}", np.int64) for k in range(len(ranges))])
    arr = np.zeros(grids[0].shape, dtype=dt)
    for k, g in enumerate(grids):
        arr[f"f{k + 1}"] = g
    return ProductArray(arr, ranges)


def _expand(v, n, default):
    if v is None:
        v = default
    if np.isscalar(v):
        return np.full(n, float(v))
    a = np.asarray(list(v) if not isinstance(v, np.ndarray) else v, dtype=np.float64)
    a = a.reshape(-1, order="F")
    assert a.size == n, f"expected {n} values, got {a.size}"
    return a


class ExaCore:
    """Progressive model store (nlp.jl:328-571).  Mutating: each `add_*` returns the handle."""

    def __init__(self, minimize=True, name="Generic"):
        self.minimize, self.name = minimize, name
        self.nvar = self.npar = self.ncon = self.nconaug = self.nobj = 0
        self.x0, self.lvar, self.uvar = [], [], []
        self.theta = []
        self.y0, self.lcon, self.ucon = [], [], []
        self.patterns: list[_Pattern] = []
        self.vars: list[Variable] = []
        self.pars: list[Parameter] = []

    # -- variables / parameters ------------------------------------------------
    def add_var(self, *ns, start=0.0, lvar=-np.inf, uvar=np.inf, name="x"):
        """nlp.jl:1116-1140."""
        n = int(np.prod([_length(s) for s in ns])) if ns else 1
        v = Variable(ns if ns else (1,), n, self.nvar, name)
        self.nvar += n
        self.x0.append(_expand(start, n, 0.0))
        self.lvar.append(_expand(lvar, n, -np.inf))
        self.uvar.append(_expand(uvar, n, np.inf))
        self.vars.append(v)
        return v

    def add_par(self, *ns, value=0.0):
        """nlp.jl:1177-1215: `add_par(core, dims...; value)` or `add_par(core, array)`."""
        if len(ns) == 1 and isinstance(ns[0], (np.ndarray, list, tuple)):
            value = np.asarray(ns[0], dtype=np.float64)
            ns = value.shape
        n = int(np.prod([_length(s) for s in ns]))
        p = Parameter(ns, n, self.npar)
        self.npar += n
        self.theta.append(_expand(value, n, 0.0))
        self.pars.append(p)
        return p

    # -- patterns ----------------------------------------------------------------
    def _probe(self, body, itr):
        it = Iterator(itr)
        tree = body(G.DataSource(it.layout)) if callable(body) else body
        return tree, it

    def _push(self, pat):
        pat.index = len(self.patterns)
        self.patterns.append(pat)
        return pat

    def add_obj(self, body, itr=range(1, 2)):
        """nlp.jl:1448-1482 (`gen.f(DataSource())` is `body(DataSource())`)."""
        tree, it = self._probe(body, itr)
        self.nobj += len(it)
        return self._push(_Pattern(KIND_OBJ, tree, it))

    def add_con(self, body=None, itr=None, *, dims=None, start=0.0, lcon=0.0, ucon=0.0):
        """nlp.jl:1551-1611.  The reference's `add_con(core, n1, n2, ...)` (empty rows, to be
        filled by `add_con!`) is `add_con(dims=(n1, n2, ...))`; dims may be ints or ranges."""
        if dims is not None:
            dims = tuple(dims) if isinstance(dims, (tuple, list)) else (dims,)
            # nlp.jl:1574,1584-1585: Null rows over 1:n or a collected product
            if len(dims) == 1:
                it = Iterator(range(1, _length(dims[0]) + 1))
            else:
                it = Iterator(product(*[range(1, _length(d) + 1) for d in dims]))
            tree, size = G.Null(None), dims
        else:
            if not callable(body) and itr is None:
                itr = range(1, 2)  # nlp.jl:1576: a single node -> one row
            tree, it = self._probe(body, itr)
            size = it.sizes  # nlp.jl:1741-1744 (_infer_subexpr_dims)
        n = len(it)
        self.y0.append(_expand(start, n, 0.0))
        self.lcon.append(_expand(lcon, n, 0.0))
        self.ucon.append(_expand(ucon, n, 0.0))
        pat = _Pattern(KIND_CON, tree, it, size=size)
        pat.offset = self.ncon
        self.ncon += n
        return self._push(pat)

    def add_con_aug(self, con_or_body, body=None, itr=None):
        """`add_con!` (nlp.jl:1680-1738).

        Forms: `add_con_aug(con, lambda d: (idx, expr), itr)`          (idx => expr)
               `add_con_aug(lambda d: g[idx] + expr, itr)`            (g[idx] += expr)
        """
        if isinstance(con_or_body, _Pattern):
            con = con_or_body
            it = Iterator(itr)
            res = body(G.DataSource(it.layout))
            idx, expr = res
        else:
            it = Iterator(body)
            res = con_or_body(G.DataSource(it.layout))
            assert isinstance(res, _AugPair), "two-argument form requires `g[idx] + expr`"
            con, idx, expr = res.con, res.idx, res.expr
        if not isinstance(expr, G.AbstractNode):
            expr = G.Null(expr)
        idxs = idx if isinstance(idx, tuple) else (idx,)
        base = con if con.kind == KIND_CON else con.base
        # nlp.jl:1726-1727: dims = size(base.itr)
        dims = base.itr.shape
        self.nconaug += len(it)
        return self._push(_Pattern(KIND_AUG, expr, it, base=base, idx=idxs, dims=dims))

    # -- meta ----------------------------------------------------------------
    def _cat(self, parts):
        return np.concatenate(parts) if parts else np.zeros(0)

    def meta(self):
        return dict(
            nvar=self.nvar, ncon=self.ncon, npar=self.npar, minimize=self.minimize,
            x0=self._cat(self.x0), lvar=self._cat(self.lvar), uvar=self._cat(self.uvar),
            y0=self._cat(self.y0), lcon=self._cat(self.lcon), ucon=self._cat(self.ucon),
            theta=self._cat(self.theta),
        )

    # -- IR --------------------------------------------------------------------
    def to_ir(self):
        """Serialise to the int64 word stream of include/exa_b200.h §IR.

        Returns `(ir_bytes, data_buffers)`; `data_buffers[k]` is the AoS numpy array
        referenced as data buffer `k`.
        """
        words: list[int] = [IR_MAGIC, IR_VERSION, self.nvar, self.npar, len(self.patterns), 0]
        bufs: list[np.ndarray] = []
        for p in self.patterns:
            words += _emit_pattern(p, bufs)
        words[5] = len(bufs)
        return struct.pack(f"<{len(words)}q", *words), bufs


def _f64_bits(v: float) -> int:
    return struct.unpack("<q", struct.pack("<d", float(v)))[0]


class _NodeTable:
    def __init__(self, itr: Iterator):
        self.itr = itr
        self.rows: list[tuple[int, int, int, int]] = []
        self.memo: dict = {}
        self.fields: list[tuple[int, int]] = []  # (byte offset, type)
        self.field_of: dict = {}

    def _field(self, path):
        key = tuple(path)
        if key in self.field_of:
            return self.field_of[key]
        dt, off = self.itr.array.dtype, 0
        for f in path:
            assert dt.names is not None, f"data access path {path} does not exist"
            name = dt.names[f - 1] if G._is_int(f) else f
            sub, o = dt.fields[name][:2]
            dt, off = sub, off + o
        assert dt in _FT_OF, f"unsupported field type {dt} at {path}"
        self.fields.append((off, _FT_OF[dt]))
        self.field_of[key] = len(self.fields) - 1
        return self.field_of[key]

    def _push(self, key, row):
        if key in self.memo:
            return self.memo[key]
        self.rows.append(row)
        self.memo[key] = len(self.rows) - 1
        return self.memo[key]

    def emit(self, n):
        if isinstance(n, G.Val):
            return self._push(n.key(), (T_VAL, 0, 0, n.value))
        if not isinstance(n, G.AbstractNode):
            if G._is_int(n):
                return self._push(("i", int(n)), (T_CONST_I, 0, 0, int(n)))
            return self._push(("f", _f64_bits(n)), (T_CONST_F, 0, 0, _f64_bits(n)))
        k = n.key()
        if k in self.memo:
            return self.memo[k]
        if isinstance(n, G.Constant):
            v = n.value
            row = (T_CONST_I, 0, 0, v) if G._is_int(v) else (T_CONST_F, 0, 0, _f64_bits(v))
        elif isinstance(n, G.Null):
            row = (T_NULL, 0, 0, _f64_bits(0.0 if n.value is None else n.value))
        elif isinstance(n, G.DataSource):
            if self.itr.range is not None:
                row = (T_DATA_SELF, 0, 0, 0)
            else:
                assert self.itr.array.dtype.names is None, "a struct data point is not a scalar"
                if () not in self.field_of:
                    self.fields.append((0, _FT_OF[self.itr.array.dtype]))
                    self.field_of[()] = len(self.fields) - 1
                row = (T_DATA_FIELD, self.field_of[()], 0, 0)
        elif isinstance(n, G.DataIndexed):
            row = (T_DATA_FIELD, self._field(n.path()), 0, 0)
        elif isinstance(n, G.Var):
            row = (T_VAR, self.emit(n.i), 0, 0)
        elif isinstance(n, G.ParameterNode):
            row = (T_PAR, self.emit(n.i), 0, 0)
        elif isinstance(n, G.Node1):
            row = (T_OP1, self.emit(n.inner), 0, G.OP1_CODE[n.op])
        elif isinstance(n, G.Node2):
            a, b = self.emit(n.inner1), self.emit(n.inner2)
            row = (T_OP2, a, b, G.OP2_CODE[n.op])
        else:  # pragma: no cover
            raise TypeError(f"cannot serialise {type(n)}")
        return self._push(k, row)


def _emit_pattern(p: _Pattern, bufs):
    it = p.itr
    tab = _NodeTable(it)
    root = tab.emit(p.tree)  # a plain Real body (SIMDFunction{<:Real}) becomes a CONST node
    idx_roots = [tab.emit(i) for i in p.idx]
    w = [p.kind, p.nitr]
    if it.range is not None:
        w += [ITR_RANGE, it.range.start, -1, 0]
    else:
        bufs.append(it.array)
        w += [ITR_AOS, 0, len(bufs) - 1, it.array.dtype.itemsize]
    w.append(len(tab.fields))
    for off, ft in tab.fields:
        w += [off, ft]
    w += [-1, -1, -1]  # o0, o1, o2: assigned behind the ABI by the counter rules
    w.append(p.base.index if p.base is not None else -1)
    w.append(len(idx_roots))
    w += idx_roots
    w += [int(d) for d in p.dims][: len(idx_roots)] + [1] * max(0, len(idx_roots) - len(p.dims))
    w.append(len(tab.rows))
    for r in tab.rows:
        w += list(r)
    w.append(root)
    w += [0, 0]  # ncomp1, ncomp2 (not supplied: recomputed by the probe)
    return w


def window_core(core: ExaCore, windows):
    """The same model with every pattern's iterator restricted to the points `[a_k, b_k)` (0-based, one pair per
    pattern, in add order); variables, parameters and expression trees are shared.  Because a pattern's slots depend
    on its own data point only, callback outputs of the windowed model are the corresponding slices of the full
    model's -- the size-independent property bench.py and the full-size tests use to check a huge model against the
    oracle on a sample (`window_slices` gives the index maps).  Models with constraint augmentations are not supported
    (their row indices refer to the base constraint's full dims)."""
    w = ExaCore(minimize=core.minimize, name=core.name)
    w.nvar, w.npar = core.nvar, core.npar
    w.x0, w.lvar, w.uvar, w.theta = core.x0, core.lvar, core.uvar, core.theta
    w.vars, w.pars = core.vars, core.pars
    assert len(windows) == len(core.patterns)
    for p, (a, b) in zip(core.patterns, windows):
        assert p.kind != KIND_AUG, "window_core: augmentation patterns are not supported"
        a, b = max(0, int(a)), min(p.nitr, int(b))
        b = max(a, b)
        if p.itr.range is not None:
            it = Iterator(range(p.itr.range.start + a, p.itr.range.start + b))
        else:
            it = Iterator(np.ascontiguousarray(p.itr.array[a:b]))
        q = _Pattern(p.kind, p.tree, it, size=it.sizes)
        if p.kind == KIND_OBJ:
            w.nobj += len(it)
        else:
            q.offset = w.ncon
            w.ncon += len(it)
            z = np.zeros(len(it))
            w.y0.append(z); w.lcon.append(z); w.ucon.append(z)
        w._push(q)
    return w


def window_slices(full_info, win_info, windows):
    """Index maps between a full model and its `window_core`: for each pattern k with window [a, b) returns a dict with
    `rows` / `jac` / `hess` = (slice in the full model's c / jac / hess vector, slice in the windowed model's).
    `full_info[k]`, `win_info[k]` are `Plan.pattern_info(k)` dicts of the two models."""
    out = []
    for f, s, (a, b) in zip(full_info, win_info, windows):
        a = max(0, int(a)); b = max(a, min(f["nitr"], int(b)))
        n = b - a
        d = {"n": n}
        d["hess"] = (slice(f["o2"] + f["o2step"] * a, f["o2"] + f["o2step"] * b), slice(s["o2"], s["o2"] + s["o2step"] * n))
        if f["kind"] == KIND_CON:
            d["rows"] = (slice(f["o0"] + a, f["o0"] + b), slice(s["o0"], s["o0"] + n))
            d["jac"] = (slice(f["o1"] + f["o1step"] * a, f["o1"] + f["o1step"] * b), slice(s["o1"], s["o1"] + s["o1step"] * n))
        out.append(d)
    return out
