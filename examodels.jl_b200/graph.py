"""Expression-graph front end: the host-side mirror of the reference's pattern layer.

This is the producer of what the B200 evaluator consumes.  It mirrors the tree
SHAPES the reference builds, because the partially-compressed sparsity layout
(slot order, dedupe classes) is a function of tree shape and of the structural
identity (Julia `===`) of variable-index expressions:

* node kinds          -> /root/reference/src/graph.jl:37-230  (Null, Constant, Var,
                         ParameterNode, DataSource, DataIndexed, Node1, Node2)
* operator dispatch   -> /root/reference/src/register.jl:56-153 (node OP node,
                         node OP Real, Real OP node, Constant folding)
* literal powers      -> /root/reference/src/specialization.jl:193-202
                         (x^1 -> x, x^2 -> abs2(x), x^p -> Node2(^, x, Val{p}))
* Constant algebra    -> /root/reference/src/specialization.jl:311-339
* sum / prod          -> /root/reference/src/graph.jl:520-567 (left fold of + / *)

Nothing here evaluates anything numerically; evaluation lives behind the C ABI
(CUDA) and, for tests only, in oracle/.
"""
from __future__ import annotations

import math
import numbers

import numpy as np

# ---------------------------------------------------------------------------
# op tables (names follow /root/reference/src/functionlist.jl:6-81)
# ---------------------------------------------------------------------------
UNIVARIATE = [
    "+", "-", "inv", "sqrt", "cbrt", "abs", "abs2", "sign", "exp", "exp2", "exp10",
    "expm1", "log", "log2", "log1p", "log10", "sin", "cos", "tan", "asin", "acos",
    "atan", "acot", "csc", "sec", "cot", "sinh", "cosh", "tanh", "asinh", "acosh",
    "csch", "sech", "coth", "sind", "cosd", "tand", "cscd", "secd", "cotd", "atand",
    "acotd", "sinpi", "cospi", "sinc", "deg2rad", "rad2deg", "signbit", "floor",
    "ceil", "atanh", "acoth",
    # SpecialFunctions extension, order of /root/reference/ext/functionlist.jl:6-104
    "erf", "erfc", "erfi", "erfcx", "digamma", "trigamma", "invdigamma", "gamma", "airyai", "airybi",
    "airyaiprime", "airybiprime", "besselj0", "bessely0", "besselj1", "bessely1", "dawson", "erfinv", "erfcinv",
]
SPECIAL_UNIVARIATE = UNIVARIATE[UNIVARIATE.index("erf"):]
BIVARIATE = ["+", "-", "*", "/", "^", "atan", "hypot", "max", "min",
             "beta", "logbeta"]  # ext/functionlist.jl:111-126
SPECIAL_BIVARIATE = ["beta", "logbeta"]
OP1_CODE = {n: i for i, n in enumerate(UNIVARIATE)}
OP2_CODE = {n: i for i, n in enumerate(BIVARIATE)}


def _is_real(v) -> bool:
    return isinstance(v, (numbers.Real, np.integer, np.floating)) and not isinstance(v, bool)


def _is_int(v) -> bool:
    return isinstance(v, (numbers.Integral, np.integer)) and not isinstance(v, bool)


def _norm_real(v):
    """Python/numpy scalar -> plain int or float (Julia Int / Float64)."""
    return int(v) if _is_int(v) else float(v)


class AbstractNode:
    """Base of all graph nodes (graph.jl:11).  Structural equality == Julia `===`."""

    __slots__ = ()

    def key(self):  # pragma: no cover - overridden
        raise NotImplementedError

    def __eq__(self, other):
        return isinstance(other, AbstractNode) and self.key() == other.key()

    def __hash__(self):
        return hash(self.key())

    # --- binary operators: register.jl:126-153 -------------------------------
    def __add__(self, o):
        return _op2("+", self, o)

    def __radd__(self, o):
        return _op2("+", o, self)

    def __sub__(self, o):
        return _op2("-", self, o)

    def __rsub__(self, o):
        return _op2("-", o, self)

    def __mul__(self, o):
        return _op2("*", self, o)

    def __rmul__(self, o):
        return _op2("*", o, self)

    def __truediv__(self, o):
        return _op2("/", self, o)

    def __rtruediv__(self, o):
        return _op2("/", o, self)

    def __pow__(self, o):
        # A Python int exponent plays the role of a Julia literal exponent
        # (Base.literal_pow, specialization.jl:193-199).
        if _is_int(o):
            return _pow_val(self, int(o))
        return _op2("^", self, o)

    def __rpow__(self, o):
        return _op2("^", o, self)

    def __neg__(self):
        return _op1("-", self)

    def __pos__(self):
        return _op1("+", self)


class Null(AbstractNode):
    """graph.jl:37-40 — run-time scalar leaf; yields an AdjointNull (no slot)."""

    __slots__ = ("value",)

    def __init__(self, value=None):
        self.value = None if value is None else _norm_real(value)

    def key(self):
        return ("Null", type(self.value).__name__, self.value)


class Constant(AbstractNode):
    """graph.jl:89-91 — compile-time constant (value in the type)."""

    __slots__ = ("value",)

    def __init__(self, value):
        self.value = _norm_real(value)

    def key(self):
        return ("Constant", type(self.value).__name__, self.value)


class Val:
    """Julia `Val{V}()` exponent child of Node2(^, x, Val{V}()) (specialization.jl:199-202)."""

    __slots__ = ("value",)

    def __init__(self, value):
        self.value = int(value)

    def key(self):
        return ("Val", self.value)

    def __eq__(self, o):
        return isinstance(o, Val) and o.value == self.value

    def __hash__(self):
        return hash(self.key())


class DataSource(AbstractNode):
    """graph.jl:183 — the data point itself (`i` of `for i in itr`)."""

    __slots__ = ("layout",)

    def __init__(self, layout=None):
        self.layout = layout  # numpy dtype (structured) or None for a scalar int/float

    def key(self):
        return ("DataSource",)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return DataIndexed(self, name)

    def __getitem__(self, j):
        return DataIndexed(self, j)

    def __iter__(self):
        return _iter_fields(self)


class DataIndexed(AbstractNode):
    """graph.jl:194-199 — field `J` of `inner` (Symbol or 1-based position)."""

    __slots__ = ("inner", "field")

    def __init__(self, inner, field):
        object.__setattr__(self, "inner", inner)
        object.__setattr__(self, "field", field)

    def key(self):
        return ("DataIndexed", self.inner.key(), self.field)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return DataIndexed(self, name)

    def __getitem__(self, j):
        return DataIndexed(self, j)

    def __iter__(self):
        return _iter_fields(self)

    def path(self):
        p, n = [], self
        while isinstance(n, DataIndexed):
            p.append(n.field)
            n = n.inner
        return tuple(reversed(p))


def _sub_dtype(node):
    """numpy dtype reached by following node's access path from the element dtype."""
    if isinstance(node, DataSource):
        return node.layout
    inner = _sub_dtype(node.inner)
    if inner is None or inner.names is None:
        return None
    f = node.field
    name = inner.names[f - 1] if _is_int(f) else f  # positions are 1-based like Julia
    return inner.fields[name][0]


def _iter_fields(node):
    """Tuple destructuring `(i, j) = d` (graph.jl:289,293: indexed_iterate)."""
    dt = _sub_dtype(node)
    if dt is None or dt.names is None:
        raise TypeError("cannot destructure a scalar data point")
    return iter([DataIndexed(node, k + 1) for k in range(len(dt.names))])


class Var(AbstractNode):
    """graph.jl:138-140 — x[i]; `i` is an int or an index-expression node."""

    __slots__ = ("i",)

    def __init__(self, i):
        self.i = int(i) if _is_int(i) else i

    def key(self):
        return ("Var", _child_key(self.i))


class ParameterNode(AbstractNode):
    """graph.jl:143-145 — θ[i]."""

    __slots__ = ("i",)

    def __init__(self, i):
        self.i = int(i) if _is_int(i) else i

    def key(self):
        return ("Par", _child_key(self.i))


class Node1(AbstractNode):
    """graph.jl:210-212."""

    __slots__ = ("op", "inner")

    def __init__(self, op, inner):
        assert op in OP1_CODE, op
        self.op, self.inner = op, inner

    def key(self):
        return ("Node1", self.op, _child_key(self.inner))


class Node2(AbstractNode):
    """graph.jl:227-230 — children may be nodes, stored Reals, or Val exponents."""

    __slots__ = ("op", "inner1", "inner2")

    def __init__(self, op, inner1, inner2):
        assert op in OP2_CODE, op
        self.op = op
        self.inner1 = _norm_real(inner1) if _is_real(inner1) else inner1
        self.inner2 = _norm_real(inner2) if _is_real(inner2) else inner2

    def key(self):
        return ("Node2", self.op, _child_key(self.inner1), _child_key(self.inner2))


def _child_key(c):
    if isinstance(c, (AbstractNode, Val)):
        return c.key()
    return (type(c).__name__, c)  # int vs float are different Julia types


# ---------------------------------------------------------------------------
# dispatch (register.jl + specialization.jl)
# ---------------------------------------------------------------------------
_PYF1 = {
    "+": lambda x: +x, "-": lambda x: -x, "inv": lambda x: 1.0 / x, "sqrt": math.sqrt,
    "cbrt": lambda x: math.copysign(abs(x) ** (1.0 / 3.0), x), "abs": abs,
    "abs2": lambda x: x * x, "sign": lambda x: (x > 0) - (x < 0), "exp": math.exp,
    "exp2": lambda x: 2.0 ** x, "exp10": lambda x: 10.0 ** x, "expm1": math.expm1,
    "log": math.log, "log2": math.log2, "log1p": math.log1p, "log10": math.log10,
    "sin": math.sin, "cos": math.cos, "tan": math.tan, "asin": math.asin,
    "acos": math.acos, "atan": math.atan, "sinh": math.sinh, "cosh": math.cosh,
    "tanh": math.tanh, "asinh": math.asinh, "acosh": math.acosh, "atanh": math.atanh,
    "floor": math.floor, "ceil": math.ceil, "deg2rad": math.radians, "rad2deg": math.degrees,
}
_PYF2 = {
    "+": lambda a, b: a + b, "-": lambda a, b: a - b, "*": lambda a, b: a * b,
    "/": lambda a, b: a / b, "^": lambda a, b: a ** b, "atan": math.atan2,
    "hypot": math.hypot, "max": max, "min": min,
}


def _special_folds():
    """Constant folding of the SpecialFunctions ops on the host (a Real argument is evaluated eagerly, register.jl:70)."""
    import scipy.special as S
    one = {"erf": S.erf, "erfc": S.erfc, "erfi": S.erfi, "erfcx": S.erfcx, "digamma": S.digamma,
           "trigamma": lambda x: S.polygamma(1, x), "gamma": S.gamma, "airyai": lambda x: S.airy(x)[0],
           "airybi": lambda x: S.airy(x)[2], "airyaiprime": lambda x: S.airy(x)[1], "airybiprime": lambda x: S.airy(x)[3],
           "besselj0": S.j0, "bessely0": S.y0, "besselj1": S.j1, "bessely1": S.y1, "dawson": S.dawsn,
           "erfinv": S.erfinv, "erfcinv": S.erfcinv}
    return ({k: (lambda x, f=f: float(f(x))) for k, f in one.items()},
            {"beta": lambda a, b: float(S.beta(a, b)), "logbeta": lambda a, b: float(S.betaln(a, b))})


def _fold1(op, v):
    f = _PYF1.get(op)
    if f is None and op in SPECIAL_UNIVARIATE and op != "invdigamma":
        _PYF1.update(_special_folds()[0])
        f = _PYF1.get(op)
    if f is None:
        raise NotImplementedError(f"constant folding of {op} is not available on the host")
    return f(v)


def _op1(op, a):
    if isinstance(a, Constant):  # register.jl:63
        return Constant(_fold1(op, a.value))
    if isinstance(a, AbstractNode):  # register.jl:60
        return Node1(op, a)
    return _fold1(op, a)


def _iszero(c):
    return isinstance(c, Constant) and c.value == 0


def _isone(c):
    return isinstance(c, Constant) and c.value == 1


def _op2(op, a, b):
    an, bn = isinstance(a, AbstractNode), isinstance(b, AbstractNode)
    if isinstance(a, Constant) and isinstance(b, Constant):  # register.jl:135
        if op in SPECIAL_BIVARIATE and op not in _PYF2:
            _PYF2.update(_special_folds()[1])
        return Constant(_PYF2[op](a.value, b.value))
    # Constant algebra, specialization.jl:311-339 (a Constant meeting a node)
    if an and bn:
        if op == "+":
            if _iszero(a):
                return b
            if _iszero(b):
                return a
        elif op == "-":
            if _iszero(a):
                return _op1("-", b)
            if _iszero(b):
                return a
        elif op == "*":
            if _iszero(a) or _iszero(b):
                return Constant(0)
            if _isone(a):
                return b
            if _isone(b):
                return a
        elif op == "/":
            if _iszero(a):
                return Constant(0)
            if _isone(b):
                return a
            if _isone(a):
                return _op1("inv", b)
        elif op == "^":
            if _iszero(a):  # graph.jl:93
                return Constant(0)
            if _isone(a):  # graph.jl:94
                return Constant(1)
            if _iszero(b):
                return Constant(1)
            if _isone(b):
                return a
            if isinstance(b, Constant) and b.value == -1:
                return _op1("inv", a)
            if isinstance(b, Constant) and b.value == 2:
                return _op1("abs2", a)
    if an or bn:
        if not an and not _is_real(a):
            return NotImplemented
        if not bn and not _is_real(b) and not isinstance(b, Val):
            return NotImplemented
        return Node2(op, a, b)
    if op in SPECIAL_BIVARIATE and op not in _PYF2:
        _PYF2.update(_special_folds()[1])
    return _PYF2[op](a, b)


def _pow_val(d1, p: int):
    """specialization.jl:197-199."""
    if p == 1:
        return d1
    if p == 2:
        return Node1("abs2", d1)
    return Node2("^", d1, Val(p))


def pow_runtime(d1, p):
    """`x^n` with a NON-literal exponent: Node2(^, x, n) (specialization.jl:195-196)."""
    return _op2("^", d1, p)


def exa_sum(f, itr):
    """specialization.jl:253-254,282-290 + graph.jl:549-550: evaluated as a left fold of `+`.

    An empty sum is `AdjointNull(0)` (graph.jl:547-548) -> Null(0).
    """
    terms = [f(k) for k in itr]
    if not terms:
        return Null(0)
    acc = terms[0]
    for t in terms[1:]:
        acc = acc + t
    return acc


def exa_prod(f, itr):
    """graph.jl:552-555: left fold of `*`; the empty product is AdjointNull(1)."""
    terms = [f(k) for k in itr]
    if not terms:
        return Null(1)
    acc = terms[0]
    for t in terms[1:]:
        acc = acc * t
    return acc


def _make_unary(name):
    def fn(x):
        return _op1(name, x)

    fn.__name__ = name
    fn.__doc__ = f"`{name}` registered as a univariate op (functionlist.jl:6-60)."
    return fn


_g = globals()
for _n in UNIVARIATE:
    if _n in ("+", "-"):
        continue
    _g[_n if _n not in ("abs",) else "abs_"] = _make_unary(_n)
abs_ = _g["abs_"]


def beta(a, b):
    """`SpecialFunctions.beta` (ext/functionlist.jl:111-118)."""
    return _op2("beta", a, b)


def logbeta(a, b):
    """`SpecialFunctions.logbeta` (ext/functionlist.jl:119-126)."""
    return _op2("logbeta", a, b)


def atan2(a, b):
    """Two-argument `atan(y, x)` (functionlist.jl:77)."""
    return _op2("atan", a, b)


def hypot(a, b):
    return _op2("hypot", a, b)


def max_(a, b):
    return _op2("max", a, b)


def min_(a, b):
    return _op2("min", a, b)
