"""Symbolic linear expressions for heat-balance equations.

Interface mirror of the reference's ``heatsim2/expression.py`` (public names
``linear_expression`` :15, ``group`` :703, ``eliminate_groups`` :712,
``no_groups`` :718, ``subst`` :724, ``crank_subst_in_groups`` :753, methods
``fullreduce`` :537, ``dictform`` :561, ``exprstr`` :646) so that boundary
plug-ins written for heatsim2 (``qz/qy/qx`` callables that combine their
operands with ``+ - * / []`` and ``group``) run unchanged.

The implementation is independent of the reference: the reference keeps an RPN
command list and rewrites it in place; here an expression is an immutable
nested-tuple syntax tree and reduction is a recursive expansion into a sum of
monomials.  The two agree on every ``dictform()`` up to the rounding of the
coefficient products (checked against the built reference in
tests/test_expression.py).

Tree nodes (plain tuples, hashable):
    ("c", value)                      constant (number, or ndarray for tensors)
    ("v", name, coef, i0, i1)         variable * coef, optional [i0,i1] index
    ("+", a, b)  ("*", a, b)  ("/", a, b)
    ("g", a)                          group: opaque to reduction until
                                      crank_subst_in_groups / eliminate_groups
"""
import numbers

import numpy as np


def _is_number(x):
    return isinstance(x, numbers.Number)


class linear_expression(object):
    """A linear (after parameter substitution) expression in named variables."""

    __slots__ = ("_n",)

    def __init__(self, *args):
        if len(args) == 0:
            # the reference's "empty, non-final" expression; value 0 here
            self._n = ("c", 0.0)
        elif len(args) == 1:
            value = args[0]
            if isinstance(value, linear_expression):
                self._n = value._n
            elif isinstance(value, tuple) and value and value[0] in ("c", "v", "+", "*", "/", "g"):
                self._n = value
            elif _is_number(value):
                self._n = ("c", value)
            elif isinstance(value, str):
                self._n = ("v", value, 1.0, None, None)
            else:
                raise ValueError("Unknown argument type")
        else:
            raise ValueError("Too many arguments (%d)" % (len(args)))

    # ------------------------------------------------------------------ algebra
    @staticmethod
    def _node(x):
        return linear_expression(x)._n

    def __add__(self, other):
        return linear_expression(("+", self._n, self._node(other)))

    def __radd__(self, other):
        return linear_expression(("+", self._node(other), self._n))

    def __sub__(self, other):
        return linear_expression(("+", self._n, ("*", self._node(other), ("c", -1))))

    def __rsub__(self, other):
        return linear_expression(("+", self._node(other), ("*", self._n, ("c", -1))))

    def __neg__(self):
        return linear_expression(("*", self._n, ("c", -1)))

    def __mul__(self, other):
        return linear_expression(("*", self._n, self._node(other)))

    def __rmul__(self, other):
        return linear_expression(("*", self._node(other), self._n))

    def __truediv__(self, other):
        return linear_expression(("/", self._n, self._node(other)))

    def __rtruediv__(self, other):
        return linear_expression(("/", self._node(other), self._n))

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __getitem__(self, index):
        """Index every not-yet-indexed variable by a 2-tuple (tensor element)."""
        assert len(index) == 2

        def f(node):
            if node[0] == "v" and node[3] is None and node[4] is None:
                return ("v", node[1], node[2], index[0], index[1])
            return node
        return linear_expression(_map_leaves(self._n, f))

    # ---------------------------------------------------------------- identity
    def __hash__(self):
        return hash(self._n)

    def __eq__(self, other):
        return isinstance(other, linear_expression) and self._n == other._n

    def __ne__(self, other):
        return not self.__eq__(other)

    # --------------------------------------------------------------- traversal
    def map_vars(self, fn):
        """Return a copy with every variable leaf ``("v",name,coef,i0,i1)``
        replaced by ``fn(name, coef, i0, i1)`` (a node tuple)."""
        return linear_expression(_map_leaves(
            self._n, lambda nd: fn(nd[1], nd[2], nd[3], nd[4]) if nd[0] == "v" else nd))

    def variables(self):
        out = []
        _collect_vars(self._n, out)
        return out

    # --------------------------------------------------------------- reduction
    def terms(self):
        """Expand into a list of ``(coef, factors)`` monomials; ``factors`` is
        a tuple of opaque atoms (variables / groups).  Zero monomials are
        dropped - this is what lets an axis-aligned tensor conductivity shed
        its cross-derivative groups (reference expression.py:299-330)."""
        return _expand(self._n)

    def fullreduce(self):
        node = None
        for coef, factors in self.terms():
            if len(factors) == 0:
                t = ("c", coef)
            else:
                t = None
                for f in factors:
                    t = f if t is None else ("*", t, f)
                if len(factors) == 1 and factors[0][0] == "v":
                    f = factors[0]
                    t = ("v", f[1], coef * f[2], f[3], f[4])
                else:
                    t = ("*", ("c", coef), t)
            node = t if node is None else ("+", node, t)
        if node is None:
            node = ("c", 0.0)
        return linear_expression(node)

    def dictform(self):
        """``{variable name: coefficient}``; the constant term, when present,
        is under the key ``""`` (reference expression.py:561-601)."""
        out = {}
        for coef, factors in self.terms():
            if len(factors) == 0:
                out[""] = out.get("", 0.0) + coef
                continue
            if len(factors) > 1 or factors[0][0] != "v":
                raise ValueError("Expression %s cannot be reduced to dictionary form" % self.exprstr())
            f = factors[0]
            if f[3] is not None or f[4] is not None:
                raise ValueError("Expressions with indexed operands such as %s cannot be reduced to dictionary form" % f[1])
            out[f[1]] = out.get(f[1], 0.0) + coef * f[2]
        return out

    # ----------------------------------------------------------------- display
    def exprstr(self):
        return _to_str(self._n)

    def __str__(self):
        return "linear_expression( %s )" % _to_str(self._n)

    __repr__ = __str__


# ---------------------------------------------------------------------- helpers
def _map_leaves(node, f):
    k = node[0]
    if k in ("c", "v"):
        return f(node)
    if k == "g":
        return ("g", _map_leaves(node[1], f))
    return (k, _map_leaves(node[1], f), _map_leaves(node[2], f))


def _collect_vars(node, out):
    k = node[0]
    if k == "v":
        out.append(node[1])
    elif k == "g":
        _collect_vars(node[1], out)
    elif k != "c":
        _collect_vars(node[1], out)
        _collect_vars(node[2], out)


def _iszero(c):
    if isinstance(c, np.ndarray):
        return False
    return c == 0.0


def _expand(node):
    k = node[0]
    if k == "c":
        return [] if _iszero(node[1]) else [(node[1], ())]
    if k == "v":
        if _iszero(node[2]):
            return []
        return [(1.0, (node,))]
    if k == "g":
        return [(1.0, (node,))]
    a = _expand(node[1])
    if k == "+":
        return a + _expand(node[2])
    b = _expand(node[2])
    if k == "*":
        out = []
        for ca, fa in a:
            for cb, fb in b:
                out.append(_mono(ca * cb, fa + fb))
        return [t for t in out if not _iszero(t[0])]
    if k == "/":
        if len(b) != 1 or len(b[0][1]) != 0:
            # allow a single pure-coefficient variable? no: divisor must be numeric
            raise ValueError("division by a non-constant expression")
        d = b[0][0]
        return [(float(ca) / d if _is_number(ca) else ca / d, fa) for ca, fa in a]
    raise ValueError("Unknown node %r" % (k,))


def _mono(coef, factors):
    """Pull variable coefficients into the monomial coefficient."""
    if len(factors) == 1 and factors[0][0] == "v" and factors[0][2] != 1.0:
        f = factors[0]
        return (coef * f[2], (("v", f[1], 1.0, f[3], f[4]),))
    if len(factors) > 1:
        fs = []
        for f in factors:
            if f[0] == "v" and f[2] != 1.0:
                coef = coef * f[2]
                f = ("v", f[1], 1.0, f[3], f[4])
            fs.append(f)
        return (coef, tuple(fs))
    return (coef, factors)


def _to_str(node):
    k = node[0]
    if k == "c":
        return str(node[1]) if isinstance(node[1], np.ndarray) else "%g" % node[1]
    if k == "v":
        if node[3] is None and node[4] is None:
            return "%g*%s" % (node[2], node[1])
        return "%g*%s[%d,%d]" % (node[2], node[1], node[3], node[4])
    if k == "g":
        return "GROUP(%s)" % _to_str(node[1])
    return "(%s %s %s)" % (_to_str(node[1]), k, _to_str(node[2]))


# ------------------------------------------------------------------ public api
def group(linear_exp):
    """Mark ``linear_exp`` as a unit that reduction must not split."""
    return linear_expression(("g", linear_expression(linear_exp)._n))


def _strip_groups(node):
    k = node[0]
    if k in ("c", "v"):
        return node
    if k == "g":
        return _strip_groups(node[1])
    return (k, _strip_groups(node[1]), _strip_groups(node[2]))


def eliminate_groups(expr):
    return linear_expression(_strip_groups(expr._n))


def _has_group(node):
    k = node[0]
    if k in ("c", "v"):
        return False
    if k == "g":
        return True
    return _has_group(node[1]) or _has_group(node[2])


def no_groups(expr):
    return not _has_group(expr._n)


def subst(expr, origvar, newvar_or_value):
    """Rename variable ``origvar`` (string argument) or replace it by a value
    (number, or a tensor that indexed occurrences pick an element from)."""
    if isinstance(newvar_or_value, str):
        def f(node):
            if node[0] == "v" and node[1] == origvar:
                return ("v", newvar_or_value, node[2], node[3], node[4])
            return node
    else:
        def f(node):
            if node[0] == "v" and node[1] == origvar:
                if node[3] is not None or node[4] is not None:
                    return ("c", newvar_or_value[node[3], node[4]] * node[2])
                return ("c", newvar_or_value * node[2])
            return node
    return linear_expression(_map_leaves(expr._n, f))


def crank_subst_in_groups(expr, groupmembers, solnum):
    """Crank-Nicolson time averaging of one sweep direction.

    Every group whose variables all belong to ``groupmembers`` is opened and
    each variable ``V`` in it becomes ``(V+'p<solnum>' + V+'m') / 2`` - the
    implicit half at the new time level of solution stage ``solnum`` and the
    explicit half at the old one (reference expression.py:753-827)."""
    members = set(groupmembers)

    def avg(node):
        if node[0] == "v":
            assert node[3] is None and node[4] is None
            return ("+", ("v", node[1] + "p" + str(solnum), node[2] / 2.0, None, None),
                    ("v", node[1] + "m", node[2] / 2.0, None, None))
        return node

    def walk(node):
        k = node[0]
        if k in ("c", "v"):
            return node
        if k == "g":
            names = []
            _collect_vars(node[1], names)
            if all(n in members for n in names):
                return _map_leaves(node[1], avg)
            return node
        return (k, walk(node[1]), walk(node[2]))

    return linear_expression(walk(expr._n))
