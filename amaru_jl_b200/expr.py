"""Filter / value expressions in x, y, z, t with the reference's tolerance semantics.

The reference evaluates ``:(x==0 && y==1)``-style expressions with tolerant comparisons
(``src/tools/expr.jl:95-126``: ``arith_tol = 1e-6``; ``==`` is ``|a-b| < tol``, ``>=`` is ``a > b - tol`` ...).
Here the same expressions are given as strings (Julia's ``&&``, ``||``, ``^`` are accepted) or as Python
callables ``f(x, y, z)``; evaluation is vectorised over numpy arrays.
"""
from __future__ import annotations

import ast

import numpy as np

ARITH_TOL = 1e-6

_FUNCS = {"abs": np.abs, "sin": np.sin, "cos": np.cos, "tan": np.tan, "exp": np.exp, "log": np.log,
          "log10": np.log10, "sqrt": np.sqrt, "max": np.maximum, "min": np.minimum}


def _cmp(op, a, b):
    if isinstance(op, ast.Eq):
        return np.abs(a - b) < ARITH_TOL
    if isinstance(op, ast.NotEq):
        return np.abs(a - b) >= ARITH_TOL
    if isinstance(op, ast.Gt):
        return a > b + ARITH_TOL
    if isinstance(op, ast.Lt):
        return a < b - ARITH_TOL
    if isinstance(op, ast.GtE):
        return a > b - ARITH_TOL
    if isinstance(op, ast.LtE):
        return a < b + ARITH_TOL
    raise ValueError("comparison not allowed in this context")


def _ev(node, env):
    if isinstance(node, ast.Expression):
        return _ev(node.body, env)
    if isinstance(node, ast.Constant):
        return node.value
    if isinstance(node, ast.Name):
        if node.id == "pi":
            return np.pi
        if node.id not in env:
            raise ValueError(f"variable {node.id} not defined for this context")
        return env[node.id]
    if isinstance(node, ast.BoolOp):
        vals = [_ev(v, env) for v in node.values]
        out = vals[0]
        for v in vals[1:]:
            out = np.logical_and(out, v) if isinstance(node.op, ast.And) else np.logical_or(out, v)
        return out
    if isinstance(node, ast.UnaryOp):
        v = _ev(node.operand, env)
        if isinstance(node.op, ast.USub):
            return -v
        if isinstance(node.op, ast.UAdd):
            return v
        if isinstance(node.op, ast.Not):
            return np.logical_not(v)
    if isinstance(node, ast.BinOp):
        a, b = _ev(node.left, env), _ev(node.right, env)
        if isinstance(node.op, ast.Add):
            return a + b
        if isinstance(node.op, ast.Sub):
            return a - b
        if isinstance(node.op, ast.Mult):
            return a * b
        if isinstance(node.op, ast.Div):
            return a / b
        if isinstance(node.op, ast.Pow):
            return a ** b
    if isinstance(node, ast.Compare):
        left = _ev(node.left, env)
        out = True
        for op, comp in zip(node.ops, node.comparators):
            right = _ev(comp, env)
            out = np.logical_and(out, _cmp(op, left, right))
            left = right
        return out
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS:
        return _FUNCS[node.func.id](*[_ev(a, env) for a in node.args])
    raise ValueError(f"operation not allowed in this context: {ast.dump(node)[:60]}")


def _parse(expr: str):
    s = expr.strip()
    if s.startswith(":(") and s.endswith(")"):
        s = s[2:-1]
    s = s.replace("&&", " and ").replace("||", " or ").replace("^", "**")
    return ast.parse(s, mode="eval")


def evaluate(expr, **vars):
    """Evaluate a number, callable or expression string over (arrays of) x, y, z, t."""
    if callable(expr):
        return expr(vars.get("x"), vars.get("y"), vars.get("z"))
    if isinstance(expr, (int, float, np.floating, np.integer)):
        return float(expr)
    return _ev(_parse(expr), vars)


def select(expr, coords):
    """Boolean mask of the points (n,3) that satisfy a filter expression."""
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    m = evaluate(expr, x=x, y=y, z=z)
    if np.ndim(m) == 0:
        m = np.full(coords.shape[0], bool(m))
    return np.asarray(m, dtype=bool)
