"""Lower the SymPy F / J expression trees of a model to CUDA device functions.

Replaces the ``lambdify`` -> NumPy step of the reference's compiler
(reference ``triflow/core/compilers.py:207-219``).  To reproduce the reference's
results to the last bit the *same printed source* that the reference executes is
used as the intermediate form: the expressions are printed by SymPy's NumPy
printer exactly as ``lambdify`` does, parsed with :mod:`ast`, and translated in
the printed association order into IEEE round-to-nearest operations without
FMA contraction (``__dadd_rn / __dmul_rn / __ddiv_rn`` on the device).

Sub-trees that depend only on ``dx``, literals and *uniform* parameters (scalar,
or one value per ensemble member) are not lowered: they are evaluated once on
the host, with NumPy and the operand types the reference sees (``dx`` a
``float64`` scalar, parameters ``float64`` arrays), and handed to the kernels as
a table of constants — ``dx**4``, ``0.5*c/dx + k/dx**2`` ....
Everything that depends on a field value, a per-node parameter array or ``x``
becomes straight-line device code with common sub-expressions shared.

``Heaviside`` lowers to the constant 1 (the reference's table entry,
``compilers.py:204-205``); ``Max/Min`` to NumPy ``maximum/minimum`` semantics.
"""

import ast
import functools
import hashlib
import inspect

import numpy as np
from sympy import lambdify

_UNARY_FUNCS = {
    "sqrt": "sqrt", "exp": "exp", "log": "log", "sin": "sin", "cos": "cos",
    "tan": "tan", "tanh": "tanh", "sinh": "sinh", "cosh": "cosh",
    "arctan": "atan", "arcsin": "asin", "arccos": "acos", "abs": "fabs",
    "absolute": "fabs", "sign": "TF_SIGN", "log10": "log10", "log2": "log2",
    "arcsinh": "asinh", "arccosh": "acosh", "arctanh": "atanh", "floor": "floor",
    "ceil": "ceil", "expm1": "expm1", "log1p": "log1p", "cbrt": "cbrt",
    "erf": "erf", "erfc": "erfc",
}
_NAMED_CONST = {"pi": np.pi, "E": np.e, "e": np.e}


def _heaviside(a, h0=None):
    # reference semantics: where(a < 0, 1, 1)  (compilers.py:204-205)
    return np.where(np.asarray(a) < 0, 1, 1)


def _host_namespace():
    ns = {name: getattr(np, name) for name in _UNARY_FUNCS if hasattr(np, name)}
    ns.update(reduce=functools.reduce, maximum=np.maximum, minimum=np.minimum,
              Heaviside=_heaviside, numpy=np, pi=np.pi, E=np.e, abs=np.abs)
    return ns


def printed_source(model, exprs):
    """The source ``lambdify`` generates for ``exprs`` (what the reference runs)."""
    stub = {"amax": None, "amin": None, "Heaviside": None}
    fn = lambdify(model._symbolic_args, expr=list(exprs), modules=[stub, "numpy"])
    return inspect.getsource(fn)


def _c_literal(v):
    v = float(v)
    if v != v:
        return "TF_NAN"
    if v in (float("inf"), float("-inf")):
        return "TF_INF" if v > 0 else "(-TF_INF)"
    s = repr(v)
    if "e" not in s and "." not in s and "n" not in s:
        s += ".0"
    return s


class _Lowering:
    """Translate the return-list of one lambdify-generated function."""

    LIT, UNI, NODE = 0, 1, 2

    def __init__(self, arg_kinds, consts):
        self.arg_kinds = arg_kinds          # name -> ("win", f, o) | ("npar", j) | ("x",) | ("uni",)
        self.consts = consts                # shared dict: source -> index
        self.lines = []
        self.cse = {}
        self.ntmp = 0
        self.n_div = self.n_divc = self.n_ops = 0

    # ---- classification
    def kind(self, node):
        k = getattr(node, "_tfk", None)
        if k is not None:
            return k
        if isinstance(node, ast.Constant):
            k = self.LIT
        elif isinstance(node, ast.Name):
            if node.id in self.arg_kinds:
                k = self.UNI if self.arg_kinds[node.id][0] == "uni" else self.NODE
            elif node.id in _NAMED_CONST:
                k = self.LIT
            else:
                raise NotImplementedError("unknown symbol %r in model expression" % node.id)
        elif isinstance(node, ast.Attribute):     # numpy.pi
            k = self.LIT
        elif isinstance(node, ast.UnaryOp):
            k = self.kind(node.operand)
        elif isinstance(node, ast.BinOp):
            k = max(self.kind(node.left), self.kind(node.right))
        elif isinstance(node, ast.Call):
            name = self._callee(node)
            if name == "Heaviside":
                k = self.LIT
            elif name == "reduce":
                k = max(self.kind(e) for e in node.args[1].elts)
            else:
                k = max(self.kind(a) for a in node.args)
        elif isinstance(node, (ast.List, ast.Tuple)):
            k = max(self.kind(e) for e in node.elts)
        else:
            raise NotImplementedError("unsupported construct in model expression: %s"
                                      % ast.dump(node)[:80])
        node._tfk = k
        return k

    @staticmethod
    def _callee(node):
        f = node.func
        return f.id if isinstance(f, ast.Name) else f.attr

    # ---- emission helpers
    def _tmp(self, expr):
        if expr in self.cse:
            return self.cse[expr]
        name = "t%d" % self.ntmp
        self.ntmp += 1
        self.lines.append("const double %s = %s;" % (name, expr))
        self.cse[expr] = name
        self.n_ops += 1
        return name

    def _const(self, node):
        src = ast.unparse(node)
        if src not in self.consts:
            self.consts[src] = len(self.consts)
        return self.consts[src]

    def _literal_value(self, node):
        if isinstance(node, ast.Call) and self._callee(node) == "Heaviside":
            return 1.0
        return float(eval(compile(ast.Expression(node), "<lit>", "eval"),
                          {"numpy": np, **_NAMED_CONST}))

    # ---- expression -> C operand
    def emit(self, node):
        k = self.kind(node)
        if k == self.LIT:
            return _c_literal(self._literal_value(node))
        if k == self.UNI:
            return "cst[%d]" % self._const(node)
        if isinstance(node, ast.Name):
            info = self.arg_kinds[node.id]
            if info[0] == "win":
                return "in.w[%d][%d]" % (info[1], info[2])
            if info[0] == "npar":
                return "in.np[%d]" % info[1]
            return "in.x"
        if isinstance(node, ast.UnaryOp):
            a = self.emit(node.operand)
            if isinstance(node.op, ast.USub):
                return self._tmp("(-%s)" % a)
            if isinstance(node.op, ast.UAdd):
                return a
            raise NotImplementedError("unary operator")
        if isinstance(node, ast.BinOp):
            return self._binop(node)
        if isinstance(node, ast.Call):
            return self._call(node)
        raise NotImplementedError(ast.dump(node)[:80])

    def _binop(self, node):
        op = node.op
        if isinstance(op, ast.Pow):
            return self._pow(node)
        a = self.emit(node.left)
        if isinstance(op, ast.Div):
            if self.kind(node.right) == self.UNI:
                self.n_divc += 1
                return self._tmp("TF_DIVC(%s, %d)" % (a, self._const(node.right)))
            self.n_div += 1
            return self._tmp("TF_DIV(%s, %s)" % (a, self.emit(node.right)))
        b = self.emit(node.right)
        name = {ast.Add: "TF_ADD", ast.Sub: "TF_SUB", ast.Mult: "TF_MUL"}.get(type(op))
        if name is None:
            raise NotImplementedError("binary operator %s" % type(op).__name__)
        return self._tmp("%s(%s, %s)" % (name, a, b))

    def _pow(self, node):
        base = self.emit(node.left)
        if self.kind(node.right) == self.LIT:
            e = self._literal_value(node.right)
            # NumPy's scalar-exponent fast paths (array ** 2 is square, etc.)
            if e == 2.0:
                return self._tmp("TF_MUL(%s, %s)" % (base, base))
            if e == 1.0:
                return base
            if e == 0.0:
                return "1.0"
            if e == -1.0:
                self.n_div += 1
                return self._tmp("TF_DIV(1.0, %s)" % base)
            if e == 0.5:
                return self._tmp("TF_SQRT(%s)" % base)
            if e == int(e) and 3 <= abs(e) <= 16:
                return self._tmp("tf_powi(%s, %d)" % (base, int(e)))
            return self._tmp("TF_POW(%s, %s)" % (base, _c_literal(e)))
        return self._tmp("TF_POW(%s, %s)" % (base, self.emit(node.right)))

    def _call(self, node):
        name = self._callee(node)
        if name == "reduce":
            which = node.args[0]
            which = which.id if isinstance(which, ast.Name) else which.attr
            fn = {"maximum": "TF_MAX", "minimum": "TF_MIN"}[which]
            elts = [self.emit(e) for e in node.args[1].elts]
            acc = elts[0]
            for e in elts[1:]:
                acc = self._tmp("%s(%s, %s)" % (fn, acc, e))
            return acc
        if name in ("maximum", "minimum"):
            fn = "TF_MAX" if name == "maximum" else "TF_MIN"
            return self._tmp("%s(%s, %s)" % (fn, self.emit(node.args[0]),
                                             self.emit(node.args[1])))
        if name in _UNARY_FUNCS and len(node.args) == 1:
            return self._tmp("%s(%s)" % (_UNARY_FUNCS[name], self.emit(node.args[0])))
        raise NotImplementedError("function %r is not supported by the CUDA compiler"
                                  % name)


class Lowered:
    """Result of :func:`lower`."""

    def uniform_table(self, dx, pars, batch):
        """Evaluate the uniform sub-expressions on the host.

        ``dx`` float64 scalar; ``pars[name]`` scalar or ``(batch,)`` array.
        Returns ``(batch, n_const_total)`` float64: constants, then (for the
        fast-division mode) their reciprocals."""
        ns = _host_namespace()
        ns["dx"] = np.float64(dx)
        for name in self.uniform_pars:
            v = np.asarray(pars[name], dtype=np.float64)
            if v.ndim == 2 and v.shape[1] == 1:           # (batch, 1): one value per member
                v = v[:, 0]
            ns[name] = np.broadcast_to(v, (batch,)).copy()
        table = np.empty((batch, 2 * max(1, self.n_const)), dtype=np.float64)
        table[:] = 1.0
        with np.errstate(all="ignore"):
            for src, j in self.consts.items():
                table[:, j] = eval(self._const_code[src], ns)
            table[:, self.n_const:2 * self.n_const] = 1.0 / table[:, :self.n_const]
        return table


def lower(model, node_pars=()):
    """Lower ``model`` for a given set of per-node (array) parameters."""
    node_pars = tuple(p for p in model._pars if p in set(node_pars))
    lo, hi = model._bounds
    p = max(-lo, hi)
    fields = list(model._dep_vars) + list(model._help_funcs)
    arg_kinds = {"x": ("x",), "dx": ("uni",)}
    for f, name in enumerate(fields):
        for o in range(-p, p + 1):
            key = name if o == 0 else "%s_%s%d" % (name, "m" if o < 0 else "p", abs(o))
            arg_kinds[key] = ("win", f, o + p)
    for name in model._pars:
        arg_kinds[name] = (("npar", node_pars.index(name)) if name in node_pars
                           else ("uni",))

    consts = {}
    out = Lowered()
    bodies = {}
    stats = {}
    uses_x = False
    for which, exprs in (("F", model.F_array.tolist()),
                         ("J", model._J_sparse_array.tolist())):
        src = printed_source(model, exprs)
        tree = ast.parse(src)
        ret = tree.body[0].body[-1].value
        assert isinstance(ret, ast.List)
        low = _Lowering(arg_kinds, consts)
        results = [low.emit(e) for e in ret.elts]
        lines = list(low.lines)
        for i, r in enumerate(results):
            lines.append("out[%d] = %s;" % (i, r))
        bodies[which] = lines
        stats[which] = dict(ops=low.n_ops, div=low.n_div, divc=low.n_divc)
        uses_x = uses_x or any("in.x" in ln for ln in lines)
        setattr(out, which + "_printed", src)

    # cheaper form of F for the solver kernels of nonlinear models: linear stencil part with
    # host-evaluated coefficients + the expanded remainder (adopted when it saves a third)
    out.f_split = False
    const_j = stats["J"]["ops"] == 0 and all("in." not in ln for ln in bodies["J"])   # -> linear form
    split = (_split_linear_part(model, fields, len(model._dep_vars), p)
             if not node_pars and not const_j else None)
    if split is not None:
        try:
            consts2 = dict(consts)
            low = _Lowering(arg_kinds, consts2)
            tree = ast.parse(printed_source(model, split))
            results = [low.emit(e) for e in tree.body[0].body[-1].value.elts]
            lines = list(low.lines) + ["out[%d] = %s;" % (i, r) for i, r in enumerate(results)]
            st = dict(ops=low.n_ops, div=low.n_div, divc=low.n_divc)
            cost = lambda t: t["ops"] + 3 * t["divc"] + 20 * t["div"]  # noqa: E731
            if (cost(st) <= 0.67 * cost(stats["F"]) and not any("in.x" in ln for ln in lines)):
                consts.update(consts2)
                bodies["Fs"], stats["Fs"], out.f_split = lines, st, True
        except Exception:  # noqa: BLE001
            pass

    nvar = model._nvar
    kk = np.asarray(model._sparse_indices[0], dtype=int)
    col = kk // nvar
    out.nvar = nvar
    out.nhelp = len(model._help_funcs)
    out.nfield = len(fields)
    out.half_width = p
    out.nnz = len(kk)
    out.j_eq = (kk % nvar).tolist()
    out.j_var = (col % nvar).tolist()
    out.j_off = (col // nvar - p).tolist()
    out.consts = consts
    out.n_const = len(consts)
    out._const_code = {s: compile(s, "<uniform:%s>" % s, "eval") for s in consts}
    out.node_pars = node_pars
    out.uniform_pars = tuple(q for q in model._pars if q not in node_pars)
    out.uses_x = uses_x
    out.stats = stats
    out.jacobian_is_constant = (stats["J"]["ops"] == 0 and not node_pars and
                                all("in." not in ln for ln in bodies["J"]))
    out.f_is_linear = bool(out.jacobian_is_constant and not model._help_funcs and not uses_x and
                           _f_equals_j_times_u(model, fields, nvar, p, out.j_eq, out.j_var, out.j_off))
    if out.f_is_linear:
        out.f_split = False
    out.fields = fields
    out.header = _render_header(out, bodies)
    out.key = hashlib.sha1(out.header.encode()).hexdigest()[:16]
    return out


def _f_equals_j_times_u(model, fields, nvar, p, j_eq, j_var, j_off):
    """True when every equation is homogeneous linear in the stencil values, F_e = sum_k J_k u_k
    identically (checked symbolically): the solver kernels may then evaluate F as that sum -- a
    few fused multiply-adds with the uniform Jacobian constants instead of the reference's
    expanded expression (same value up to rounding; `model.F` keeps the reference's form)."""
    import sympy as sp
    F = model.F_array.tolist()
    J = model._J_sparse_array.tolist()
    syms = {}
    for e in F + J:
        for a in e.free_symbols:
            syms[a.name] = a
    try:
        for e in range(nvar):
            acc = F[e]
            for k in range(len(J)):
                if j_eq[k] != e:
                    continue
                o = j_off[k]
                name = fields[j_var[k]] if o == 0 else "%s_%s%d" % (fields[j_var[k]], "m" if o < 0 else "p", abs(o))
                if name not in syms:
                    return False
                acc = acc - J[k] * syms[name]
            if sp.expand(acc) != 0 and sp.simplify(acc) != 0:
                return False
    except Exception:  # noqa: BLE001  (anything unexpected: keep the general form)
        return False
    return True


def _split_linear_part(model, fields, ndep, p):
    """Every equation as a sum over its monomials in the stencil values, each with ONE collected
    coefficient (kept together for the printer, so that the lowering hoists it to the host):
    F_e = sum_m c_m * m(u).  The same function as the reference's expanded expression -- which
    divides every term separately -- with one multiplication per monomial.  Returns the list of
    expressions, or None."""
    import sympy as sp
    F = model.F_array.tolist()
    syms = {}
    for e in F:
        for a in e.free_symbols:
            syms[a.name] = a
    us = []
    for f in fields:
        for o in range(-p, p + 1):
            name = f if o == 0 else "%s_%s%d" % (f, "m" if o < 0 else "p", abs(o))
            if name in syms:
                us.append(syms[name])
    if "x" in syms:
        us.append(syms["x"])
    if not us:
        return None
    out = []
    try:
        for e in F:
            if sp.count_ops(e) > 400:
                return None
            groups = {}
            for term in sp.Add.make_args(sp.expand(e)):
                coeff, dep = term.as_independent(*us)
                groups[dep] = groups.get(dep, 0) + coeff
            parts = []
            for dep, coeff in groups.items():
                coeff = sp.together(coeff)
                if coeff == 0:
                    continue
                parts.append(dep if coeff == 1 else sp.UnevaluatedExpr(coeff) * dep)
            if not parts:
                return None
            out.append(sp.Add(*parts))
    except Exception:  # noqa: BLE001  (anything unexpected: keep the reference's form)
        return None
    return out


def _switch(name, values):
    cases = "".join(" case %d: return %d;" % (i, v) for i, v in enumerate(values))
    return ("TF_HD constexpr int %s(int k) { switch (k) {%s default: return 0; } }"
            % (name, cases))


def _render_header(L, bodies):
    h = []
    h.append("// generated by triflow_b200.codegen -- do not edit")
    h.append("#define TF_NVAR %d" % L.nvar)
    h.append("#define TF_NHELP %d" % L.nhelp)
    h.append("#define TF_NFIELD %d" % L.nfield)
    h.append("#define TF_P %d" % L.half_width)
    h.append("#define TF_WW %d" % (2 * L.half_width + 1))
    h.append("#define TF_NNZ %d" % L.nnz)
    h.append("#define TF_NCONST %d" % L.n_const)
    h.append("#define TF_NNODEPAR %d" % len(L.node_pars))
    h.append("#define TF_USES_X %d" % int(L.uses_x))
    h.append("#define TF_F_LINEAR %d" % int(getattr(L, "f_is_linear", False)))
    h.append("#define TF_F_SPLIT %d" % int(getattr(L, "f_split", False)))
    h.append('#include "tf_model_prelude.h"')
    h.append(_switch("tf_j_eq", L.j_eq))
    h.append(_switch("tf_j_var", L.j_var))
    h.append(_switch("tf_j_off", L.j_off))
    emit = [("F", "TF_NVAR"), ("J", "TF_NNZ")] + ([("Fs", "TF_NVAR")] if getattr(L, "f_split", False) else [])
    for which, n in emit:
        h.append("template <bool TF_FD> TF_HD TF_INLINE void tf_model_%s(const double* "
                 "TF_RESTRICT cst, const TfNodeIn& in, double (&out)[%s]) {" % (which, n))
        h.append("  (void)cst; (void)in;")
        h.extend("  " + ln for ln in bodies[which])
        h.append("}")
    return "\n".join(h) + "\n"
