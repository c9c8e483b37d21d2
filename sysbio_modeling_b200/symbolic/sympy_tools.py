"""Symbolic front-end: model file / RHS callable -> SymPy -> {Python RHS, CUDA device functions}.

Role on the hot path: the reference turns a ``#*!``-delimited model
file into SymPy expressions, derives the forward-sensitivity equations
``dS_ij/dt = df_i/dp_j + sum_k df_i/dy_k S_kj`` and writes Python source that numba compiles
(reference ``symbolic/sympy_tools.py:130-146, 162-216, 272-397``).  This module keeps that front
end (same section markers, same ``yout[N + i*P_free + j]`` layout for the Python emitter, which
feeds the CPU oracle) and adds the CUDA emitter the B200 kernels are built from.

The CUDA emitter does NOT print the expanded sensitivity equations.  It prints three pieces that
the integrator kernels combine per lane (see ``csrc/kernels/sbm_ode_kernels.cuh``):

* ``sbm_f``      - the RHS f(t, y, p);
* ``sbm_fJG``    - f plus the structural non-zeros of J = df/dy and G = df/dp_free, CSE'd together;
* ``sbm_col``    - one column of the augmented RHS: ``f`` for the state column, else
                   ``J*s + G[:, j]`` for sensitivity column j, with the lane-dependent pick of
                   ``G[:, j]`` done by integer selects (ALU pipe, not the FP64 pipe);
* ``sbm_jac_dense`` / ``sbm_dfdt`` - dense state Jacobian and time derivative for the Rosenbrock kernel.

Nothing here is copied from the reference; the model-file grammar and output layouts are kept so
that reference users can reuse their model files.
"""
from __future__ import annotations

import hashlib
import inspect
import io
import textwrap
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter
from sympy.printing.pycode import PythonCodePrinter

CATEGORIES = OrderedDict([
    ('Parameters', False), ('Variables', False), ('Conservation Laws', True),
    ('Rate Laws', True), ('Differential Equations', True), ('Imports', False),
])

T_SYM = sp.Symbol('t', real=True)


# ----------------------------------------------------------------------------------------------
# Parsing
# ----------------------------------------------------------------------------------------------
def _model_text(model):
    """Accept a path, an open file, a text blob or a python function (reference :285-297)."""
    if hasattr(model, 'py_func'):          # numba dispatcher
        model = model.py_func
    if inspect.isfunction(model):
        return inspect.getsource(model)
    if isinstance(model, str):
        if '#*!' in model:
            return model
        with open(model, 'r') as fh:
            return fh.read()
    if hasattr(model, 'read'):
        text = model.read()
        return text
    raise TypeError('model must be a path, file handle, source text or function')


def _chunk(text, category):
    start = text.find('#*! %s Start' % category)
    end = text.find('#*! %s End' % category)
    if start < 0 or end < 0:
        return []
    return text[start:end].split('\n')[1:]


def _clean(line):
    line = line.split('#')[0]
    return line.replace(' ', '').replace('\t', '').strip()


def parse_model_file(model):
    """Parse a ``#*!``-delimited model description into ordered dicts of SymPy objects.

    Same section names as the reference (``sympy_tools.py:300-311``).  Returns a dict with keys
    'Parameters', 'Variables' (name -> Symbol, in file order), 'Conservation Laws', 'Rate Laws',
    'Differential Equations' (name -> expression) and 'Imports' (raw lines).
    Unlike the reference, every declared name is bound to a plain Symbol before the right-hand
    sides are parsed, so species called ``S``, ``E``, ``N``... do not collide with SymPy globals.
    """
    text = _model_text(model)
    names = OrderedDict()
    parsed = {}
    for category in ('Parameters', 'Variables'):
        d = OrderedDict()
        for raw in _chunk(text, category):
            line = _clean(raw)
            if not line or '=' not in line:
                continue
            lhs = line[:line.find('=')]
            d[lhs] = sp.Symbol(lhs, real=True)
            names[lhs] = d[lhs]
        parsed[category] = d

    # left-hand sides of laws are names too (they get substituted away)
    law_lines = {}
    for category in ('Conservation Laws', 'Rate Laws', 'Differential Equations'):
        rows = []
        for raw in _chunk(text, category):
            line = _clean(raw)
            if not line or '=' not in line:
                continue
            lhs, rhs = line[:line.find('=')], line[line.find('=') + 1:]
            rows.append((lhs, rhs))
            if category != 'Differential Equations' and lhs not in names:
                names[lhs] = sp.Symbol(lhs, real=True)
        law_lines[category] = rows

    symtab = dict(names)
    symtab['t'] = T_SYM
    for category, rows in law_lines.items():
        d = OrderedDict()
        for lhs, rhs in rows:
            d[lhs] = sp.sympify(rhs, locals=symtab)
        parsed[category] = d
    parsed['Imports'] = [l for l in _chunk(text, 'Imports') if l.strip()]
    return parsed


# ----------------------------------------------------------------------------------------------
# Symbolic model container
# ----------------------------------------------------------------------------------------------
@dataclass
class SymbolicModel:
    """f(t, y, p) with its structural derivatives.

    ``J[(i, k)] = df_i/dy_k`` and ``G[(i, j)] = df_i/dp_free[j]`` hold non-zero entries only.
    ``free`` lists the non-fixed parameters in ``param_order`` order - this is the column order of
    the reference's sensitivity block (``sympy_tools.py:135-145``: variable-major, free-param-minor).
    """
    name: str
    variables: list
    var_syms: list
    param_order: list
    param_syms: list
    f: list
    fixed: list = field(default_factory=list)
    J: dict = field(default_factory=dict)
    G: dict = field(default_factory=dict)
    dfdt: list = field(default_factory=list)

    def __post_init__(self):
        self.fixed = [p for p in self.param_order if p in set(self.fixed)]
        self.free = [p for p in self.param_order if p not in set(self.fixed)]
        self.free_idx = [self.param_order.index(p) for p in self.free]
        self._derive()

    @property
    def n_vars(self):
        return len(self.variables)

    @property
    def n_params(self):
        return len(self.param_order)

    @property
    def n_free(self):
        return len(self.free)

    @property
    def autonomous(self):
        return all(d == 0 for d in self.dfdt)

    def _derive(self):
        self.J, self.G = {}, {}
        for i, fi in enumerate(self.f):
            for k, yk in enumerate(self.var_syms):
                d = sp.diff(fi, yk)
                if d != 0:
                    self.J[(i, k)] = d
            for j, m in enumerate(self.free_idx):
                d = sp.diff(fi, self.param_syms[m])
                if d != 0:
                    self.G[(i, j)] = d
        self.dfdt = [sp.diff(fi, T_SYM) for fi in self.f]

    def with_fixed(self, fixed_params):
        fixed_params = list(fixed_params or [])
        for p in fixed_params:
            if p not in self.param_order:
                raise KeyError('%s not in model parameters' % p)
        return SymbolicModel(self.name, self.variables, self.var_syms, self.param_order,
                             self.param_syms, self.f, fixed_params)

    def sens_symbols(self):
        """Symbols ``sens_<var>_<par>`` in the reference's augmented-state order."""
        return [[sp.Symbol('sens_%s_%s' % (v.lstrip('_'), p), real=True) for p in self.free]
                for v in self.variables]

    def sensitivity_equations(self):
        """Expanded dS_ij/dt, row-major (i, j) - reference ``_derive_sensitivity_equations``."""
        S = self.sens_symbols()
        out = []
        for i in range(self.n_vars):
            for j in range(self.n_free):
                e = self.G.get((i, j), sp.Integer(0))
                for k in range(self.n_vars):
                    if (i, k) in self.J:
                        e = e + self.J[(i, k)] * S[k][j]
                out.append(e)
        return out

    def fingerprint(self):
        h = hashlib.sha256()
        h.update(repr((self.variables, self.param_order, self.fixed,
                       [sp.srepr(e) for e in self.f])).encode())
        return h.hexdigest()[:16]


def _state_symbol_for(eq_name, variables):
    """Map a ``d_x`` left-hand side to the declared state it differentiates (``x`` or ``_x``)."""
    base = eq_name[2:] if eq_name.startswith('d_') else eq_name
    for cand in (base, '_' + base):
        if cand in variables:
            return cand
    raise KeyError('differential equation %s does not match any declared variable' % eq_name)


def symbolic_model_from_dict(model_dict, name='Model', fixed_params=None, simplify=False):
    """Expand rate and conservation laws into the ODEs (reference :351-353) and build the model."""
    params = model_dict['Parameters']
    variables = model_dict['Variables']
    rate_laws = model_dict.get('Rate Laws') or {}
    cons_laws = model_dict.get('Conservation Laws') or {}
    eqns = model_dict['Differential Equations']

    def expand(e):
        # laws may refer to each other: substitute to a fixed point
        for _ in range(8):
            new = e
            if rate_laws:
                new = new.subs({sp.Symbol(k, real=True): v for k, v in rate_laws.items()})
            if cons_laws:
                new = new.subs({sp.Symbol(k, real=True): v for k, v in cons_laws.items()})
            if new == e:
                break
            e = new
        return sp.simplify(e) if simplify else e

    by_var = {}
    for lhs, rhs in eqns.items():
        by_var[_state_symbol_for(lhs, variables)] = expand(rhs)
    missing = [v for v in variables if v not in by_var]
    if missing:
        raise KeyError('no differential equation for variables: %s' % ', '.join(missing))
    var_names = list(variables.keys())
    return SymbolicModel(name, var_names, [variables[v] for v in var_names],
                         list(params.keys()), list(params.values()),
                         [by_var[v] for v in var_names], list(fixed_params or []))


def symbolic_model_from_file(model, name='Model', fixed_params=None, simplify=False):
    return symbolic_model_from_dict(parse_model_file(model), name, fixed_params, simplify)


def trace_rhs_callable(model, n_vars, param_order, name='Model', fixed_params=None):
    """Recover f symbolically from a reference-ABI callable ``model(y, t, yout, p)``.

    The reference's ``OdeModel`` is handed plain Python functions (``ode_model.py:27-29``; RHS ABI
    ``void f(y, t, yout, p)``, :50-51).  Calling such a function on SymPy symbols executes its
    arithmetic symbolically, which is how ``OdeModel(backend='cuda')`` stays a drop-in for models
    that exist only as generated Python (e.g. reference ``tests/test_utils/jittable_model.py``).
    """
    if hasattr(model, 'py_func'):
        model = model.py_func
    ys = [sp.Symbol('y%d' % i, real=True) for i in range(n_vars)]
    ps = [sp.Symbol(str(p), real=True) for p in param_order]
    yout = [sp.Integer(0)] * n_vars
    model(ys, T_SYM, yout, ps)
    return SymbolicModel(name, ['y%d' % i for i in range(n_vars)], ys, [str(p) for p in param_order],
                         ps, [sp.sympify(e) for e in yout], list(fixed_params or []))


# ----------------------------------------------------------------------------------------------
# Reference-shaped API (dict in / dict out)
# ----------------------------------------------------------------------------------------------
def process_model_dict(model_dict, fixed_params=None, calculate_model_sensitivities=True,
                       simplify_subexpressions=False, calculate_model_jacobian=False):
    """Reference-shaped driver (``sympy_tools.py:325-397``): fills 'Sensitivity Equations',
    'Model Jacobian Equations' and 'Subexpressions' and stores the SymbolicModel under 'Symbolic'."""
    sm = symbolic_model_from_dict(model_dict, fixed_params=fixed_params)
    model_dict['Symbolic'] = sm
    if fixed_params:
        for p in fixed_params:
            model_dict['Parameters'][p] = 'fixed'
    sens = None
    if calculate_model_sensitivities:
        sens = OrderedDict()
        eqs = sm.sensitivity_equations()
        q = 0
        for v in sm.variables:
            for p in sm.free:
                sens['d_sens_%s_%s' % (v.lstrip('_'), p)] = eqs[q]
                q += 1
    model_dict['Sensitivity Equations'] = sens
    jac = None
    if calculate_model_jacobian:
        jac = OrderedDict()
        for i, vi in enumerate(sm.variables):
            for k, vk in enumerate(sm.variables):
                jac['d_%s_d_%s' % (vi, vk)] = sm.J.get((i, k), sp.Integer(0))
    model_dict['Model Jacobian Equations'] = jac
    sub = None
    if simplify_subexpressions:
        exprs = list(sm.f) + (list(sens.values()) if sens else [])
        rep, red = sp.cse(exprs, optimizations='basic')
        sub = OrderedDict(rep)
        model_dict['Reduced Equations'] = red
    model_dict['Subexpressions'] = sub
    return model_dict


# ----------------------------------------------------------------------------------------------
# Operation counting (the ALGORITHMIC flop model of DESIGN.md)
# ----------------------------------------------------------------------------------------------
def count_flops(exprs):
    """1 per add/sub/mul/div; integer power k costs k-1 multiplies (+1 divide if negative);
    real pow / exp / log / other functions cost 1 each.  An FMA therefore counts 2."""
    total = 0
    for e in exprs:
        for node in sp.preorder_traversal(e):
            if isinstance(node, sp.Add) or isinstance(node, sp.Mul):
                n = len(node.args)
                # a leading -1 coefficient is a sign flip folded into the add/sub, not a multiply
                if isinstance(node, sp.Mul) and node.args[0] == -1:
                    n -= 1
                total += max(n - 1, 0)
            elif isinstance(node, sp.Pow):
                ex = node.exp
                if ex.is_Integer:
                    k = abs(int(ex))
                    total += max(k - 1, 0) + (1 if ex < 0 else 0)
                else:
                    total += 1
            elif isinstance(node, sp.Function):
                total += 1
    return total


# ----------------------------------------------------------------------------------------------
# Printers
# ----------------------------------------------------------------------------------------------
class sbm_rcp(sp.Function):
    """Explicit reciprocal node: lets CSE share 1/x between x**-1, x**-2, ... (one divide each)."""
    nargs = 1

    def fdiff(self, argindex=1):
        return -sbm_rcp(self.args[0]) ** 2


class sbm_exp(sp.Function):
    """Opaque exp/log pair used to print x**a (a symbolic) as exp(a*log(x)): SymPy would otherwise fold
    it back into a Pow.  Lets CSE share log(x) between x**a and d(x**a)/da = x**a*log(x)."""
    nargs = 1


class sbm_log(sp.Function):
    nargs = 1


def _share_reciprocals(e, pow_via_explog=True):
    """Rewrite x**-k as sbm_rcp(x)**k so that the emitted code divides once per denominator, and
    (optionally) x**a with a non-numeric exponent as exp(a*log(x)) - valid for positive bases, which is
    what real powers of concentrations require anyway (a negative base gives NaN either way)."""
    e = e.replace(lambda n: n.is_Pow and n.exp.is_Integer and n.exp < 0,
                  lambda n: sbm_rcp(n.base) ** (-n.exp))
    if pow_via_explog:
        e = e.replace(lambda n: n.is_Pow and not n.exp.is_number,
                      lambda n: sbm_exp(n.exp * sbm_log(n.base)))
        e = e.replace(lambda n: isinstance(n, sp.log), lambda n: sbm_log(n.args[0]))
    return e


class _CudaPrinter(C99CodePrinter):
    """C printer that expands small integer powers into multiplies and reciprocals."""

    def _print_sbm_rcp(self, expr):
        # branch-free reciprocal defined in csrc/sbm_device_math.h (MUFU seed + Newton steps)
        return 'sbm_rcp(%s)' % self._print(expr.args[0])

    def _print_sbm_exp(self, expr):
        return 'sbm_exp(%s)' % self._print(expr.args[0])      # csrc/sbm_device_math.h

    def _print_sbm_log(self, expr):
        return 'sbm_log(%s)' % self._print(expr.args[0])

    def _print_Pow(self, expr):
        base, ex = expr.base, expr.exp
        if ex.is_Integer and 1 <= abs(int(ex)) <= 8:
            b = self.parenthesize(base, 1000)
            prod = '*'.join([b] * abs(int(ex)))
            if ex > 0:
                return '(%s)' % prod
            return '(1.0/(%s))' % prod
        if ex == sp.Rational(1, 2):
            return 'sqrt(%s)' % self._print(base)
        if ex == -sp.Rational(1, 2):
            return 'rsqrt(%s)' % self._print(base)
        return 'pow(%s, %s)' % (self._print(base), self._print(ex))

    def _print_Integer(self, expr):
        return '%d.0' % int(expr)

    def _print_Rational(self, expr):
        return '(%d.0/%d.0)' % (expr.p, expr.q)


class _PyPrinter(PythonCodePrinter):
    def _print_Pow(self, expr, rational=False):
        base, ex = expr.base, expr.exp
        if ex.is_Integer and 1 <= abs(int(ex)) <= 4:
            b = self.parenthesize(base, 1000)
            prod = '*'.join([b] * abs(int(ex)))
            return '(%s)' % prod if ex > 0 else '(1.0/(%s))' % prod
        return '(%s)**(%s)' % (self.parenthesize(base, 1000), self._print(ex))


_PY_FUNCS = {'math.exp': 'exp', 'math.log': 'log', 'math.sqrt': 'sqrt', 'math.sin': 'sin',
             'math.cos': 'cos', 'math.tanh': 'tanh'}


def _py(expr):
    s = _PyPrinter({'fully_qualified_modules': True}).doprint(expr)
    for k, v in _PY_FUNCS.items():
        s = s.replace(k, v)
    return s


# ----------------------------------------------------------------------------------------------
# Python emitter (reference ABI; feeds the CPU oracle and the reference arm of bench.py)
# ----------------------------------------------------------------------------------------------
def emit_python(sm, sensitivities=True, cse=True, fcn_name=None):
    """Source of ``model(y, t, yout, p)`` or ``sens_model(y, t, yout, p)`` in the reference's
    generated-file format (``sympy_tools.py:38-127,185-195``): ``yout[0:N] = f`` and
    ``yout[N + i*P_free + j] = dS_ij/dt``.  Pure arithmetic, so ``numba.jit`` accepts it
    (``ode_model.py:50-51``)."""
    fcn_name = fcn_name or ('sens_model' if sensitivities else 'model')
    exprs = list(sm.f)
    S = sm.sens_symbols()
    if sensitivities:
        exprs += sm.sensitivity_equations()
    if cse:
        rep, red = sp.cse(exprs, symbols=sp.numbered_symbols('x_'), optimizations='basic')
    else:
        rep, red = [], exprs
    pad = '    '
    L = ['from math import exp, log, sqrt, sin, cos, tanh', '',
         'ordered_params = [%s]' % ', '.join("'%s'" % p for p in sm.param_order),
         'n_vars = %d' % sm.n_vars,
         'free_params = [%s]' % ', '.join("'%s'" % p for p in sm.free), '', '',
         'def %s(y, t, yout, p):' % fcn_name]
    for m, ps in enumerate(sm.param_syms):
        L.append(pad + '%s = p[%d]' % (ps, m))
    for i, ys in enumerate(sm.var_syms):
        L.append(pad + '%s = y[%d]' % (ys, i))
    if sensitivities:
        for i in range(sm.n_vars):
            for j in range(sm.n_free):
                L.append(pad + '%s = y[%d]' % (S[i][j], sm.n_vars + i * sm.n_free + j))
    for lhs, rhs in rep:
        L.append(pad + '%s = %s' % (lhs, _py(rhs)))
    for q, e in enumerate(red):
        L.append(pad + 'yout[%d] = %s' % (q, _py(e)))
    return '\n'.join(L) + '\n'


def make_ode_model(sm, sensitivities=False, cse=True, output_fh=None):
    """Exec the emitted Python and return the callable (reference ``make_ode_model`` :162-216)."""
    src = emit_python(sm, sensitivities=sensitivities, cse=cse)
    if output_fh is not None:
        output_fh.write(src)
    ns = {}
    exec(compile(src, '<sbm generated %s>' % sm.name, 'exec'), ns)
    fn = ns['sens_model' if sensitivities else 'model']
    fn.n_vars = sm.n_vars
    fn.ordered_params = list(sm.param_order)
    fn.state_variables = list(sm.variables)
    fn.source = src
    return fn


# ----------------------------------------------------------------------------------------------
# CUDA emitter
# ----------------------------------------------------------------------------------------------
@dataclass
class CudaModelSource:
    source: str
    n_vars: int
    n_params: int
    n_free: int
    free_idx: list
    nnz_J: int
    nnz_G: int
    flops_f: int          # one evaluation of f
    flops_fJG: int        # f + J + G non-zeros, CSE'd, evaluated once
    flops_col: int        # one sensitivity column: J*s + G[:, j]
    flops_sens_rhs: int   # ALGORITHMIC cost of one augmented RHS evaluation (non-redundant)
    flops_jac_dense: int
    autonomous: bool
    fingerprint: str


def _emit_block(lines, rep, outs, printer, pad='    '):
    for lhs, rhs in rep:
        lines.append(pad + 'const double %s = %s;' % (lhs, printer.doprint(rhs)))
    for target, e in outs:
        lines.append(pad + '%s = %s;' % (target, printer.doprint(e)))



def ct_layout(n_vars, n_free, cols=0):
    """Shape of the CTA-transposed sensitivity kernel (csrc/kernels/sbm_dp5_ct.cuh): W warps per CTA, each
    carrying C sensitivity columns of N states (three register-resident vectors per column), W*C >= n_free.
    Small systems (N * NF <= 24) run one warp per CTA = one thread per trajectory.  Larger ones take the
    smallest C that needs at most 11 warps (168 registers per thread; measured on the MAPK model: C = 2 x 11
    warps beats C = 3 x 8 warps, because the unrolled stage code is smaller and more warps hide latency).
    Returns (W, C), or (0, 0) when the model does not fit (the caller then uses the warp-group kernel)."""
    if n_free < 1 or n_vars > 24:
        return 0, 0
    cmax = min(max(24 // n_vars, 1), n_free)
    if cols and cols * n_vars <= 32:
        C = min(cols, n_free)
    elif n_vars * n_free <= 24:
        C = n_free
    else:
        C = cmax
        for c in range(1, cmax + 1):
            if -(-n_free // c) <= 11:
                C = c
                break
    W = -(-n_free // C)
    if W > (32 if cols else 11):
        return 0, 0
    return W, C


def ct_share_warps(W, n_shares):
    """Warps that evaluate the model's primal shares: spread over the schedulers (warp w runs on scheduler w % 4)
    that host the fewest warps, highest warp index first."""
    count = [len(range(q, W, 4)) for q in range(4)]
    free = [sorted(range(q, W, 4)) for q in range(4)]
    out = []
    while len(out) < min(n_shares, W):
        q = min((q for q in range(4) if free[q]), key=lambda q: (count[q], -q))
        out.append(free[q].pop())
        count[q] += 1
    return out


def _instr_cost(e):
    """Rough instruction count of an expression on the FP64 pipe (balancing only)."""
    c = count_flops([e])
    for node in sp.preorder_traversal(e):
        if isinstance(node, sbm_rcp):
            c += 6
        elif isinstance(node, (sbm_log, sp.log)):
            c += 34
        elif isinstance(node, (sbm_exp, sp.exp)):
            c += 29
        elif isinstance(node, sp.Pow) and not node.exp.is_Integer:
            c += 60
    return c


def _emit_primal_shares(L, pr, N, f_x, J_x, G_x, gkeys, W, C, NF, ysyms, XW=0, fname='sbm_primal_share',
                        n_shares=None, head=False, jpos=None, jslot=None, jrep=None):
    """Split the model's primal part - f, the non-zeros of J = df/dy and of G = df/dp - into W shares, one per
    warp of the CTA-transposed kernel.  Each share is straight-line code that recomputes the sub-expressions it
    needs (a share that owns several outputs of one rate law computes its reciprocal once), reads only the stage
    values of the states it depends on, and stores its outputs straight into shared memory ([entry][lane]).
    Greedy longest-processing-time assignment on an instruction-count estimate; the starting load of a warp
    is the work of its real sensitivity columns, so warps with empty column slots take more of the model."""
    outs = [('f', i, e) for i, e in enumerate(f_x)] + [('J', q, e) for q, e in enumerate(J_x)] + \
           [('G', q, e) for q, e in enumerate(G_x)]
    rep, red = sp.cse([e for _, _, e in outs], symbols=sp.numbered_symbols('s_'), optimizations='basic')
    rsym = [r[0] for r in rep]
    rexpr = dict(rep)
    rset = set(rsym)
    rcost = {k: _instr_cost(v) for k, v in rep}
    deps = {k: [x for x in v.free_symbols if x in rset] for k, v in rep}

    def closure(expr):
        todo = [x for x in expr.free_symbols if x in rset]
        seen = set()
        while todo:
            x = todo.pop()
            if x in seen:
                continue
            seen.add(x)
            todo.extend(deps[x])
        return seen
    items = []
    for (kind, idx, _), e in zip(outs, red):
        cl = closure(e)
        items.append((kind, idx, e, cl, _instr_cost(e) + 2 + sum(rcost[x] for x in cl)))
    # starting load of a warp = estimated instructions of its sensitivity columns per stage: the J entries are
    # loaded once per warp, each column costs nnz(J) FMAs plus ~12 instructions per state for the tableau
    # updates and their loads/stores; warp 0 also runs the step controller
    nnzJ = len(J_x)
    import os as _os
    if n_shares is None:
        n_shares = int(_os.environ.get('SBM_CT_SHARES', '0')) or (1 if XW else min(W, 3))
    WT = W + XW                                 # XW warps carry no columns: they only evaluate the model
    share_warps = [int(x) for x in _os.environ.get('SBM_CT_SHARE_WARPS', '').split(',') if x.strip()] \
        if not head else []
    if not share_warps:
        share_warps = list(range(W, WT)) + [w for w in ct_share_warps(WT, n_shares + XW) if w < W][:max(n_shares - XW, 0)]
    colwork = []
    for w in range(WT):
        real = max(0, min(C, NF - w * C)) if w < W else 0
        colwork.append((float(nnzJ + real * (nnzJ + 12 * N)) if real else 0.0) + (60.0 if (w == 0 and not XW) else 0.0))
    loads = []
    for w in range(WT):
        if head:
            # the model at the input of stage 2, evaluated while the column warps have nothing else to do - except the
            # start of their step (accept copy, Jacobian rows of an output time: ~1,000 clocks measured, about a
            # third of the whole model), so the model warp takes the largest cluster (the longest dependent chain)
            base = float(_os.environ.get('SBM_CT_HEAD_COLBASE', '200')) if w < W else 0.0
        elif XW:
            # several shares with a model warp: balance the load of the schedulers (warp w runs on scheduler w % 4), not
            # of the warps - the model warp's scheduler hosts one column warp fewer.  Measured on C4: 1 / 4 / 7 / 10
            # shares = 141.8 / 124.5 / 118.9 / 123.1 k solves/s, i.e. ANY slice of the model on a column warp costs
            # more than the balance gains (the column warps' own latency is the critical path): default 1 share.
            base = sum(colwork[v] for v in range(w % 4, WT, 4)) + (80.0 if w >= W else 0.0)
        else:
            base = colwork[w]
        loads.append(base + (0.0 if w in share_warps else 1e9))
    have = [set() for _ in range(WT)]
    mine = [[] for _ in range(WT)]
    ycomp = [set() for _ in range(WT)]
    yset = set(ysyms)
    Y_COST = 3                                  # 2 LDS + 1 FMA to form the stage value of one state

    def ys_of(e, cl):
        out = set(x for x in e.free_symbols if x in yset)
        for x in cl:
            out |= set(z for z in rexpr[x].free_symbols if z in yset)
        return out

    def assign(w, group):
        for kind, idx, e, cl, _ in group:
            loads[w] += _instr_cost(e) + 2 + sum(rcost[x] for x in cl if x not in have[w]) + \
                Y_COST * len(ys_of(e, cl) - ycomp[w])
            have[w] |= cl
            ycomp[w] |= ys_of(e, cl)
            mine[w].append((kind, idx, e))

    def incr(w, group):
        h, yc, v = set(have[w]), set(ycomp[w]), 0.0
        for kind, idx, e, cl, _ in group:
            v += _instr_cost(e) + 2 + sum(rcost[x] for x in cl if x not in h) + Y_COST * len(ys_of(e, cl) - yc)
            h |= cl
            yc |= ys_of(e, cl)
        return v
    # clusters: J and G outputs that share an expensive sub-expression (a reciprocal, exp, log ...) stay in
    # one share - in a reaction network these are the outputs of one rate law
    jg = [it for it in items if it[0] != 'f']
    parent = list(range(len(jg)))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    owner = {}
    for n, it in enumerate(jg):
        for x in it[3]:
            if rcost[x] >= 4:
                if x in owner:
                    parent[find(n)] = find(owner[x])
                else:
                    owner[x] = n
    clusters = {}
    for n, it in enumerate(jg):
        clusters.setdefault(find(n), []).append(it)
    groups = sorted(clusters.values(), key=lambda g: -sum(it[4] for it in g))
    for g in groups:
        w = min(share_warps, key=lambda w: loads[w] + incr(w, g))
        assign(w, g)
    for it in sorted([it for it in items if it[0] == 'f'], key=lambda it: -it[4]):
        w = min(share_warps, key=lambda w: loads[w] + incr(w, [it]))
        assign(w, [it])
    L.append('// primal shares of the CTA-transposed kernel: estimated instructions per warp = %s'
             % ', '.join('%d' % int(v) for v in loads))
    L.append('template <int PS, int YS = 32>   // YS: stride of Pp / fl (32: [state][lane] in shared memory, 1: a register array)')
    L.append('// Pp: partial stage value y + h sum_{j<s} a_{s+1,j} f_j (prepared one iteration earlier), fl: f_s, cl = h a_{s+1,s}')
    L.append('__device__ __forceinline__ void %s(const int share, const double t,' % fname)
    L.append('        const double* __restrict__ Pp, const double* __restrict__ fl, const double cl,')
    L.append('        const double* __restrict__ p, double* __restrict__ Jo, double* __restrict__ Go,')
    L.append('        double* __restrict__ fo)')
    L.append('{')
    L.append('    switch (share) {')
    ypat = {str(sy): i for i, sy in enumerate(ysyms)}
    for w in range(WT):
        L.append('    case %d: {' % w)
        need_syms = set()
        for x in have[w]:
            need_syms |= rexpr[x].free_symbols
        for _, _, e in mine[w]:
            need_syms |= e.free_symbols
        need = sorted(ypat[str(x)] for x in need_syms if str(x) in ypat)
        for i in need:
            # stage value of state i = prepared partial sum + the newest stage derivative
            L.append('        const double Y%d = fma(cl, fl[%d * YS], Pp[%d * YS]);' % (i, i, i))
        body = []
        for k in rsym:
            if k in have[w]:
                body.append('        const double %s = %s;' % (k, pr.doprint(rexpr[k])))
        for kind, idx, e in mine[w]:
            if kind == 'f':
                body.append('        fo[%d * 32] = %s;' % (idx, pr.doprint(e)))
            elif kind == 'J':
                if jrep is not None and jrep.get(idx) is not None:
                    body.append('#if SBM_CT_JTM')
                    body.append('        Jo[SBM_JPOS(%d, %d)] = %s;' % (idx, jpos[idx], pr.doprint(e)))
                    body.append('#endif')
                else:
                    body.append('        Jo[SBM_JPOS(%d, %d)] = %s;' % (jslot[idx] if jslot is not None else idx, jpos[idx], pr.doprint(e)))
            else:
                i, j = gkeys[idx]
                body.append('        Go[SBM_GIDX(%d)] = %s;' % (j * (N + (N & 1)) + i, pr.doprint(e)))
        import re as _re
        for ln in body:
            ln = _re.sub(r'p\[(\d+)\]', r'p[\1 * PS]', ln)
            ln = _re.sub(r'y\[(\d+)\]', r'Y\1', ln)
            L.append(ln)
        L.append('    } break;')
    L.append('    default: break;')
    L.append('    }')
    L.append('}')
    L.append('')
    return [int(v) for v in loads]


def emit_cuda(sm, pow_via_explog=True, ct_cols=0):
    """Emit the device functions the integrator kernels are compiled against (NVRTC, sm_100a)."""
    pr = _CudaPrinter()
    N, NP, NF = sm.n_vars, sm.n_params, sm.n_free
    jkeys = sorted(sm.J.keys())
    gkeys = sorted(sm.G.keys())

    # substitute array accesses for the named symbols
    ysub = {s: sp.Symbol('y[%d]' % i, real=True) for i, s in enumerate(sm.var_syms)}
    psub = {s: sp.Symbol('p[%d]' % m, real=True) for m, s in enumerate(sm.param_syms)}
    sub = {}
    sub.update(ysub)
    sub.update(psub)

    def X(e):
        return _share_reciprocals(sp.sympify(e).xreplace(sub), pow_via_explog)

    f_x = [X(e) for e in sm.f]
    J_x = [X(sm.J[k]) for k in jkeys]
    G_x = [X(sm.G[k]) for k in gkeys]

    L = []
    L.append('// ---- generated by sysbio_modeling_b200.symbolic.emit_cuda ----')
    L.append('// states : %s' % ', '.join(sm.variables))
    L.append('// params : %s' % ', '.join(sm.param_order))
    L.append('// free   : %s' % ', '.join(sm.free))
    L.append('#define SBM_N %d' % N)
    L.append('#define SBM_NP %d' % NP)
    L.append('#define SBM_NF %d' % NF)
    L.append('#define SBM_NJ %d' % max(len(jkeys), 1))
    L.append('#define SBM_NG %d' % max(len(gkeys), 1))
    L.append('#define SBM_AUTONOMOUS %d' % (1 if sm.autonomous else 0))
    L.append('#define SBM_HAS_FREE_IDX 1')
    L.append('// model-parameter index of each free (sensitivity) parameter')
    L.append('__device__ __constant__ int sbm_free_idx[%d] = {%s};'
             % (max(NF, 1), ', '.join(str(m) for m in (sm.free_idx or [0]))))
    L.append('')

    # --- f only
    rep, red = sp.cse(f_x, symbols=sp.numbered_symbols('a_'), optimizations='basic')
    flops_f = count_flops([r for _, r in rep] + list(red))
    L.append('__device__ __forceinline__ void sbm_f(const double t, const double* __restrict__ y,')
    L.append('        const double* __restrict__ p, double* __restrict__ f)')
    L.append('{')
    _emit_block(L, rep, [('f[%d]' % i, e) for i, e in enumerate(red)], pr)
    L.append('}')
    L.append('')

    # --- f + J + G non-zeros
    allx = f_x + J_x + G_x
    rep, red = sp.cse(allx, symbols=sp.numbered_symbols('b_'), optimizations='basic')
    flops_fJG = count_flops([r for _, r in rep] + list(red))
    outs = [('f[%d]' % i, red[i]) for i in range(N)]
    outs += [('Jv[%d]' % q, red[N + q]) for q in range(len(jkeys))]
    outs += [('Gv[%d]' % q, red[N + len(jkeys) + q]) for q in range(len(gkeys))]
    L.append('// Jv[q] = df_i/dy_k for (i,k) in: %s' % ' '.join('%d:(%d,%d)' % (q, i, k)
                                                             for q, (i, k) in enumerate(jkeys)))
    L.append('// Gv[q] = df_i/dp_free[j] for (i,j) in: %s' % ' '.join('%d:(%d,%d)' % (q, i, j)
                                                                   for q, (i, j) in enumerate(gkeys)))
    L.append('// PS = stride of the parameter vector (32 when it lives in shared memory as [param][lane])')
    L.append('template <int PS = 1>')
    L.append('__device__ __forceinline__ void sbm_fJG(const double t, const double* __restrict__ y,')
    L.append('        const double* __restrict__ p, double* __restrict__ f,')
    L.append('        double* __restrict__ Jv, double* __restrict__ Gv)')
    L.append('{')
    blk = []
    _emit_block(blk, rep, outs, pr)
    import re as _re
    L.extend(_re.sub(r'p\[(\d+)\]', r'p[\1 * PS]', ln) for ln in blk)
    L.append('}')
    L.append('')

    # --- one column of the augmented RHS
    # cid == 0 : the state column (out = f);  cid == j+1 : sensitivity w.r.t. free parameter j
    L.append('__device__ __forceinline__ void sbm_col(const double* __restrict__ f,')
    L.append('        const double* __restrict__ Jv, const double* __restrict__ Gv,')
    L.append('        const double* __restrict__ s, const int cid, double* __restrict__ out)')
    L.append('{')
    flops_col = 0
    for i in range(N):
        terms = ['Jv[%d]*s[%d]' % (q, k) for q, (ii, k) in enumerate(jkeys) if ii == i]
        gsel = [(q, j) for q, (ii, j) in enumerate(gkeys) if ii == i]
        L.append('    {')
        L.append('        double g = 0.0;')
        for q, j in gsel:
            L.append('        g = (cid == %d) ? Gv[%d] : g;' % (j + 1, q))
        acc = 'g'
        for tm in terms:
            acc = 'fma(%s, %s, %s)' % (tm.split('*')[0], tm.split('*')[1], acc)
        flops_col += 2 * len(terms)
        L.append('        const double d = %s;' % acc)
        L.append('        out[%d] = (cid == 0) ? f[%d] : d;' % (i, i))
        L.append('    }')
    L.append('}')
    L.append('')

    # --- entries of J that are equal up to sign (conservation: a reaction leaves one pool and enters another, so the rows
    #     of a conserved pair differ only in sign) share ONE slot of the shared-memory J buffer of the CTA-transposed kernel:
    #     fewer stores for the model warp, fewer loads for every column warp (the DFMA takes the sign for free)
    jslot, jsign, jrep, n_ju = {}, {}, {}, 0
    # candidates by numerical evaluation at two random points (cheap), confirmed symbolically (exact)
    syms = sorted(set().union(*[e.free_symbols for e in J_x]) if J_x else [], key=str)
    vals = []
    if syms:
        rng_ = np.random.default_rng(12345)
        fn_ = sp.lambdify(syms, list(J_x), modules=[{'sbm_rcp': lambda x: 1.0 / x, 'sbm_log': np.log,
                                                                                         'sbm_exp': np.exp}, 'numpy'])
        for _ in range(2):
            vals.append(np.array(fn_(*list(0.5 + rng_.random(len(syms)))), dtype=np.float64))
    for q, e in enumerate(J_x):
        hit = None
        for q2 in range(q):
            if jrep[q2] is not None:
                continue
            sgn = 0
            if vals:
                a0, b0, a1, b1 = vals[0][q], vals[0][q2], vals[1][q], vals[1][q2]
                tol_ = 1e-12 * max(abs(a0), abs(a1), 1e-300)
                if abs(a0 - b0) <= tol_ and abs(a1 - b1) <= tol_:
                    sgn = 1
                elif abs(a0 + b0) <= tol_ and abs(a1 + b1) <= tol_:
                    sgn = -1
            else:
                sgn = 1 if e == J_x[q2] else (-1 if e == -J_x[q2] else 0)
            if sgn and sp.simplify(e - sgn * J_x[q2]) == 0:
                hit = (q2, sgn)
                break
        if hit is None:
            jrep[q] = None
            jslot[q], jsign[q] = n_ju, 1
            n_ju += 1
        else:
            jrep[q] = hit[0]
            jslot[q], jsign[q] = jslot[hit[0]], hit[1]
    L.append('#define SBM_NJU %d' % max(n_ju, 1))
    # --- row i of J*w for C columns at once, J read from shared memory ([entry][lane]); used by the
    #     CTA-transposed kernel (kernels/sbm_dp5_ct.cuh), where `i` is a constant after unrolling
    L.append('#ifndef SBM_JLOAD')
    L.append('#define SBM_JLOAD(q) Jp[(q) * 32]')
    L.append('#endif')
    L.append('template <int C>')
    L.append('__device__ __forceinline__ void sbm_jrow(const int i, const double* __restrict__ Jp,')
    L.append('        const double (&w)[C][SBM_N], double (&acc)[C])')
    L.append('{')
    for i in range(N):
        ent = [(q, k) for q, (ii, k) in enumerate(jkeys) if ii == i]
        L.append('    %sif (i == %d) {' % ('else ' if i else '', i))
        for sl in sorted(set(jslot[q] for q, k in ent)):
            L.append('        const double j%d = SBM_JLOAD(%d);' % (sl, sl))
        if ent:
            expr = 'acc[c]'
            for q, k in ent:
                expr = 'fma(%sj%d, w[c][%d], %s)' % ('-' if jsign[q] < 0 else '', jslot[q], k, expr)
            L.append('#pragma unroll')
            L.append('        for (int c = 0; c < C; ++c) acc[c] = %s;' % expr)
        L.append('    }')
    L.append('}')
    # --- the same for a row PAIR with the entries of J in registers (loaded from tensor memory, where the model warp's
    #     J arrives through tcgen05.cp): the entries of rows 2 rp and 2 rp + 1 are adjacent, each row pair padded to
    #     SBM_JRP doubles (a power of two: one tcgen05.ld per row pair)
    NP2 = (N + 1) // 2
    rp_ent = [[(q, ii, k) for q, (ii, k) in enumerate(jkeys) if ii // 2 == rp] for rp in range(NP2)]
    JRP = 2
    while JRP < max([len(e) for e in rp_ent] + [1]):
        JRP *= 2
    jpos = {}
    for rp in range(NP2):
        for n_, (q, ii, k) in enumerate(rp_ent[rp]):
            jpos[q] = rp * JRP + n_
    L.append('#define SBM_JRP %d' % JRP)
    L.append('#define SBM_NJP %d' % (NP2 * JRP))
    L.append('template <int C>')
    L.append('__device__ __forceinline__ void sbm_jrow_tm(const int rp, const double (&jr)[SBM_JRP],')
    L.append('        const double (&w)[C][SBM_N], double (&acc0)[C], double (&acc1)[C])')
    L.append('{')
    for rp in range(NP2):
        L.append('    %sif (rp == %d) {' % ('else ' if rp else '', rp))
        for r_, accn in ((2 * rp, 'acc0'), (2 * rp + 1, 'acc1')):
            ent = [(n_, k) for n_, (q, ii, k) in enumerate(rp_ent[rp]) if ii == r_]
            if ent:
                expr = '%s[c]' % accn
                for n_, k in ent:
                    expr = 'fma(jr[%d], w[c][%d], %s)' % (n_, k, expr)
                L.append('#pragma unroll')
                L.append('        for (int c = 0; c < C; ++c) %s[c] = %s;' % (accn, expr))
        L.append('    }')
    L.append('}')
    W_ct, C_ct = ct_layout(N, NF, ct_cols)
    import os as _os
    # one extra warp without columns: it evaluates the whole model and runs the step controller, so that the column
    # warps (the critical path) carry no share of the model.  At most 12 warps = 168 registers per thread.
    XW_ct = 1 if 2 <= W_ct <= 11 else 0
    if 'SBM_CT_XW' in _os.environ:
        XW_ct = int(_os.environ['SBM_CT_XW']) if W_ct else 0
    L.append('// dense G of the CTA-transposed kernel: element e = parameter * NE + state (NE = N rounded up to even) lives')
    L.append('// at double index (e/2)*64 + 2*lane + (e&1): the two states of a row pair sit side by side (128-bit loads)')
    L.append('#define SBM_GIDX(e) ((((e) >> 1) * 64) + ((e) & 1))')
    L.append('#define SBM_CT_W %d' % W_ct)
    L.append('#define SBM_CT_C %d' % C_ct)
    L.append('#define SBM_CT_XW %d' % XW_ct)
    # S1 / S4 of the sensitivity columns live in tensor memory: ceil(W/4) warps per lane quadrant, 4*C*NE columns each
    NE_ct = N + (N & 1)
    # ... except for small systems run by one thread per trajectory, whose five vectors fit the register file
    regv = 1 if (W_ct == 1 and C_ct * NE_ct <= 16) else 0
    if 'SBM_CT_REGV' in _os.environ:
        regv = int(_os.environ['SBM_CT_REGV']) if W_ct else 0
    L.append('#define SBM_CT_REGV %d' % regv)
    # producer / consumer hand-over of J and G (mbarriers, 3 buffers) when a dedicated model warp evaluates the whole model
    n_sh = int(_os.environ.get('SBM_CT_SHARES', '0')) or (1 if XW_ct else min(W_ct, 3))
    # (measured on C4: 129 k solves/s against 143.6 k with one CTA barrier per stage - the single model warp then also
    #  prepares the stage values alone and becomes the critical path; kept as an option, off by default)
    asyn = 0
    if 'SBM_CT_ASYNC' in _os.environ:
        asyn = int(_os.environ['SBM_CT_ASYNC']) if (XW_ct == 1 and n_sh == 1) else 0
    L.append('#define SBM_CT_ASYNC %d' % asyn)
    tm_cols = ((W_ct + 3) // 4) * 4 * C_ct * NE_ct
    L.append('#define SBM_CT_TMEM %d' % (1 if (W_ct and not regv and tm_cols <= 512) else 0))
    # J of a stage handed to the column warps through tensor memory (tcgen05.cp broadcasts the model warp's [lane][16 B]
    # blocks to the four lane quadrants) instead of W x NJ shared-memory loads per stage: needs the model warp, the
    # synchronous hand-over and room for two buffers of 2 * NJP columns
    jtm_ok = XW_ct == 1 and not regv and not asyn and JRP <= 8 and tm_cols + 4 * NP2 * JRP <= 512
    # (measured on C4: 120 k solves/s against 157 k - every tcgen05.cp costs the issuing model warp ~55 clocks, 16 per stage:
    #  the critical path grows by more than the column warps save; kept as an option, off by default)
    jtm = 0
    if 'SBM_CT_JTM' in _os.environ:
        jtm = int(_os.environ['SBM_CT_JTM']) if jtm_ok else 0
    L.append('#define SBM_CT_JTM %d' % jtm)
    # lagged step-size control (csrc/kernels/sbm_dp5_ct.cuh, SBM_CT_LAG): needs the model warp with the whole model in one share
    lag = 1 if (XW_ct == 1 and n_sh == 1 and not regv and not asyn and not jtm) else 0
    if 'SBM_CT_LAG' in _os.environ:
        lag = int(_os.environ['SBM_CT_LAG']) if lag else 0
    L.append('#define SBM_CT_LAG %d' % lag)
    L.append('#if SBM_CT_JTM')
    L.append('#define SBM_JPOS(q, qp) ((((qp) >> 1) * 64) + ((qp) & 1))')
    L.append('#else')
    L.append('#define SBM_JPOS(q, qp) ((q) * 32)')
    L.append('#endif')
    if W_ct:
        ysyms = [sp.Symbol('y[%d]' % i, real=True) for i in range(N)]
        _emit_primal_shares(L, pr, N, f_x, J_x, G_x, gkeys, W_ct, C_ct, NF, ysyms, XW_ct, jpos=jpos, jslot=jslot, jrep=jrep)
        # a second, wider split for the first stage of a step: its step size is only known once the previous step
        # has been judged, so nothing overlaps it and every warp is free to help
        # (measured on C4 with the model warp: 1 / 2 / 3 / 4 / 6 head shares = 144.2 / 141.7 / 145.1 / 144.5 / 142.9 k solves/s;
        #  final build, (shares, column base load): (3, 150) 157.3, (3, 200) 157.9, (4, 100) 157.7, (4, 200) 158.8, (5, 150) 157.2 k)
        n_head = int(_os.environ.get('SBM_CT_HEAD_SHARES', '0')) or min(W_ct + XW_ct, 4)
        L.append('#define SBM_CT_HEAD_SHARES %d' % n_head)
        _emit_primal_shares(L, pr, N, f_x, J_x, G_x, gkeys, W_ct, C_ct, NF, ysyms, XW_ct, fname='sbm_primal_head',
                            n_shares=n_head, head=True, jpos=jpos, jslot=jslot, jrep=jrep)

    # --- dense Jacobian (row-major Jd[i*N+k]) and df/dt for the Rosenbrock kernel
    rep, red = sp.cse(J_x, symbols=sp.numbered_symbols('c_'), optimizations='basic')
    flops_jd = count_flops([r for _, r in rep] + list(red))
    L.append('__device__ __forceinline__ void sbm_jac_dense(const double t, const double* __restrict__ y,')
    L.append('        const double* __restrict__ p, double* __restrict__ Jd)')
    L.append('{')
    L.append('    #pragma unroll')
    L.append('    for (int q = 0; q < SBM_N*SBM_N; ++q) Jd[q] = 0.0;')
    _emit_block(L, rep, [('Jd[%d]' % (i * N + k), red[q]) for q, (i, k) in enumerate(jkeys)], pr)
    L.append('}')
    L.append('')
    dfdt_x = [X(e) for e in sm.dfdt]
    rep, red = sp.cse(dfdt_x, symbols=sp.numbered_symbols('d_'), optimizations='basic')
    L.append('__device__ __forceinline__ void sbm_dfdt(const double t, const double* __restrict__ y,')
    L.append('        const double* __restrict__ p, double* __restrict__ ft)')
    L.append('{')
    _emit_block(L, rep, [('ft[%d]' % i, e) for i, e in enumerate(red)], pr)
    L.append('}')
    L.append('')

    # --- directional derivative of the non-zeros of J and G along a state direction u (same sparse layout as sbm_fJG):
    #     DJv[q] = sum_l dJ_q/dy_l u[l],  DGv[q] = sum_l dG_q/dy_l u[l].  The Rosenbrock kernel with sensitivities needs it: the
    #     Jacobian of the augmented system [y | S_j] has the off-diagonal blocks B_j = d(J S_j + G_j)/dy, and a Rosenbrock
    #     method keeps its order only with the exact Jacobian (B_j u = DJ(u) S_j + DG(u)[:, j]).
    if sm.autonomous and N <= 12 and NF >= 1:
        usyms = [sp.Symbol('u[%d]' % l, real=True) for l in range(N)]

        def ddir(e):
            e = sp.sympify(e)
            return sum((sp.diff(e, v) * usyms[l] for l, v in enumerate(sm.var_syms) if e.has(v)), sp.Integer(0))
        dJ_x = [X(ddir(sm.J[k])) for k in jkeys]
        dG_x = [X(ddir(sm.G[k])) for k in gkeys]
        rep, red = sp.cse(dJ_x + dG_x, symbols=sp.numbered_symbols('h_'), optimizations='basic')
        outs = [('DJv[%d]' % q, red[q]) for q in range(len(jkeys))]
        outs += [('DGv[%d]' % q, red[len(jkeys) + q]) for q in range(len(gkeys))]
        L.append('#define SBM_HAS_DJG 1')
        L.append('__device__ __forceinline__ void sbm_dJG(const double t, const double* __restrict__ y,')
        L.append('        const double* __restrict__ p, const double* __restrict__ u,')
        L.append('        double* __restrict__ DJv, double* __restrict__ DGv)')
        L.append('{')
        _emit_block(L, rep, outs, pr)
        L.append('}')
        L.append('')

    src = '\n'.join(L)
    flops_sens = flops_fJG + NF * flops_col + len(gkeys)
    return CudaModelSource(source=src, n_vars=N, n_params=NP, n_free=NF, free_idx=list(sm.free_idx),
                           nnz_J=len(jkeys), nnz_G=len(gkeys), flops_f=flops_f, flops_fJG=flops_fJG,
                           flops_col=flops_col, flops_sens_rhs=flops_sens, flops_jac_dense=flops_jd,
                           autonomous=sm.autonomous, fingerprint=sm.fingerprint())
