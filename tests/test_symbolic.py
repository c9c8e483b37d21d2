"""CPU: the symbolic front end (reference symbolic/sympy_tools.py) - parsing, sensitivity derivation, the
Python emitter in the reference's generated-file layout and the CUDA emitter's structure and op counts.
Intent of the reference's (stale) tests/test_sympy_tools.py: generated model == hand model under odeint."""
import numpy as np
import sympy as sp
from scipy.integrate import odeint

from sysbio_modeling_b200 import model_files
from sysbio_modeling_b200.symbolic import (parse_model_file, process_model_dict, symbolic_model_from_file,
                                           trace_rhs_callable, emit_python, emit_cuda, make_ode_model, count_flops)


def test_parse_sections_and_name_collisions():
    d = parse_model_file(model_files.path('reversible_mm'))
    assert list(d['Parameters']) == ['k1f', 'k1r', 'k2f', 'k2r']
    assert list(d['Variables']) == ['S', 'E', 'ES', 'P']          # S, E are SymPy globals: must stay Symbols
    assert all(isinstance(v, sp.Symbol) for v in d['Variables'].values())
    assert set(d['Rate Laws']) == {'r1', 'r2'}
    sm = symbolic_model_from_file(model_files.path('reversible_mm'))
    S, E, ES, P = sm.var_syms
    k1f, k1r, k2f, k2r = sm.param_syms
    assert sp.simplify(sm.f[0] - (-(k1f * E * S - k1r * ES))) == 0
    assert sm.n_free == 4 and len(sm.J) == 14 and sm.autonomous


def test_generated_python_matches_hand_written_rhs():
    sm = symbolic_model_from_file(model_files.path('michaelis_menten'))
    f = make_ode_model(sm, sensitivities=False)
    p = np.array([1e-3, 0.001, 0.01, 0.01, 0.001])

    def hand(y, t):
        v = p[0] * y[0] / (p[1] + y[0])
        return [p[2] - v - p[3] * y[0], v - p[4] * y[1]]
    out = np.zeros(2)

    def gen(y, t):
        f(y, t, out, p)
        return out.copy()
    t = np.linspace(0, 100, 30)
    assert np.allclose(odeint(hand, [0, 0], t, rtol=1e-10, atol=1e-12),
                       odeint(gen, [0, 0], t, rtol=1e-10, atol=1e-12), rtol=1e-7, atol=1e-12)


def test_sensitivity_layout_and_values_vs_finite_differences():
    """yout[N + i*P + j] = d/dt (dx_i/dp_j) (reference sympy_tools.py:135-145 ordering)."""
    sm = symbolic_model_from_file(model_files.path('michaelis_menten'))
    s = make_ode_model(sm, sensitivities=True)
    assert 'yout[11]' in s.source and 'sens_s_vmax = y[2]' in s.source and 'sens_p_k_deg_p = y[11]' in s.source
    p = np.array([1e-3, 0.002, 0.01, 0.01, 0.001])
    out = np.zeros(12)

    def rhs(y, t):
        s(y, t, out, p)
        return out.copy()
    t = np.linspace(0, 50, 6)
    sol = odeint(rhs, np.zeros(12), t, rtol=1e-11, atol=1e-13)
    f = make_ode_model(sm, sensitivities=False)
    o2 = np.zeros(2)
    for j in range(5):
        def run(pp):
            def r(y, tt):
                f(y, tt, o2, pp)
                return o2.copy()
            return odeint(r, np.zeros(2), t, rtol=1e-11, atol=1e-13)
        h = 1e-3 * p[j]
        pp, pm = p.copy(), p.copy()
        pp[j] += h
        pm[j] -= h
        fd = (run(pp) - run(pm)) / (2 * h)
        for i in range(2):
            assert np.allclose(sol[:, 2 + i * 5 + j], fd[:, i], rtol=1e-3, atol=1e-5)


def test_fixed_parameters_drop_columns():
    sm = symbolic_model_from_file(model_files.path('michaelis_menten'), fixed_params=['km', 'k_deg_p'])
    assert sm.free == ['vmax', 'k_synt_s', 'k_deg_s'] and sm.free_idx == [0, 2, 3]
    cs = emit_cuda(sm)
    assert cs.n_free == 3 and '#define SBM_NF 3' in cs.source and 'sbm_free_idx[3] = {0, 2, 3}' in cs.source
    d = process_model_dict(parse_model_file(model_files.path('michaelis_menten')), fixed_params=['km'])
    assert d['Parameters']['km'] == 'fixed' and len(d['Sensitivity Equations']) == 8
    assert list(d['Sensitivity Equations'])[0] == 'd_sens_s_vmax'


def test_trace_callable_equals_model_file():
    sm = symbolic_model_from_file(model_files.path('michaelis_menten'))
    f = make_ode_model(sm, False)
    tr = trace_rhs_callable(f, 2, sm.param_order)
    sub = dict(zip(tr.var_syms, sm.var_syms))
    sub.update(dict(zip(tr.param_syms, sm.param_syms)))
    for a, b in zip(tr.f, sm.f):
        assert sp.simplify(a.xreplace(sub) - b) == 0
    assert tr.fingerprint() != sm.fingerprint()          # different symbol names, same maths


def test_cuda_emitter_structure_and_flop_model():
    sm = symbolic_model_from_file(model_files.path('mapk_kholodenko'))
    cs = emit_cuda(sm)
    assert (cs.n_vars, cs.n_params, cs.n_free, cs.nnz_J, cs.nnz_G) == (8, 22, 22, 26, 44)
    for fn in ('sbm_f', 'sbm_fJG', 'sbm_col', 'sbm_jac_dense', 'sbm_dfdt'):
        assert 'void %s(' % fn in cs.source
    # one reciprocal per distinct denominator: 10 Michaelis-Menten terms + Ki + Hill denominator + 1/MAPK_PP
    assert cs.source.split('void sbm_fJG')[1].split('void sbm_col')[0].count('sbm_rcp(') == 13
    # algorithmic cost of one augmented RHS = f,J,G once + 22 x (2 flops per Jacobian non-zero) + 44 adds
    assert cs.flops_col == 2 * 26
    assert cs.flops_sens_rhs == cs.flops_fJG + 22 * cs.flops_col + 44
    assert 1200 < cs.flops_sens_rhs < 1500
    assert count_flops([sp.sympify('a*b + c')]) == 2 and count_flops([sp.sympify('a**3/b')]) == 4


def test_ct_layout_and_primal_shares_cover_every_output_once():
    """The CTA-transposed kernel splits the model's primal part (f, nnz J, nnz G) into per-warp shares: every output
    must be written by exactly one share, into the slot the kernel reads it from."""
    import re
    from sysbio_modeling_b200 import model_files
    from sysbio_modeling_b200.symbolic import symbolic_model_from_file, emit_cuda
    from sysbio_modeling_b200.symbolic.sympy_tools import ct_layout, ct_share_warps
    assert ct_layout(8, 22) == (11, 2)          # MAPK: 11 warps x 2 columns
    assert ct_layout(4, 4) == (1, 4)            # small systems: one thread per trajectory
    assert ct_layout(25, 10) == (0, 0)          # does not fit: warp-group kernel
    assert ct_layout(8, 22, cols=3) == (8, 3)
    assert sorted(ct_share_warps(11, 3)) == [3, 7, 10]
    for name in ('mapk_kholodenko', 'reversible_mm', 'michaelis_menten'):
        sm = symbolic_model_from_file(model_files.path(name), name=name)
        cs = emit_cuda(sm)
        src = cs.source
        W = int(re.search(r'#define SBM_CT_W (\d+)', src).group(1))
        C = int(re.search(r'#define SBM_CT_C (\d+)', src).group(1))
        assert (W, C) == ct_layout(sm.n_vars, sm.n_free)
        body = src[src.index('void sbm_primal_share('):]
        body = body[:body.index('\n}\n')]
        f_idx = sorted(int(x) for x in re.findall(r'fo\[(\d+) \* 32\] =', body))
        # J entries equal up to sign share a slot of the J buffer (SBM_NJU slots); the first argument of SBM_JPOS is the
        # slot, written once - the duplicates only exist for the padded tensor-memory layout (#if SBM_CT_JTM)
        plain = re.sub(r'#if SBM_CT_JTM\n.*\n#endif\n', '', body)
        j_idx = sorted(int(x) for x in re.findall(r'Jo\[SBM_JPOS\((\d+), \d+\)\] =', plain))
        nju = int(re.search(r'#define SBM_NJU (\d+)', src).group(1))
        assert 1 <= nju <= cs.nnz_J
        g_idx = sorted(int(x) for x in re.findall(r'Go\[SBM_GIDX\((\d+)\)\] =', body))
        assert f_idx == list(range(sm.n_vars))
        assert j_idx == list(range(nju))
        # dense G slot = parameter * N + state, one per structural non-zero
        ne = sm.n_vars + (sm.n_vars & 1)
        assert g_idx == sorted(j * ne + i for (i, j) in sm.G.keys())
        XW = int(re.search(r'#define SBM_CT_XW (\d+)', src).group(1))
        assert body.count('case ') == W + XW


def test_directional_derivative_of_J_and_G_for_the_rosenbrock_sensitivity_kernel():
    """``sbm_dJG`` (DJv[q] = sum_l dJ_q/dy_l u_l, same for G) is emitted for autonomous models of at most 12 states, and its
    formula - re-derived here from the symbolic model - equals a central difference of J and G along the direction."""
    src = emit_cuda(symbolic_model_from_file(model_files.path('robertson'))).source
    assert '#define SBM_HAS_DJG 1' in src and 'sbm_dJG(' in src
    assert 'sbm_dJG(' not in emit_cuda(symbolic_model_from_file(model_files.path('egfr_signalling'))).source   # 25 states
    sm = symbolic_model_from_file(model_files.path('mapk_kholodenko'))
    src = emit_cuda(sm).source
    assert src.count('DJv[') == len(sm.J) and src.count('DGv[') == len(sm.G)
    rng = np.random.default_rng(0)
    mod = model_files.module('mapk_kholodenko')
    yv = np.array(mod.initial_conditions) * (1 + 0.1 * rng.random(8))
    pv = np.array([mod.nominal[k] for k in sm.param_order])
    u = rng.standard_normal(8)
    usyms = sp.symbols('u0:8', real=True)
    args = list(sm.var_syms) + list(sm.param_syms)
    for table in (sm.J, sm.G):
        keys = sorted(table.keys())
        val = sp.lambdify(args, [table[k] for k in keys], 'numpy')
        dirs = [sum(sp.diff(table[k], v) * usyms[l] for l, v in enumerate(sm.var_syms)) for k in keys]
        dval = sp.lambdify(args + list(usyms), dirs, 'numpy')
        eps = 1e-5
        fd = (np.array(val(*(list(yv + eps * u) + list(pv))), dtype=float) -
              np.array(val(*(list(yv - eps * u) + list(pv))), dtype=float)) / (2 * eps)
        an = np.array(dval(*(list(yv) + list(pv) + list(u))), dtype=float)
        assert np.allclose(an, fd, rtol=1e-6, atol=1e-9 * (1 + np.abs(fd).max()))
