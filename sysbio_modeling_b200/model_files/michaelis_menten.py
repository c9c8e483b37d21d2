"""Two-state irreversible Michaelis-Menten with substrate synthesis/degradation (C1, 'sum' mapping).

Same equations and parameter order as the reference fixture
``tests/test_utils/michelis_menten_model.py:13-32``.
"""
n_vars = 2
initial_conditions = [0.0, 0.0]


def michaelis_menten(y, t, p):
    #*! Parameters Start
    vmax = p[0]
    km = p[1]
    k_synt_s = p[2]
    k_deg_s = p[3]
    k_deg_p = p[4]
    #*! Parameters End

    #*! Variables Start
    _s = y[0]
    _p = y[1]
    #*! Variables End

    #*! Rate Laws Start
    v_cat = vmax * _s / (km + _s)
    #*! Rate Laws End

    #*! Differential Equations Start
    d_s = k_synt_s - v_cat - k_deg_s * _s
    d_p = v_cat - k_deg_p * _p
    #*! Differential Equations End
    return [d_s, d_p]
