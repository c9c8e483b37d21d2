"""One-state synthesis/degradation model dy/dt = k_synt - k_deg*y (BASELINE config C1).

Same equations and parameter order as the reference fixture ``tests/test_utils/simple_model.py:9-22``.
"""
n_vars = 1
initial_conditions = [0.0]


def simple_synth_deg(y, t, p):
    #*! Parameters Start
    k_deg = p[0]
    k_synt = p[1]
    #*! Parameters End

    #*! Variables Start
    _y = y[0]
    #*! Variables End

    #*! Differential Equations Start
    d_y = k_synt - k_deg * _y
    #*! Differential Equations End
    return [d_y]
