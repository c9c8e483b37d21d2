"""Reversible Michaelis-Menten, explicit enzyme complex (BASELINE config C2).

S + E <-> ES <-> E + P with mass-action steps; 4 states, 4 parameters (20 augmented states).
"""
n_vars = 4
initial_conditions = [1.0, 0.1, 0.0, 0.0]
nominal = {'k1f': 10.0, 'k1r': 1.0, 'k2f': 5.0, 'k2r': 0.1}


def reversible_mm(y, t, p):
    #*! Parameters Start
    k1f = p[0]
    k1r = p[1]
    k2f = p[2]
    k2r = p[3]
    #*! Parameters End

    #*! Variables Start
    S = y[0]
    E = y[1]
    ES = y[2]
    P = y[3]
    #*! Variables End

    #*! Rate Laws Start
    r1 = k1f * E * S - k1r * ES
    r2 = k2f * ES - k2r * E * P
    #*! Rate Laws End

    #*! Differential Equations Start
    d_S = -r1
    d_E = -r1 + r2
    d_ES = r1 - r2
    d_P = r2
    #*! Differential Equations End
    return [d_S, d_E, d_ES, d_P]
