"""Robertson's stiff chemical kinetics (BASELINE config C3)."""
n_vars = 3
initial_conditions = [1.0, 0.0, 0.0]
nominal = {'a': 0.04, 'b': 1.0e4, 'c': 3.0e7}


def robertson(y, t, p):
    #*! Parameters Start
    a = p[0]
    b = p[1]
    c = p[2]
    #*! Parameters End

    #*! Variables Start
    y1 = y[0]
    y2 = y[1]
    y3 = y[2]
    #*! Variables End

    #*! Differential Equations Start
    d_y1 = -a * y1 + b * y2 * y3
    d_y2 = a * y1 - b * y2 * y3 - c * y2 * y2
    d_y3 = c * y2 * y2
    #*! Differential Equations End
    return [d_y1, d_y2, d_y3]
