"""Kholodenko (2000) MAPK cascade with negative feedback (BASELINE config C4).

8 states, 22 parameters -> 8 + 8*22 = 184 sensitivity-augmented states.
"""
n_vars = 8
initial_conditions = [90.0, 10.0, 280.0, 10.0, 10.0, 280.0, 10.0, 10.0]
nominal = {'V1': 2.5, 'n': 1.0, 'Ki': 9.0, 'K1': 10.0, 'V2': 0.25, 'K2': 8.0, 'k3': 0.025, 'K3': 15.0,
           'k4': 0.025, 'K4': 15.0, 'V5': 0.75, 'K5': 15.0, 'V6': 0.75, 'K6': 15.0, 'k7': 0.025,
           'K7': 15.0, 'k8': 0.025, 'K8': 15.0, 'V9': 0.5, 'K9': 15.0, 'V10': 0.5, 'K10': 15.0}


def mapk_kholodenko(y, t, p):
    #*! Parameters Start
    V1 = p[0]
    n = p[1]
    Ki = p[2]
    K1 = p[3]
    V2 = p[4]
    K2 = p[5]
    k3 = p[6]
    K3 = p[7]
    k4 = p[8]
    K4 = p[9]
    V5 = p[10]
    K5 = p[11]
    V6 = p[12]
    K6 = p[13]
    k7 = p[14]
    K7 = p[15]
    k8 = p[16]
    K8 = p[17]
    V9 = p[18]
    K9 = p[19]
    V10 = p[20]
    K10 = p[21]
    #*! Parameters End

    #*! Variables Start
    MKKK = y[0]
    MKKK_P = y[1]
    MKK = y[2]
    MKK_P = y[3]
    MKK_PP = y[4]
    MAPK = y[5]
    MAPK_P = y[6]
    MAPK_PP = y[7]
    #*! Variables End

    #*! Rate Laws Start
    v1 = V1 * MKKK / ((1 + (MAPK_PP / Ki) ** n) * (K1 + MKKK))
    v2 = V2 * MKKK_P / (K2 + MKKK_P)
    v3 = k3 * MKKK_P * MKK / (K3 + MKK)
    v4 = k4 * MKKK_P * MKK_P / (K4 + MKK_P)
    v5 = V5 * MKK_PP / (K5 + MKK_PP)
    v6 = V6 * MKK_P / (K6 + MKK_P)
    v7 = k7 * MKK_PP * MAPK / (K7 + MAPK)
    v8 = k8 * MKK_PP * MAPK_P / (K8 + MAPK_P)
    v9 = V9 * MAPK_PP / (K9 + MAPK_PP)
    v10 = V10 * MAPK_P / (K10 + MAPK_P)
    #*! Rate Laws End

    #*! Differential Equations Start
    d_MKKK = v2 - v1
    d_MKKK_P = v1 - v2
    d_MKK = v6 - v3
    d_MKK_P = v3 + v5 - v4 - v6
    d_MKK_PP = v4 - v5
    d_MAPK = v10 - v7
    d_MAPK_P = v7 + v9 - v8 - v10
    d_MAPK_PP = v8 - v9
    #*! Differential Equations End
    return [d_MKKK, d_MKKK_P, d_MKK, d_MKK_P, d_MKK_PP, d_MAPK, d_MAPK_P, d_MAPK_PP]
