"""EGFR-scale signalling network (BASELINE config C5): receptor activation, Shc / Grb2 / Sos
recruitment, Ras activation with ERK -> Sos negative feedback, and the Raf / MEK / ERK cascade.

25 species, 52 parameters, mass-action binding steps and Michaelis-Menten enzymatic steps (a hand-written network
in the spirit of the published EGFR models; the reference ships no model of this size).  Used for the ensemble /
SSR-only workload: many parameter sets ("walkers"), residuals only.
"""
n_vars = 25
initial_conditions = [100.0, 0.0, 0.0, 0.0, 0.0,      # R, RL, R2, R2P, Ri
                      80.0, 0.0, 40.0, 0.0, 0.0,      # G, R2PG, S, R2PGS, GS
                      0.0, 60.0, 0.0, 0.0, 0.0,       # SosI, Shc, R2PShc, ShcP, ShcPG
                      120.0, 0.0,                     # RasD, RasT
                      100.0, 0.0,                     # Raf, RafP
                      200.0, 0.0, 0.0,                # MEK, MEKP, MEKPP
                      200.0, 0.0, 0.0]                # ERK, ERKP, ERKPP
nominal = {'L': 1.0, 'kf1': 0.05, 'kr1': 0.06, 'kf2': 0.01, 'kr2': 0.1, 'k3': 1.0, 'V3p': 4.0, 'K3p': 50.0,
           'kf4': 0.003, 'kr4': 0.05, 'kf5': 0.01, 'kr5': 0.06, 'kf6': 0.0015, 'kr6': 0.03, 'kf7': 0.009, 'kr7': 0.04,
           'k8': 0.5, 'K8': 50.0, 'V9': 5.0, 'K9': 30.0, 'k10': 0.05, 'K10': 60.0, 'V11': 3.0, 'K11': 20.0,
           'k12': 0.15, 'K12': 100.0, 'k13': 0.15, 'K13': 100.0, 'V14': 4.0, 'K14': 60.0, 'V15': 4.0, 'K15': 60.0,
           'k16': 0.3, 'K16': 100.0, 'k17': 0.3, 'K17': 100.0, 'V18': 3.0, 'K18': 60.0, 'V19': 3.0, 'K19': 60.0,
           'k20': 0.02, 'K20': 40.0, 'k21': 0.01, 'kf22': 0.002, 'kr22': 0.05, 'k23': 0.2, 'kf24': 0.003,
           'kr24': 0.1, 'V25': 2.0, 'K25': 40.0, 'k26': 0.005, 'k27': 0.002}


def egfr_signalling(y, t, p):
    #*! Parameters Start
    L = p[0]
    kf1 = p[1]
    kr1 = p[2]
    kf2 = p[3]
    kr2 = p[4]
    k3 = p[5]
    V3p = p[6]
    K3p = p[7]
    kf4 = p[8]
    kr4 = p[9]
    kf5 = p[10]
    kr5 = p[11]
    kf6 = p[12]
    kr6 = p[13]
    kf7 = p[14]
    kr7 = p[15]
    k8 = p[16]
    K8 = p[17]
    V9 = p[18]
    K9 = p[19]
    k10 = p[20]
    K10 = p[21]
    V11 = p[22]
    K11 = p[23]
    k12 = p[24]
    K12 = p[25]
    k13 = p[26]
    K13 = p[27]
    V14 = p[28]
    K14 = p[29]
    V15 = p[30]
    K15 = p[31]
    k16 = p[32]
    K16 = p[33]
    k17 = p[34]
    K17 = p[35]
    V18 = p[36]
    K18 = p[37]
    V19 = p[38]
    K19 = p[39]
    k20 = p[40]
    K20 = p[41]
    k21 = p[42]
    kf22 = p[43]
    kr22 = p[44]
    k23 = p[45]
    kf24 = p[46]
    kr24 = p[47]
    V25 = p[48]
    K25 = p[49]
    k26 = p[50]
    k27 = p[51]
    #*! Parameters End

    #*! Variables Start
    R = y[0]
    RL = y[1]
    R2 = y[2]
    R2P = y[3]
    Ri = y[4]
    G = y[5]
    R2PG = y[6]
    S = y[7]
    R2PGS = y[8]
    GS = y[9]
    SosI = y[10]
    Shc = y[11]
    R2PShc = y[12]
    ShcP = y[13]
    ShcPG = y[14]
    RasD = y[15]
    RasT = y[16]
    Raf = y[17]
    RafP = y[18]
    MEK = y[19]
    MEKP = y[20]
    MEKPP = y[21]
    ERK = y[22]
    ERKP = y[23]
    ERKPP = y[24]
    #*! Variables End

    #*! Rate Laws Start
    v1 = kf1 * L * R - kr1 * RL
    v2 = kf2 * RL * RL - kr2 * R2
    v3 = k3 * R2 - V3p * R2P / (K3p + R2P)
    v4 = kf4 * R2P * G - kr4 * R2PG
    v5 = kf5 * R2PG * S - kr5 * R2PGS
    v6 = kf6 * G * S - kr6 * GS
    v7 = kf7 * R2P * GS - kr7 * R2PGS
    v8 = k8 * R2PGS * RasD / (K8 + RasD)
    v9 = V9 * RasT / (K9 + RasT)
    v10 = k10 * RasT * Raf / (K10 + Raf)
    v11 = V11 * RafP / (K11 + RafP)
    v12 = k12 * RafP * MEK / (K12 + MEK)
    v13 = k13 * RafP * MEKP / (K13 + MEKP)
    v14 = V14 * MEKPP / (K14 + MEKPP)
    v15 = V15 * MEKP / (K15 + MEKP)
    v16 = k16 * MEKPP * ERK / (K16 + ERK)
    v17 = k17 * MEKPP * ERKP / (K17 + ERKP)
    v18 = V18 * ERKPP / (K18 + ERKPP)
    v19 = V19 * ERKP / (K19 + ERKP)
    v20 = k20 * ERKPP * S / (K20 + S)
    v21 = k21 * SosI
    v22 = kf22 * R2P * Shc - kr22 * R2PShc
    v23 = k23 * R2PShc
    v24 = kf24 * ShcP * G - kr24 * ShcPG
    v25 = V25 * ShcP / (K25 + ShcP)
    v26 = k26 * R2P - k27 * Ri
    #*! Rate Laws End

    #*! Differential Equations Start
    d_R = -v1
    d_RL = v1 - 2 * v2
    d_R2 = v2 - v3
    d_R2P = v3 - v4 - v7 - v22 + v23 - v26
    d_Ri = v26
    d_G = -v4 - v6 - v24
    d_R2PG = v4 - v5
    d_S = -v5 - v6 - v20 + v21
    d_R2PGS = v5 + v7
    d_GS = v6 - v7
    d_SosI = v20 - v21
    d_Shc = -v22 + v25
    d_R2PShc = v22 - v23
    d_ShcP = v23 - v24 - v25
    d_ShcPG = v24
    d_RasD = v9 - v8
    d_RasT = v8 - v9
    d_Raf = v11 - v10
    d_RafP = v10 - v11
    d_MEK = v15 - v12
    d_MEKP = v12 - v13 + v14 - v15
    d_MEKPP = v13 - v14
    d_ERK = v19 - v16
    d_ERKP = v16 - v17 + v18 - v19
    d_ERKPP = v17 - v18
    #*! Differential Equations End
    return [d_R, d_RL, d_R2, d_R2P, d_Ri, d_G, d_R2PG, d_S, d_R2PGS, d_GS, d_SosI, d_Shc, d_R2PShc, d_ShcP,
            d_ShcPG, d_RasD, d_RasT, d_Raf, d_RafP, d_MEK, d_MEKP, d_MEKPP, d_ERK, d_ERKP, d_ERKPP]
