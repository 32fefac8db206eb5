"""Second, independent restatement of the carbonate system solved by co2calc.F90.

TEST INFRASTRUCTURE ONLY.  Where the C oracle (co2calc_oracle.c) follows the
reference line by line, this file deliberately does NOT: it writes total
alkalinity from the textbook speciation (DOE 1994 / Dickson, Sabine & Christian
2007, ch. 2) and finds the root with plain bisection to machine precision.  It
shares no algebra with talk_row (co2calc.F90:1001-1092) beyond the equilibrium
constants, so agreement between the two pins the residual function and the
Newton/bisection solver of the oracle.

Equilibrium constants: the same published fits the reference uses
(co2calc.F90:320-777), vectorised, without the level-index quirk (pass
`deep` explicitly).
"""
import numpy as np

T0_KELVIN = 273.15
RHO_SW = 1.026
SALT_MIN = 0.1


def press_bar(depth_m):
    return 0.059808 * (np.exp(-0.025 * depth_m) - 1.0) + 0.100766 * depth_m + 2.28405e-7 * depth_m ** 2


def constants(temp, salt, depth_m=0.0, deep=False):
    """dict of k1,k2,kb,k1p,k2p,k3p,ksi,kw,ks,kf,bt,st,ft (mol/kg scales as in the reference)."""
    t = np.asarray(temp, dtype=np.float64)
    s = np.maximum(np.asarray(salt, dtype=np.float64), SALT_MIN)
    tk = T0_KELVIN + t
    lntk = np.log(tk)
    itk = 1.0 / tk
    sq = np.sqrt(s)
    ion = 19.924 * s / (1000.0 - 1.005 * s)
    sqi = np.sqrt(ion)
    scl = s / 1.80655
    ln1m = np.log(1.0 - 0.001005 * s)
    P = press_bar(np.asarray(depth_m, dtype=np.float64))
    RT = 83.1451 * tk

    def pcorr(dV, kappa):
        if not deep:
            return 1.0
        return np.exp((-dV + 0.5 * kappa * P) * P / RT)

    out = {}
    # Lueker et al. 2000 (total scale); NOT pressure corrected in the reference (quirk Q2)
    out["k1"] = 10.0 ** (-(3633.86 * itk - 61.2172 + 9.67770 * lntk - 0.011555 * s + 0.0001152 * s * s))
    out["k2"] = 10.0 ** (-(471.78 * itk + 25.9290 - 3.16967 * lntk - 0.01781 * s + 0.0001122 * s * s))
    out["kb"] = np.exp((-8966.90 - 2890.53 * sq - 77.942 * s + 1.728 * s * sq - 0.0996 * s * s) * itk
                       + (148.0248 + 137.1942 * sq + 1.62142 * s)
                       + (-24.4344 - 25.085 * sq - 0.2474 * s) * lntk + 0.053105 * sq * tk) \
        * pcorr(-29.48 + (0.1622 - 0.002608 * t) * t, -2.84e-3)
    out["k1p"] = np.exp(-4576.752 * itk + 115.525 - 18.453 * lntk + (-106.736 * itk + 0.69171) * sq
                        + (-0.65643 * itk - 0.01844) * s) \
        * pcorr(-14.51 + (0.1211 - 0.000321 * t) * t, (-2.67 + 0.0427 * t) * 1e-3)
    out["k2p"] = np.exp(-8814.715 * itk + 172.0883 - 27.927 * lntk + (-160.340 * itk + 1.3566) * sq
                        + (0.37335 * itk - 0.05778) * s) \
        * pcorr(-23.12 + (0.1758 - 0.002647 * t) * t, (-5.15 + 0.09 * t) * 1e-3)
    out["k3p"] = np.exp(-3070.75 * itk - 18.141 + (17.27039 * itk + 2.81197) * sq
                        + (-44.99486 * itk - 0.09984) * s) \
        * pcorr(-26.57 + (0.202 - 0.003042 * t) * t, (-4.08 + 0.0714 * t) * 1e-3)
    out["ksi"] = np.exp(-8904.2 * itk + 117.385 - 19.334 * lntk + (-458.79 * itk + 3.5913) * sqi
                        + (188.74 * itk - 1.5998) * ion + (-12.1652 * itk + 0.07871) * ion * ion + ln1m) \
        * pcorr(-29.48 + (0.1622 - 0.002608 * t) * t, -2.84e-3)
    out["kw"] = np.exp(-13847.26 * itk + 148.9652 - 23.6521 * lntk
                       + (118.67 * itk - 5.977 + 1.0495 * lntk) * sq - 0.01615 * s) \
        * pcorr(-20.02 + (0.1119 - 0.001409 * t) * t, (-5.13 + 0.0794 * t) * 1e-3)
    out["ks"] = np.exp(-4276.1 * itk + 141.328 - 23.093 * lntk
                       + (-13856.0 * itk + 324.57 - 47.986 * lntk) * sqi
                       + (35474.0 * itk - 771.54 + 114.723 * lntk) * ion
                       - 2698.0 * itk * ion * sqi + 1776.0 * itk * ion * ion + ln1m) \
        * pcorr(-18.03 + (0.0466 + 0.000316 * t) * t, (-4.53 + 0.09 * t) * 1e-3)
    st = 0.14 / 96.062 * scl
    out["kf"] = np.exp(1590.2 * itk - 12.641 + 1.525 * sqi + ln1m + np.log(1.0 + st / out["ks"])) \
        * pcorr(-9.78 - (0.009 + 0.000942 * t) * t, (-3.91 + 0.054 * t) * 1e-3)
    out["bt"] = 0.000232 / 10.811 * scl
    out["st"] = st
    out["ft"] = 0.000067 / 18.9984 * scl
    # Weiss & Price 1980 fugacity-corrected solubility (used by co2calc_1point only)
    tk100 = tk / 100.0
    out["ff"] = np.exp(-162.8301 + 218.2968 / tk100 + 90.9241 * np.log(tk100) - 1.47696 * tk100 ** 2
                       + s * (0.025695 - 0.025225 * tk100 + 0.0049867 * tk100 ** 2))
    return out


def alkalinity_residual(h, K, dic, ta, pt, sit):
    """TA(H) - TA from the species concentrations (mol/kg)."""
    k1, k2 = K["k1"], K["k2"]
    d = h * h + k1 * h + k1 * k2
    hco3 = dic * k1 * h / d
    co3 = dic * k1 * k2 / d
    boh4 = K["bt"] * K["kb"] / (K["kb"] + h)
    oh = K["kw"] / h
    dp = h ** 3 + K["k1p"] * h * h + K["k1p"] * K["k2p"] * h + K["k1p"] * K["k2p"] * K["k3p"]
    h3po4 = pt * h ** 3 / dp
    hpo4 = pt * K["k1p"] * K["k2p"] * h / dp
    po4 = pt * K["k1p"] * K["k2p"] * K["k3p"] / dp
    sioh3 = sit * K["ksi"] / (K["ksi"] + h)
    # total -> free scale for the sulfate / fluoride / free-proton terms
    z = 1.0 + K["st"] / K["ks"]
    hfree = h / z
    hso4 = K["st"] / (1.0 + K["ks"] / hfree)
    hf = K["ft"] / (1.0 + K["kf"] / h)
    return hco3 + 2.0 * co3 + boh4 + oh + hpo4 + 2.0 * po4 + sioh3 - hfree - hso4 - hf - h3po4 - ta


def solve_h(temp, salt, dic_mmol, ta_mmol, pt_mmol, sit_mmol, depth_m=0.0, deep=False, ph_lo=2.0, ph_hi=12.0):
    """[H+] (mol/kg, total scale) by bisection in log space to machine precision."""
    K = constants(temp, salt, depth_m, deep)
    v2m = 1.0 / (1e6 * RHO_SW)
    dic = np.maximum(dic_mmol, SALT_MIN / 35.0 * 1944.0) * v2m
    ta = np.maximum(ta_mmol, SALT_MIN / 35.0 * 2225.0) * v2m
    pt = np.maximum(pt_mmol, 0.0) * v2m
    sit = np.maximum(sit_mmol, 0.0) * v2m
    lo = np.full(np.shape(dic), -ph_hi, dtype=np.float64)   # log10 H
    hi = np.full(np.shape(dic), -ph_lo, dtype=np.float64)
    flo = alkalinity_residual(10.0 ** lo, K, dic, ta, pt, sit)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        fm = alkalinity_residual(10.0 ** mid, K, dic, ta, pt, sit)
        same = np.sign(fm) == np.sign(flo)
        lo = np.where(same, mid, lo)
        flo = np.where(same, fm, flo)
        hi = np.where(same, hi, mid)
    return 10.0 ** (0.5 * (lo + hi)), K, dic
