"""Generate the polynomial used by rsk_atan2 (pyradiosky_b200/csrc/rsk_device.cuh).

atan(t) = t + t*s*P(s), s = t*t, t in [0, 1];  P fitted with mpmath.chebyfit at 60 digits and rounded to
double.  Prints the coefficients (lowest order first) and the maximum error of the double-precision Horner
evaluation against mpmath over a dense grid.
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60
DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 21


def g(s):
    if s == 0:
        return mp.mpf(-1) / 3
    t = mp.sqrt(s)
    return (mp.atan(t) / t - 1) / s


coef_hi_first = mp.chebyfit(g, [0, 1], DEG + 1)
coef = [float(c) for c in coef_hi_first[::-1]]  # lowest first
ts = np.linspace(0.0, 1.0, 200001)
s = ts * ts
p = np.zeros_like(s)
for c in coef[::-1]:
    p = p * s + c
approx = ts + ts * s * p
exact = np.array([float(mp.atan(mp.mpf(float(t)))) for t in ts[::50]])
err = np.abs(approx[::50] - exact)
print("degree", DEG, "max abs err", err.max(), "max rel err", (err[1:] / exact[1:]).max())
print("static const double kAtanP[%d] = {" % len(coef))
for c in coef:
    print("    %s," % float(c).hex(), " //", repr(c))
print("};")
