"""Tables of the sweep-grade angle routines in pyradiosky_b200/csrc/rsk_device.cuh (rsk_azimuth_alt_sweep):
atan(k/16), 1/sqrt(1 + (k/16)^2) and the octant table of the azimuth, 40-digit mpmath rounded to double."""
import mpmath as mp

mp.mp.dps = 40
N = 16


def hx(x):
    return float(x).hex()


atan_tab = [(mp.atan(mp.mpf(k) / N), mp.mpf(k) / N) for k in range(N + 1)]
ck = [1 / mp.sqrt(1 + (mp.mpf(k) / N) ** 2) for k in range(N + 1)]
print("__device__ const double2 kRskAtanTab16[%d] = {" % (N + 1 + 15))
print("    " + ", ".join("{%s, %s}" % (hx(a), hx(t)) for a, t in atan_tab) + ",")
print("    " + ", ".join("{0, 0}" for _ in range(15)) + "};")
print("__device__ const double kRskCk16[%d] = {" % (N + 1 + 15))
print("    " + ", ".join(hx(c) for c in ck) + ",")
print("    " + ", ".join("0" for _ in range(15)) + "};")
print("__device__ const double2 kRskAzTab16[%d] = {" % (8 * (N + 1) + 8))
for octant in range(8):
    ys, xs, sw = (octant >> 2) & 1, (octant >> 1) & 1, octant & 1
    cells = []
    for k in range(N + 1):
        base, sgn = mp.atan(mp.mpf(k) / N), 1
        if sw:
            base, sgn = mp.pi / 2 - base, -sgn
        if xs:
            base, sgn = mp.pi - base, -sgn
        if ys:
            base, sgn = -base, -sgn
        if base < 0:
            base += 2 * mp.pi
        if ys and base == 0:
            base = 2 * mp.pi  # like eraAnp: an angle just below zero wraps to 2 pi
        t = mp.mpf(k) / N
        tail = ("0x0.0p+0" if sgn > 0 else "-0x0.0p+0") if k == 0 else hx(t if sgn > 0 else -t)
        cells.append("{%s, %s}" % (hx(base), tail))
    print("    " + ", ".join(cells) + ",  // octant %d" % octant)
print("    " + ", ".join("{0, 0}" for _ in range(8)) + "};")
