"""Coefficients of the polynomial used by the device fast-exp (csrc/fastmath.cuh): e^r on [-ln2/2, ln2/2]
(argument reduced Cody-Waite style, x = k ln2 + r).
Chebyshev interpolation in 60-digit arithmetic (near-minimax), rounded to fp64, then the Horner evaluation
is emulated in fp64 with fused multiply-adds and compared with mpmath to report the worst relative error."""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 60
DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 11


def cheb_coeffs(f, a, b, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    ys = [f(x) for x in xs]
    # solve the Vandermonde system in high precision (monomial basis around 0)
    A = mp.matrix(n, n)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
    c = mp.lu_solve(A, mp.matrix(ys))
    return [c[j] for j in range(n)]


HALF = mp.log(2) / 2 * mp.mpf('1.0001')
coef = cheb_coeffs(lambda x: mp.exp(x), -HALF, HALF, DEG)
c64 = [float(c) for c in coef]
c64[0] = 1.0
print("static const double EXP_C[%d] = {" % (DEG + 1))
for c in c64:
    print("    %s," % float.hex(c), " // %.17g" % c)
print("};")

# fp64 emulation of Horner with FMA (use mpmath rounding to emulate fma exactly)
def fma(a, b, c):
    return float(mp.mpf(a) * mp.mpf(b) + mp.mpf(c))

rng = np.random.default_rng(0)
worst = 0.0
H = float(mp.log(2) / 2)
for r in np.concatenate([rng.uniform(-H, H, 20000), [-H, H, 0.0, 1e-300, -1e-17]]):
    p = c64[DEG]
    for k in range(DEG - 1, -1, -1):
        p = fma(p, float(r), c64[k])
    exact = mp.exp(mp.mpf(float(r)))
    worst = max(worst, abs((mp.mpf(p) - exact) / exact))
print("degree", DEG, "worst relative error %.3e (ulp = 1.1e-16 .. 2.2e-16)" % float(worst))
