"""Generate the fp32 polynomial coefficients of the canonical math spec ("Math").

Least-squares fit on Chebyshev nodes (near-minimax) with the target evaluated in 60-digit
arithmetic (mpmath), coefficients rounded to fp32 and printed as hex-float literals.  The
output is pasted into oracle/pz_oracle.cpp and probzelus-haskell_b200/csrc/pz_math.cuh;
both sides must carry IDENTICAL literals (tests/test_oracle.py checks the oracle's
accuracy against libm; the GPU parity tests check bit-equality with the oracle).
"""
import numpy as np
import mpmath as mp

mp.mp.dps = 60

def cheb_nodes(a, b, n=801):
    k = np.arange(n)
    x = np.cos(np.pi * (k + 0.5) / n)
    return 0.5 * (a + b) + 0.5 * (b - a) * x

def fit(fun, a, b, ncoef):
    x = cheb_nodes(a, b)
    y = np.array([float(fun(mp.mpf(float(v)))) for v in x])
    A = np.stack([x ** p for p in range(ncoef)], axis=1)
    c, *_ = np.linalg.lstsq(A, y, rcond=None)
    return np.float32(c)

def show(name, c):
    print(name, "= {" + ", ".join(float(v).hex().replace("0000000p", "p") + "f" for v in c) + "}")

def horner(c, x):
    p = np.float32(0) * x
    for v in c[::-1]:
        p = np.float32(p * x + v)
    return p

# 2^f = 1 + f*E(f) on [-0.5, 0.5]
cE = fit(lambda f: mp.log(2) if f == 0 else (mp.power(2, f) - 1) / f, -0.5, 0.5, 7)
show("EXP2", cE)
f = np.linspace(-0.5, 0.5, 200001).astype(np.float32)
p = np.float32(np.float32(horner(cE, f) * f) + np.float32(1))
print("   exp2 max rel err", np.max(np.abs(p.astype(np.float64) / 2.0 ** f.astype(np.float64) - 1)))

# ln(1+t) = t - t^2/2 + t^3 * P(t), t in [sqrt(.5)-1, sqrt(2)-1]
a, b = np.sqrt(0.5) - 1, np.sqrt(2) - 1
cL = fit(lambda t: mp.mpf(1) / 3 if t == 0 else (mp.log1p(t) - t + t * t / 2) / t ** 3, a, b, 8)
show("LOG", cL)
t = np.linspace(a, b, 200001).astype(np.float32)
t2 = np.float32(t * t)
r = np.float32(np.float32(t - np.float32(0.5) * t2) + np.float32(np.float32(t2 * t) * horner(cL, t)))
print("   log max abs err", np.max(np.abs(r.astype(np.float64) - np.log1p(t.astype(np.float64)))))

# sin(x) = x + x^3 * S(x^2), cos(x) = 1 - x^2/2 + x^4 * C(x^2) on [-pi/4, pi/4]
q = np.pi / 4
cS = fit(lambda z: -mp.mpf(1) / 6 if z == 0 else (mp.sin(mp.sqrt(z)) - mp.sqrt(z)) / mp.sqrt(z) ** 3, 0, q * q, 3)
cC = fit(lambda z: mp.mpf(1) / 24 if z == 0 else (mp.cos(mp.sqrt(z)) - 1 + z / 2) / z ** 2, 0, q * q, 3)
show("SIN", cS)
show("COS", cC)
x = np.linspace(-q, q, 200001).astype(np.float32)
z = np.float32(x * x)
s = np.float32(x + np.float32(np.float32(x * z) * horner(cS, z)))
co = np.float32(np.float32(np.float32(1) - np.float32(0.5) * z) + np.float32(np.float32(z * z) * horner(cC, z)))
print("   sin max abs err", np.max(np.abs(s - np.sin(x.astype(np.float64)))))
print("   cos max abs err", np.max(np.abs(co - np.cos(x.astype(np.float64)))))
for nm, v in (("LOG2E", 1.4426950408889634), ("NEG2LN2", -2 * np.log(2)), ("HALFPI", np.pi / 2)):
    print(nm, float(np.float32(v)).hex())
