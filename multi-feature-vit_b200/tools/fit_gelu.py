"""Derives the coefficients of normal_cdf_fast() in csrc/common.cuh.

Phi(x) = 1 / (1 + exp(-x * P(x^2))), P a degree-4 polynomial in x^2 fitted (iteratively re-weighted least squares ->
near-minimax) so that max |x * (sigma(x P) - Phi(x))| is minimal on |x| <= 7, then re-checked in float32 arithmetic on
|x| <= 12 against scipy's ndtr.  Prints the coefficients pre-multiplied by -log2(e) (the kernel uses ex2).
"""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import ndtr

x = np.linspace(-7, 7, 28001)
Phi = ndtr(x)


def model(c, x):
    x2 = x * x
    p = np.zeros_like(x)
    for ck in c[::-1]:
        p = p * x2 + ck
    return p * x


def resid(c):
    s = 1 / (1 + np.exp(-model(c, x)))
    return np.concatenate([x * (s - Phi), 0.3 * (s - Phi)])


c = np.array([1.5958, 0.0713, 0, 0, 0])
w = np.ones(2 * len(x))
for _ in range(200):
    c = least_squares(lambda c: resid(c) * w, c, method="lm").x
    e = np.abs(resid(c))
    w = w * (1 + 3 * e / e.max())
    w /= w.mean()
LOG2E = 1.4426950408889634
q = (-c * LOG2E).astype(np.float32)
xs = np.linspace(-12, 12, 2400001).astype(np.float32)
x2 = xs * xs
p = np.full_like(xs, q[4])
for k in (3, 2, 1, 0):
    p = p * x2 + q[k]
with np.errstate(over="ignore"):
    t = np.exp2((p * xs).astype(np.float32))
cdf = np.float32(1) / (np.float32(1) + t)
xd = xs.astype(np.float64)
print("max |gelu err|  %.3e" % np.abs(xs * cdf - xd * ndtr(xd)).max())
pdf = np.exp2((np.float32(-0.5 * LOG2E) * x2).astype(np.float32)) * np.float32(0.3989422804014327)
print("max |gelu' err| %.3e" % np.abs(cdf + xs * pdf - (ndtr(xd) + xd * np.exp(-0.5 * xd * xd) / np.sqrt(2 * np.pi))).max())
print("q (Horner, highest degree last):", [float(v) for v in q])
