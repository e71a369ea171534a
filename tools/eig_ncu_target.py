"""ncu target: a few reduced eigensolves of the graded LOBPCG-like 111 x 111 matrix (and one
positive definite 399 x 399) with the one-sided solver."""
import sys

import numpy as np

sys.path.insert(0, ".")
from diaglib_b200 import kernels as K

rng = np.random.default_rng(12)
k = 111
d = np.concatenate([np.arange(7.0, 44.0), 50 + 1e3 * rng.random(37), 1e6 + 1e7 * rng.random(37)])
cpl = rng.standard_normal((k, k))
cpl = 1e-3 * (cpl + cpl.T) * np.sqrt(np.outer(d, d)) / d.max() ** 0.5
a = np.diag(d) + cpl
np.fill_diagonal(a, d)
for _ in range(3):
    w, z, sw = K.sym_eig(a)
print("k=111 sweeps", sw, "path", K.sym_eig.last_path)
k = 399
s = np.random.default_rng(k).standard_normal((k, k))
b = np.diag(np.arange(1.0, k + 1)) + 0.02 * (s + s.T) + (s @ s.T) / (4 * k)
w, z, sw = K.sym_eig(b)
print("k=399 sweeps", sw)
