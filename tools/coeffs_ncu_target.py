"""ncu target: get_coeffs on the eigenvectors of a graded LOBPCG-like 111 x 111 reduced matrix
(n_max = 37, n_act = 37 and 20), the shape of the C3 solve."""
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
w, z, sw = K.sym_eig(a)
z = np.asfortranarray(z)
for n_act in (37, 37, 20):
    u_p, st = K.get_coeffs(z, k, 37, n_act)
    print("n_act", n_act, st)
