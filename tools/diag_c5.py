import sys
import numpy as np
sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K, problems as P
from oracle import oracle as O
n = 1 << 13
D.init(0); O.set_threads(16)
csr = P.toy_sparse(n); O.set_csr(*csr); D.set_csr(*csr)
n_targ, n_max = 128, 133
ev_o = P.guess(n, n_max); ev_g = ev_o.copy(order="F")
ro = O.lobpcg(ev_o, n_targ, 6, 1e-8)
eig = np.zeros(n_max)
D.lobpcg_driver(False, False, n, n_targ, n_max, 6, 1e-8, 0.0, None, None, None, eig, ev_g)
h = D.last_history(n_max)
for it in range(len(h["it"])):
    d = np.abs(h["eig"][it] - ro["hist_eig"][it]) / np.abs(ro["hist_eig"][it])
    print("it", it + 1, "n_act", h["n_act"][it], ro["n_act"][it], "max rel eig diff", d.max(), "argmax", d.argmax(),
          "gpu", h["eig"][it][[0, 1, 127, 132]], "ora", ro["hist_eig"][it][[0, 1, 127, 132]])
print(D.last_stats(), ro["stats"])
# get_coeffs at the big shape vs oracle
rng = np.random.default_rng(0)
len_u = 399
s = rng.standard_normal((len_u, len_u)); a = np.diag(np.arange(1.0, len_u + 1)) + 0.02 * (s + s.T)
_, z, _ = K.sym_eig(a)
z = np.asfortranarray(z)
for n_act in (133, 100):
    lu = 133 + 2 * n_act
    zz = np.asfortranarray(z[:lu, :lu])
    up, st = K.get_coeffs(zz, lu, 133, n_act)
    ux_ref, up_ref = O.get_coeffs(zz, lu, 133, n_act)
    print("get_coeffs", n_act, st, "diff", np.abs(up - up_ref).max(), "orth", np.abs(up.T @ up - np.eye(n_act)).max(), np.abs(zz[:, :133].T @ up).max())
