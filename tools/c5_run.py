"""C5 at full size on one GPU: toy_sparse n = 2^22, 128 roots, n_max = 133,
max_dav = 10 (lda = 1330), Davidson; then LOBPCG (len_a = 399).  `c5_run.py 22 c4`: the FCI-like
matrix of C4 (101 entries per row, strides up to 2^20), 16 roots, n_max = 21, at n = 2^22 on one GPU.  Size-independent checks:
returned residuals recomputed on the host, orthonormality, agreement of the two drivers."""
import json
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K, problems as P

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 22
problem = sys.argv[2] if len(sys.argv) > 2 else "c5"
n = 1 << bits
n_targ, n_max, tol = (128, 133, 1e-8) if problem == "c5" else (16, 21, 1e-8)   # C5 / C4 of SURVEY 8d
D.init(0)
t0 = time.time()
csr = P.toy_sparse(n) if problem == "c5" else P.fci_like(n)
D.set_csr(*csr)
a = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(n, n))
g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (0.1 / np.sqrt(n / 12.0)))
out = dict(problem=problem, nnz=int(len(csr[1])), n=n, n_targ=n_targ, n_max=n_max, setup_s=round(time.time() - t0, 1))
eigs = {}
for drv in ("davidson", "lobpcg"):
    dev = K.DeviceArray.from_numpy(g)
    eig = np.zeros(n_max)
    K.timer_start()
    if drv == "davidson":
        ok = D.davidson_driver(False, n, n_targ, n_max, 100, tol, 10, 0.0, None, None, eig, dev)
    else:
        ok = D.lobpcg_driver(False, False, n, n_targ, n_max, 100, tol, 0.0, None, None, None, eig, dev)
    ms = K.timer_stop_ms()
    x = dev.numpy()[:, :n_targ]
    dev.free()
    res = a @ x - x * eig[:n_targ]
    its = len(D.last_history(n_max)["it"])
    eigs[drv] = eig[:n_targ].copy()
    out[drv] = dict(ok=ok, its=its, ms=round(ms, 1), rms_res_max=float((np.linalg.norm(res, axis=0) / np.sqrt(n)).max()),
                    max_res=float(np.abs(res).max()), ortho_err=float(np.abs(x.T @ x - np.eye(n_targ)).max()),
                    phases={k: round(float(v), 3) for k, v in D.last_timers().items() if v}, stats=D.last_stats())
    D.lib().diaglib_b200_release_workspace()
out["drivers_agree_rel"] = float(np.abs(eigs["davidson"] - eigs["lobpcg"]).max() / np.abs(eigs["lobpcg"]).max())
print(json.dumps(out))
