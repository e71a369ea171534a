"""Generalized problem A x = lambda B x on the C3 matrix (one GPU): LOBPCG (gen_eig) and
gen_david_driver next to the standard drivers on the same matrix and guess; time to converge,
iterations and phase split.  B = problems.metric_like(A) (same pattern, SPD)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K, problems as P

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = nx ** 3
n_targ, n_max = 32, 37
D.init(0)
csr = P.lap3d(nx, nx, nx, delta=1.0)
D.set_csr(*csr)
D.set_csr_b(*P.metric_like(csr))
g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (0.1 / np.sqrt(n / 12.0)))
dg = K.DeviceArray.from_numpy(g)
dev = K.DeviceArray((n, n_max))
eig = np.zeros(n_max)
out = {}
for drv in ("lobpcg", "lobpcg_gen_eig", "davidson", "gen_david"):
    for rep in range(2):
        D.lib().diaglib_b200_d2d(dev.ptr, dg.ptr, g.nbytes)
        K.timer_start()
        if drv == "lobpcg":
            ok = D.lobpcg_driver(False, False, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, dev)
        elif drv == "lobpcg_gen_eig":
            ok = D.lobpcg_driver(False, True, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, dev)
        elif drv == "davidson":
            ok = D.davidson_driver(False, n, n_targ, n_max, 200, 1e-8, 10, 0.0, None, None, eig, dev)
        else:
            ok = D.gen_david_driver(False, n, n_targ, n_max, 200, 1e-8, 10, 0.0, None, None, None, eig, dev)
        ms = K.timer_stop_ms()
    its = len(D.last_history(n_max)["it"])
    out[drv] = dict(ok=ok, its=its, ms=round(ms, 1), its_per_s=round(its / ms * 1e3, 2), eig0=float(eig[0]),
                    phases={k: round(float(v), 4) for k, v in D.last_timers().items() if v})
print(json.dumps(dict(n=n, n_targ=n_targ, n_max=n_max, **out)))
