"""CPU oracle on C4's matrix at a size the host can hold (n = 2^22, the
same generator as the 2^26 run): FCI-like, 16 roots of 21, Davidson-Liu, max_dav = 10, tol 1e-8,
lowest-diagonal unit start + 10 % noise.  Writes the fixture the GPU runs compare with.
usage: python tools/c4_oracle.py [bits=22] [out=tests/golden/c4_oracle_n22.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diaglib_b200 import problems as P  # noqa: E402
from oracle import oracle as O  # noqa: E402

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 22
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", f"c4_oracle_n{bits}.json")
n, n_targ, n_max, tol, max_dav, noise = 1 << bits, 16, 21, 1e-8, 10, 0.1
O.set_threads(os.cpu_count() or 1)
t0 = time.time()
csr = P.fci_like(n)
O.set_csr(*csr)
g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (noise / np.sqrt(n / 12.0)))
t_gen = time.time() - t0
t0 = time.time()
r = O.davidson(g, n_targ, 100, tol, max_dav)
wall = time.time() - t0
res = {"problem": "C4 fci_like", "bits": bits, "n": n, "nnz": int(len(csr[1])), "n_targ": n_targ, "n_max": n_max, "tol": tol,
       "max_dav": max_dav, "noise": noise, "ok": bool(r["ok"]), "iterations": int(len(r["it"])), "wall_s": wall,
       "gen_s": t_gen, "threads": O.get_threads(), "eig": [float(x) for x in r["eig"]],
       "rms": [float(x) for x in r["rms"][-1]], "max": [float(x) for x in r["max"][-1]],
       "n_act": [int(x) for x in r["n_act"]], "timers_s": {k: float(v) for k, v in r["timers"].items()},
       "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
json.dump(res, open(out, "w"))
print(json.dumps({k: res[k] for k in ("n", "iterations", "ok", "wall_s", "gen_s", "threads")}))
