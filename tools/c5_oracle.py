"""CPU oracle on C5's problem at a size the host solves in minutes: toy_sparse n = 2^bits, 128 roots
of 133, Davidson-Liu with max_dav = 10 (lda = 1330) and LOBPCG (len_a = 399), tol 1e-8, random
start (guess_evec(4), as the reference's own test).  Writes the fixture the GPU test compares with.
usage: python tools/c5_oracle.py [bits=18] [out=tests/golden/c5_oracle_n18.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diaglib_b200 import problems as P  # noqa: E402
from oracle import oracle as O  # noqa: E402

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 18
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", f"c5_oracle_n{bits}.json")
n, n_targ, n_max, tol, max_dav, noise = 1 << bits, 128, 133, 1e-8, 10, 0.1
O.set_threads(os.cpu_count() or 1)
csr = P.toy_sparse(n)
O.set_csr(*csr)
res = {"problem": "C5 toy_sparse", "bits": bits, "n": n, "nnz": int(len(csr[1])), "n_targ": n_targ, "n_max": n_max, "tol": tol,
       "max_dav": max_dav, "noise": noise, "guess": "lowest-diag unit + 10% noise", "threads": O.get_threads()}
for drv in ("davidson", "lobpcg"):
    g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (noise / np.sqrt(n / 12.0)))
    t0 = time.time()
    r = O.davidson(g, n_targ, 100, tol, max_dav) if drv == "davidson" else O.lobpcg(g, n_targ, 100, tol)
    res[drv] = {"ok": bool(r["ok"]), "iterations": int(len(r["it"])), "wall_s": time.time() - t0,
                "eig": [float(x) for x in r["eig"]], "n_act": [int(x) for x in r["n_act"]],
                "rms_max": float(r["rms"][-1][:n_targ].max()), "max_max": float(r["max"][-1][:n_targ].max())}
    print(drv, res[drv]["ok"], res[drv]["iterations"], round(res[drv]["wall_s"], 1), flush=True)
res["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
json.dump(res, open(out, "w"))
