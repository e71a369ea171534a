"""Repeats one LOBPCG solve of the scaled-down benchmark workload (lap3d nx^3, 32 roots) and checks that
every repetition reproduces the first one bit for bit (iteration count, eigenvalue history).
usage: python tools/determinism_check.py [nx=128] [reps=5]"""
import hashlib
import sys

import numpy as np

sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import problems as P

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n, n_targ = nx ** 3, 32
n_max = P.n_eig_rule(n_targ)
D.init(0)
csr = P.lap3d(nx, nx, nx, delta=1.0)
D.set_csr(*csr)
import os
if not os.environ.get("DIAGLIB_B200_DET_NATURAL"):
    D.set_csr_row_order(P.tile_order_3d(nx, nx, nx, tile=(64, 2, 2), curve="morton"))
g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (0.1 / np.sqrt(n / 12.0)))
first = None
bad = 0
for r in range(reps):
    ev = g.copy(order="F")
    eig = np.zeros(n_max)
    ok = D.lobpcg_driver(False, False, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, ev)
    h = D.last_history(n_max)
    sig = hashlib.sha1(np.asarray(h["eig"]).tobytes() + ev.tobytes()).hexdigest()[:12]
    st = D.last_stats()
    print(f"rep {r}: ok={ok} its={len(h['it'])} passes={st['ortho_cd_passes']} sweeps={st['ortho_vs_x_sweeps']} sig={sig}", flush=True)
    if first is None:
        first = sig
    bad += sig != first
print("DETERMINISTIC" if bad == 0 else f"NOT DETERMINISTIC ({bad} of {reps} differ)")
