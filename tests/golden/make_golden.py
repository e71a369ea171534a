"""Regenerates the committed fixtures under tests/golden/.

  toy_dense_eigs.json : lowest eigenvalues of the reference's own test matrix
      a(i,i)=i+1, a(i,j)=1/(i+j), n=1000 (main.f90:311-317) from dense LAPACK dsyev — the
      only known-answer data the reference's test (main.f90:321-342) defines.  Computed with
      two independent LAPACK builds (scipy-OpenBLAS dsyev through the oracle, numpy eigvalsh).
  c1_oracle_history.json : iteration history of the oracle on config C1 (n=1000, 10 roots of
      15, tol 1e-8, max_dav 20; main.f90:14-18) for LOBPCG and Davidson-Liu, guess from
      diaglib_b200.problems.guess(seed=1).  Pins the oracle against accidental edits; the
      reference itself records no iteration counts.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diaglib_b200 import problems as P  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    n, n_want = 1000, 10
    n_eig = P.n_eig_rule(n_want)
    a = P.toy_dense(n)
    w1, _ = O.dsyev(a)
    w2 = np.linalg.eigvalsh(a)
    assert np.abs(w1 - w2).max() < 1e-10
    json.dump({"n": n, "source": "dense LAPACK dsyev on main.f90:311-317 matrix",
               "max_abs_diff_two_lapack_builds": float(np.abs(w1 - w2).max()),
               "eig": [float(v) for v in w1[:20]]}, open(os.path.join(HERE, "toy_dense_eigs.json"), "w"), indent=1)
    O.set_dense(a)
    out = {}
    ev = P.guess(n, n_eig)
    r = O.lobpcg(ev, n_want, 100, 1e-8, matvec="oracle_dense_matvec")
    out["lobpcg"] = {"ok": r["ok"], "iterations": int(len(r["it"])), "eig": [float(v) for v in r["eig"]],
                     "n_act": [int(v) for v in r["n_act"]]}
    ev = P.guess(n, n_eig)
    r = O.davidson(ev, n_want, 100, 1e-8, 20, matvec="oracle_dense_matvec")
    out["davidson"] = {"ok": r["ok"], "iterations": int(len(r["it"])), "eig": [float(v) for v in r["eig"]],
                       "n_act": [int(v) for v in r["n_act"]]}
    json.dump(out, open(os.path.join(HERE, "c1_oracle_history.json"), "w"), indent=1)
    print("lobpcg its", out["lobpcg"]["iterations"], "davidson its", out["davidson"]["iterations"])


if __name__ == "__main__":
    main()
