"""Regenerates the committed fixtures under tests/golden/.

  toy_dense_eigs.json : lowest eigenvalues of the reference's own test matrix
      a(i,i)=i+1, a(i,j)=1/(i+j), n=1000 (main.f90:311-317) from dense LAPACK dsyev — the
      only known-answer data the reference's test (main.f90:321-342) defines.  Computed with
      two independent LAPACK builds (scipy-OpenBLAS dsyev through the oracle, numpy eigvalsh).
  c1_oracle_history.json : iteration history of the oracle on config C1 (n=1000, 10 roots of
      15, tol 1e-8, max_dav 20; main.f90:14-18) for LOBPCG and Davidson-Liu, guess from
      diaglib_b200.problems.guess(seed=1).  Pins the oracle against accidental edits; the
      reference itself records no iteration counts.

  c3_oracle_nx{32,128,256}.json : the oracle's complete LOBPCG solve of the benchmark workload
      (C3: lap3d nx^3, 32 roots of 37, tol 1e-8, lowest-diagonal start + 10 % noise) as written by
      `python bench.py --impl reference --nx NX` (gpurun_out/oracle_c3_nxNX.json, copied here).
      nx = 32 is re-checked against the oracle on CPU (tests/test_oracle.py), nx = 128 is the
      GPU parity test at 2 M rows (tests/test_gpu_drivers.py), nx = 256 is the headline problem
      (16.7 M rows, about 8 minutes of host time on the GPU box) and backs bench.py's parity block
      when no fresh run is present.  The thread count of the generating run is recorded inside.

  c3_oracle_acc_nx{64,128,256}.json : the same solves by the oracle in its DIAGNOSTIC mode (reduced
      eigenproblems through dpotrf + dgesvj instead of dsyev): `python tools/oracle_spread.py NX T 0 OUT 1`.
      They quantify how much of an iteration-count difference is dsyev's accuracy (bench.parity_block).
  c4_oracle_n22.json : the oracle's Davidson-Liu solve of C4's matrix at n = 2^22 (tools/c4_oracle.py).
  c5_oracle_n18.json : the oracle's Davidson-Liu and LOBPCG solves of C5's problem at n = 2^18
      (128 roots of 133, lda = 1330; tools/c5_oracle.py).

Run from the repo root:  python tests/golden/make_golden.py   (the first two files)
                         python bench.py --impl reference --nx 32 && cp gpurun_out/oracle_c3_nx32.json tests/golden/c3_oracle_nx32.json
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
