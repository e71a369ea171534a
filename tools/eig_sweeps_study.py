"""Why the reduced eigensolver needs ~8 sweeps inside a solve (3-4 on random test matrices).
Input: the reduced matrices of a CPU oracle solve, dumped with ORACLE_DUMP_ARED=<file> (one record per
LOBPCG iteration: len_u, then the len_u x len_u matrix column by column).  A numpy emulation of the
kernel's algorithm (sorted diagonal, Cholesky factor, one-sided Jacobi on its columns with the
rounding-floor rotation criterion |x.y| > sqrt(k) eps sum|x_i y_i|) prints, per sweep, the largest
rotated ratio and the number of rotations, for the plain factor and for three preconditioners of the
Drmac-Veselic kind (complete pivoting, one and two Cholesky-LR steps).

Finding (lap3d 64^3, 32 roots, the benchmark's start): 8-9 sweeps in every variant - a linear phase of
five sweeps (1, 0.4, 0.13, 0.07, 0.01) before the quadratic one; the ratio that has to fall below 1e-10
is relative to each pair's own rounding floor, which is what buys eigenvectors accurate relative to
every Ritz value, and no preconditioner shortens it.  So the sweep count is the price of the accuracy
criterion, not of the ordering or the pivoting.

usage:  ORACLE_DUMP_ARED=/tmp/ared.bin python -c "..."   (any oracle.lobpcg call), then
        python tools/eig_sweeps_study.py /tmp/ared.bin [matrix indices ...]"""
import sys

import numpy as np

EPS = np.finfo(float).eps


def load(path):
    raw, mats, i = np.fromfile(path), [], 0
    while i < len(raw):
        k = int(raw[i])
        mats.append(raw[i + 1:i + 1 + k * k].reshape(k, k).T.copy())
        i += 1 + k * k
    return mats


def sweeps_on(L, stop=1e-10, maxsw=30):
    L, k = L.copy(), L.shape[0]
    tol, hist = EPS * np.sqrt(k), []
    for _ in range(maxsw):
        mx, nrot = 0.0, 0
        for p in range(k - 1):
            for q in range(p + 1, k):
                x, y = L[:, p], L[:, q]
                apq, sab = x @ y, np.abs(x * y).sum()
                if abs(apq) > tol * sab:
                    mx, nrot = max(mx, abs(apq) / max(sab, 1e-300)), nrot + 1
                    zeta = (y @ y - x @ x) / (2 * apq)
                    t = np.sign(zeta) / (abs(zeta) + np.sqrt(1 + zeta * zeta)) if zeta != 0 else 1.0
                    c = 1 / np.sqrt(1 + t * t)
                    s = t * c
                    L[:, p], L[:, q] = c * x - s * y, s * x + c * y
        hist.append((mx, nrot))
        if mx <= stop:
            break
    return hist


def chol_pivoted(a):
    k, a, L = a.shape[0], a.copy(), np.zeros_like(a)
    for j in range(k):
        p = j + np.argmax(np.diag(a)[j:])
        if p != j:
            a[[j, p], :], a[:, [j, p]], L[[j, p], :] = a[[p, j], :], a[:, [p, j]], L[[p, j], :]
        L[j, j] = np.sqrt(a[j, j])
        L[j + 1:, j] = a[j + 1:, j] / L[j, j]
        a[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    return L


if __name__ == "__main__":
    mats = load(sys.argv[1])
    which = [int(v) for v in sys.argv[2:]] or [1, len(mats) // 2, len(mats) - 3]
    fmt = lambda h: f"{len(h)} sweeps: " + " ".join(f"{m:.1e}/{n}" for m, n in h)  # noqa: E731
    for idx in which:
        a = np.tril(mats[idx]) + np.tril(mats[idx], -1).T
        pm = np.argsort(-np.diag(a), kind="stable")
        L0 = np.linalg.cholesky(a[np.ix_(pm, pm)])
        print(f"matrix {idx} (k = {a.shape[0]})")
        print("  sorted diagonal (the kernel):", fmt(sweeps_on(L0)))
        print("  complete pivoting:           ", fmt(sweeps_on(chol_pivoted(a))))
        L1 = np.linalg.cholesky(L0.T @ L0)
        print("  one Cholesky-LR step:        ", fmt(sweeps_on(L1)))
        print("  two Cholesky-LR steps:       ", fmt(sweeps_on(np.linalg.cholesky(L1.T @ L1))))
