"""Synthetic problem generators for the diaglib hot path.

Every generator is a pure function of (seed, global row[, global column]) built on a
stateless splitmix64 hash, so that the CPU oracle, the CUDA path and every row shard of
a multi-GPU run see bit-identical data.  Rows are generated for a contiguous global row
range [r0, r1) and returned as CSR with GLOBAL column indices (int32) and int64 rowptr.

Problems (names follow BASELINE.json `configs`):
  toy_dense   C1  the reference's own test matrix, main.f90:311-317
  toy_sparse  C2/C5  the same entries kept only at |i-j| in {1,2,4,...}
  lap3d       C3  3-D 7-point Laplacian (Dirichlet) + permuted-progression diagonal
  fci_like    C4  strong diagonal + ~100 banded off-diagonals with hashed symmetric values
  guess       uniform(-0.5,0.5) start vectors, mirrors guess_evec(4), main.f90:1362-1367
"""
from __future__ import annotations

import numpy as np

_U64 = np.uint64
_GOLD = _U64(0x9E3779B97F4A7C15)
_M1 = _U64(0xBF58476D1CE4E5B9)
_M2 = _U64(0x94D049BB133111EB)


def splitmix64(x):
    """One splitmix64 output step applied element-wise to uint64 input."""
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=_U64) + _GOLD
        z = (z ^ (z >> _U64(30))) * _M1
        z = (z ^ (z >> _U64(27))) * _M2
        return z ^ (z >> _U64(31))


def hash_u01(seed: int, idx):
    """U[0,1) double from (seed, idx): top 53 bits of splitmix64(splitmix64(seed) ^ idx)."""
    s = splitmix64(np.array([seed], dtype=_U64))[0]
    z = splitmix64(np.asarray(idx, dtype=_U64) ^ s)
    return (z >> _U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def guess(n_glob: int, n_max: int, r0: int = 0, r1: int | None = None, seed: int = 1):
    """Start vectors evec(i,j) = U(-0.5,0.5), element index i + n_glob*j (column-major,
    Fortran order), rows [r0,r1).  Mirrors guess_evec(4) (main.f90:1362-1367)."""
    r1 = n_glob if r1 is None else r1
    out = np.empty((r1 - r0, n_max), dtype=np.float64, order="F")
    rows = np.arange(r0, r1, dtype=_U64)
    for j in range(n_max):
        out[:, j] = hash_u01(seed, rows + _U64(n_glob) * _U64(j)) - 0.5
    return out


def bijection(idx, bits: int, seed: int = 1):
    """Seeded bijection on `bits`-bit integers: rounds of invertible xorshift / odd-multiply
    steps modulo 2**bits."""
    mask = _U64((1 << bits) - 1)
    x = np.asarray(idx, dtype=_U64) & mask
    ks = splitmix64(np.arange(4, dtype=_U64) + _U64(seed) * _U64(1000003))
    sh = _U64(max(1, bits // 2))
    with np.errstate(over="ignore"):
        for r in range(4):
            x = (x * (ks[r] | _U64(1))) & mask
            x = x ^ (x >> sh)
            x = (x + (ks[r] >> _U64(17))) & mask
    return x


def _compress(cols, vals, valid):
    """Row-wise compress candidate (n,k) arrays into CSR."""
    counts = valid.sum(axis=1).astype(np.int64)
    rowptr = np.zeros(valid.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, cols[valid].astype(np.int32), vals[valid].astype(np.float64)


def toy_dense(n: int = 1000):
    """a(i,i)=i+1, a(i,j)=1/(i+j), 1-based (main.f90:311-317).  Column-major n x n."""
    i = np.arange(1, n + 1, dtype=np.float64)
    a = 1.0 / (i[:, None] + i[None, :])
    a[np.arange(n), np.arange(n)] = i + 1.0
    return np.asfortranarray(a)


def toy_sparse(n: int, r0: int = 0, r1: int | None = None):
    """C2/C5: the toy matrix restricted to |i-j| in {1,2,4,...,<n}.  Returns
    (rowptr, col, val, diag) for rows [r0,r1)."""
    r1 = n if r1 is None else r1
    rows = np.arange(r0, r1, dtype=np.int64)
    offs = []
    k = 1
    while k < n:
        offs.append(k)
        k *= 2
    offsets = np.array([-o for o in reversed(offs)] + [0] + offs, dtype=np.int64)
    cols = rows[:, None] + offsets[None, :]
    valid = (cols >= 0) & (cols < n)
    with np.errstate(divide="ignore"):
        vals = 1.0 / ((rows[:, None] + 1) + (cols + 1)).astype(np.float64)
    diag = (rows + 2).astype(np.float64)
    vals[:, len(offs)] = diag
    rowptr, col, val = _compress(cols, vals, valid)
    return rowptr, col, val, diag


def lap3d_diag(rows, bits: int, delta: float = 1.0, seed: int = 1):
    """d_i = 6 + delta*(1 + pi(i)), pi a seeded bijection on `bits`-bit integers."""
    return 6.0 + delta * (1.0 + bijection(rows, bits, seed).astype(np.float64))


def lap3d(nx: int, ny: int, nz: int, r0: int = 0, r1: int | None = None, delta: float = 1.0, seed: int = 1):
    """C3: 7-point Laplacian on an nx*ny*nz grid, Dirichlet, off-diagonals -1, diagonal
    lap3d_diag.  i = x + nx*(y + ny*z).  nx*ny*nz must be a power of two."""
    n = nx * ny * nz
    bits = n.bit_length() - 1
    assert 1 << bits == n, "lap3d needs a power-of-two number of sites"
    r1 = n if r1 is None else r1
    rows = np.arange(r0, r1, dtype=np.int64)
    x = rows % nx
    y = (rows // nx) % ny
    z = rows // (nx * ny)
    offsets = np.array([-nx * ny, -nx, -1, 0, 1, nx, nx * ny], dtype=np.int64)
    cols = rows[:, None] + offsets[None, :]
    valid = np.stack([z > 0, y > 0, x > 0, np.ones_like(x, dtype=bool), x < nx - 1, y < ny - 1, z < nz - 1], axis=1)
    vals = np.full(cols.shape, -1.0)
    diag = lap3d_diag(rows, bits, delta, seed)
    vals[:, 3] = diag
    rowptr, col, val = _compress(cols, vals, valid)
    return rowptr, col, val, diag


def _morton3(ix, iy, iz):
    """interleave the bits of three non-negative integer arrays (z-order curve key)"""
    key = np.zeros(ix.shape, dtype=np.int64)
    for b in range(21):
        key |= ((ix >> b) & 1) << (3 * b) | ((iy >> b) & 1) << (3 * b + 1) | ((iz >> b) & 1) << (3 * b + 2)
    return key


def tile_order_3d(nx: int, ny: int, nz: int, tile=(32, 4, 2), z0: int = 0, z1: int | None = None, curve: str = "morton"):
    """Processing order for the rows of a 3-D grid stencil (row i = x + nx*(y + ny*z)) owned by a
    rank holding the planes [z0, z1): the grid is cut into tiles of tile = (tx, ty, tz) sites (one
    CTA of the SpMM = 256 consecutive entries = one 32x4x2 tile), tiles are visited along a
    z-order curve ("morton") or plane by plane ("sweep").  Returns LOCAL row indices (int32
    permutation of [0, nx*ny*(z1-z0))).  Pure locality hint for diaglib_b200.set_csr_row_order."""
    z1 = nz if z1 is None else z1
    tx, ty, tz = tile
    lz = z1 - z0
    ntx, nty, ntz = -(-nx // tx), -(-ny // ty), -(-lz // tz)
    bx, by, bz = np.meshgrid(np.arange(ntx), np.arange(nty), np.arange(ntz), indexing="ij")
    bx, by, bz = bx.ravel(), by.ravel(), bz.ravel()
    tkey = _morton3(bx, by, bz) if curve == "morton" else (bz * nty + by) * ntx + bx
    t = np.argsort(tkey, kind="stable")
    bx, by, bz = bx[t], by[t], bz[t]
    # rows of every tile, x fastest, then y, then z: shape (tiles, tz, ty, tx)
    x = (bx * tx)[:, None, None, None] + np.arange(tx)[None, None, None, :]
    y = (by * ty)[:, None, None, None] + np.arange(ty)[None, None, :, None]
    z = (bz * tz)[:, None, None, None] + np.arange(tz)[None, :, None, None]
    rows = (x + nx * (y + ny * z)).reshape(-1)
    if nx % tx or ny % ty or lz % tz:
        ok = ((x < nx) & (y < ny) & (z < lz)).reshape(-1)
        rows = rows[ok]
    return rows.astype(np.int32)


def fci_strides(n_strides: int = 50, bandwidth: int = 1 << 20, seed: int = 1):
    """Fixed seeded set of distinct positive strides <= bandwidth (sorted ascending)."""
    out: list[int] = []
    seen = set()
    c = 0
    while len(out) < n_strides:
        s = int(splitmix64(np.array([c + 7919 * seed], dtype=_U64))[0] % _U64(bandwidth)) + 1
        c += 1
        if s not in seen:
            seen.add(s)
            out.append(s)
    return np.array(sorted(out), dtype=np.int64)


def fci_like(n: int, r0: int = 0, r1: int | None = None, n_strides: int = 50, bandwidth: int = 1 << 20,
             big_delta: float = 0.1, seed: int = 1):
    """C4: d_i = 1 + Delta*pi(i); off-diagonals at i +- s_k with values U(-0.01,0.01) hashed
    from (min(i,j), max(i,j)) so the matrix is exactly symmetric.  n must be a power of two."""
    bits = n.bit_length() - 1
    assert 1 << bits == n
    r1 = n if r1 is None else r1
    strides = fci_strides(n_strides, min(bandwidth, max(1, n // 2)), seed)
    rows = np.arange(r0, r1, dtype=np.int64)
    offsets = np.concatenate([-strides[::-1], [0], strides])
    cols = rows[:, None] + offsets[None, :]
    valid = (cols >= 0) & (cols < n)
    lo = np.minimum(rows[:, None], cols).astype(_U64)
    hi = np.maximum(rows[:, None], cols).astype(_U64)
    vals = (hash_u01(seed + 17, lo * _U64(n) + hi) - 0.5) * 0.02
    diag = 1.0 + big_delta * bijection(rows, bits, seed).astype(np.float64)
    vals[:, len(strides)] = diag
    rowptr, col, val = _compress(cols, vals, valid)
    return rowptr, col, val, diag


def csr_to_dense(n_cols: int, rowptr, col, val):
    n = len(rowptr) - 1
    a = np.zeros((n, n_cols))
    for i in range(n):
        a[i, col[rowptr[i]:rowptr[i + 1]]] = val[rowptr[i]:rowptr[i + 1]]
    return a


def n_eig_rule(n_want: int) -> int:
    """Search-space rule of the reference driver: n_eig = min(2*n_want, n_want+5) (main.f90:354)."""
    return min(2 * n_want, n_want + 5)


def guess_lowest_diag(diag_glob, n_max: int, r0: int = 0, r1: int | None = None):
    """Unit start vectors on the n_max smallest diagonal entries (ties: lowest index first),
    mirrors guess_evec(1) (main.f90:1337-1347).  diag_glob is the GLOBAL diagonal."""
    diag_glob = np.asarray(diag_glob)
    n = diag_glob.shape[0]
    r1 = n if r1 is None else r1
    order = np.argsort(diag_glob, kind="stable")[:n_max]
    out = np.zeros((r1 - r0, n_max), dtype=np.float64, order="F")
    for j, ipos in enumerate(order):
        if r0 <= ipos < r1:
            out[ipos - r0, j] = 1.0
    return out


def metric_like(csr, eps: float = 0.02, seed: int = 7, r0: int = 0):
    """SPD metric B for the generalized problem A x = lambda B x (gen_eig branch,
    diaglib.f90:299-302): same sparsity pattern as the matrix, unit-ish diagonal
    b_ii = 1 + 0.3 u(i) and symmetric off-diagonals eps * u(min(i,j), max(i,j)) / row-length scale,
    diagonally dominant by construction.  csr holds rows [r0, r0 + n_loc) with GLOBAL column
    indices.  Returns (rowptr, col, val)."""
    rowptr, col, _, _ = csr
    n = len(rowptr) - 1
    rows = r0 + np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    c = col.astype(np.int64)
    lo, hi = np.minimum(rows, c), np.maximum(rows, c)
    h = splitmix64((lo * np.int64(2654435761) + hi + np.int64(seed) * np.int64(1000003)).astype(np.uint64))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    maxlen = max(1, int(np.diff(rowptr).max()))
    val = eps * (u - 0.5) * (2.0 / maxlen)
    dg = rows == c
    val[dg] = 1.0 + 0.3 * u[dg]
    return rowptr.copy(), col.copy(), val


def caslr_like(n: int, r0: int = 0, r1: int | None = None, seed: int = 11):
    """Linear-response test problem in the spirit of main.f90:528-600 on the toy_sparse pattern
    (|i-j| in {1,2,4,...}): apb = A+B (diag 5+i, off-diag 1/(i+j)), amb = A-B (diag 2+i, off-diag
    0.2/(i+j)), sigma SPD (diag 1+u, small symmetric off-diagonals), delta antisymmetric (small).
    Returns dict(apb, amb, spd, smd = (rowptr, col_global, val) for rows [r0, r1), aa_diag,
    sigma_diag)."""
    rowptr, col, val, _ = toy_sparse(n, r0, r1)
    r1 = n if r1 is None else r1
    rows = r0 + np.repeat(np.arange(r1 - r0, dtype=np.int64), np.diff(rowptr))
    c = col.astype(np.int64)
    dg = rows == c
    i1, j1 = rows + 1, c + 1
    apb = np.where(dg, 5.0 + i1, 1.0 / (i1 + j1))
    amb = np.where(dg, 2.0 + i1, 0.2 / (i1 + j1))
    lo, hi = np.minimum(rows, c), np.maximum(rows, c)
    h = splitmix64((lo * np.int64(2654435761) + hi + np.int64(seed) * np.int64(1000003)).astype(np.uint64))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    h2 = splitmix64(h)
    u2 = (h2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    maxlen = max(1, int(np.diff(rowptr).max()))
    sigma = np.where(dg, 1.0 + u, 0.2 * (u - 0.5) / maxlen)
    delta = np.where(dg, 0.0, np.sign(c - rows) * 0.2 * u2 / maxlen)       # delta(j,i) = -delta(i,j)
    mk = lambda v_: (rowptr.copy(), col.copy(), np.ascontiguousarray(v_, dtype=np.float64))  # noqa: E731
    return dict(apb=mk(apb), amb=mk(amb), spd=mk(sigma + delta), smd=mk(sigma - delta),
                aa_diag=0.5 * (apb[dg] + amb[dg]), sigma_diag=sigma[dg])
