"""Synthetic operators and start vectors of the BASELINE configs.

Pure numpy; no CUDA needed.  Every generator can produce a row range [r0, r1) of the
global operator so that each rank of a row-partitioned run builds only its own shard.
Inputs to both the CUDA path and the CPU oracle come from here, so the two see
bit-identical matrices and start vectors.

The counter-based generator is splitmix64 (all uint64, wrap-around):
  u(seed, i) = (splitmix64(splitmix64(seed) + i) >> 11) * 2^-53   in [0, 1)
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed, start, count):
    """u(seed, i) for i in [start, start+count)."""
    base = splitmix64(np.uint64(seed))
    with np.errstate(over="ignore"):
        idx = base + np.arange(start, start + count, dtype=np.uint64)
    return (splitmix64(idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def start_vector(n, seed=7, dtype=np.float64, chunk=1 << 22):
    """x_i = 2u(seed,i)-1 (complex: re from seed, im from seed+1), normalised to unit 2-norm."""
    x = np.empty(n, dtype=dtype)
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        re = 2.0 * uniform01(seed, s, c) - 1.0
        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            x[s : s + c] = re + 1j * (2.0 * uniform01(seed + 1, s, c) - 1.0)
        else:
            x[s : s + c] = re
    nrm = np.sqrt(np.vdot(x, x).real)
    return x / nrm


def dense_symmetric(n, seed=1):
    """cfg 1: A = (R + R^T)/2, R_ij = 2u(seed, i*n+j) - 1."""
    r = (2.0 * uniform01(seed, 0, n * n) - 1.0).reshape(n, n)
    return 0.5 * (r + r.T)


def _compress(cols, vals, mask):
    counts = mask.sum(axis=1).astype(np.int64)
    rowptr = np.zeros(mask.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, cols[mask].astype(np.int32), vals[mask]


def laplacian2d_csr(N, r0=0, r1=None):
    """cfg 2: 2D 5-point Dirichlet Laplacian on an N x N grid, row-major index r = i*N + j,
    diagonal 4, neighbours -1.  Returns (rowptr int64, col int32, val float64) of rows [r0, r1)."""
    n = N * N
    r1 = n if r1 is None else r1
    r = np.arange(r0, r1, dtype=np.int64)
    i, j = r // N, r % N
    cols = np.stack([r - N, r - 1, r, r + 1, r + N], axis=1)
    mask = np.stack([i > 0, j > 0, np.ones_like(i, bool), j < N - 1, i < N - 1], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), cols.shape)
    return _compress(cols, vals, mask)


def laplacian2d_eigenvalues(N, count):
    """lowest `count` eigenvalues 4 - 2cos(i pi/(N+1)) - 2cos(j pi/(N+1)) (ascending)."""
    k = np.arange(1, min(N, 64) + 1)
    c = 2.0 - 2.0 * np.cos(k * np.pi / (N + 1))
    return np.sort((c[:, None] + c[None, :]).ravel())[:count]


def convdiff3d_csr(M, gamma=(0.3, 0.2, 0.1), r0=0, r1=None):
    """cfg 3: A = Tx(x)I(x)I + I(x)Ty(x)I + I(x)I(x)Tz, T_d = tridiag(-1-g_d, 2, -1+g_d),
    index r = (ix*M + iy)*M + iz; real, non-symmetric."""
    n = M * M * M
    r1 = n if r1 is None else r1
    r = np.arange(r0, r1, dtype=np.int64)
    ix, iy, iz = r // (M * M), (r // M) % M, r % M
    gx, gy, gz = gamma
    cols = np.stack([r - M * M, r - M, r - 1, r, r + 1, r + M, r + M * M], axis=1)
    mask = np.stack([ix > 0, iy > 0, iz > 0, np.ones_like(r, bool), iz < M - 1, iy < M - 1, ix < M - 1], axis=1)
    vals = np.broadcast_to(
        np.array([-1.0 - gx, -1.0 - gy, -1.0 - gz, 6.0, -1.0 + gz, -1.0 + gy, -1.0 + gx]), cols.shape
    )
    return _compress(cols, vals, mask)


def convdiff3d_eigenvalues(M, gamma=(0.3, 0.2, 0.1), count=5):
    """largest `count` eigenvalues 6 + 2 sum_d sqrt(1-g_d^2) cos(k_d pi/(M+1)) (descending)."""
    k = np.arange(1, min(M, 12) + 1)
    c = [2.0 * np.sqrt(1.0 - g * g) * np.cos(k * np.pi / (M + 1)) for g in gamma]
    lam = 6.0 + c[0][:, None, None] + c[1][None, :, None] + c[2][None, None, :]
    return np.sort(lam.ravel())[::-1][:count]


def heisenberg_csr(L, J=1.0, pbc=True, r0=0, r1=None, chunk=1 << 20):
    """cfg 4: spin-1/2 Heisenberg chain/ring, H = J sum_i [SzSz + (S+S- + S-S+)/2]_{i,i+1}; basis =
    bit strings 0..2^L-1 (bit i = spin i).  Diagonal J/4 (#aligned - #anti-aligned bonds), stored
    even when 0; off-diagonal J/2 to the state with bits i,i+1 flipped for each anti-aligned bond.
    Columns ascending within a row."""
    n = 1 << L
    r1 = n if r1 is None else r1
    nb = L if (pbc and L > 2) else L - 1
    rowptrs, colss, valss = [np.zeros(1, np.int64)], [], []
    base = 0
    for s0 in range(r0, r1, chunk):
        s = np.arange(s0, min(s0 + chunk, r1), dtype=np.int64)
        cols = np.empty((s.size, nb + 1), dtype=np.int64)
        mask = np.empty((s.size, nb + 1), dtype=bool)
        aligned = np.zeros(s.size, dtype=np.int64)
        for b in range(nb):
            i, j = b, (b + 1) % L
            anti = ((s >> i) & 1) != ((s >> j) & 1)
            cols[:, b] = s ^ ((1 << i) | (1 << j))
            mask[:, b] = anti
            aligned += ~anti
        cols[:, nb] = s
        mask[:, nb] = True
        vals = np.empty(cols.shape, dtype=np.float64)
        vals[:, :nb] = 0.5 * J
        vals[:, nb] = 0.25 * J * (aligned - (nb - aligned))
        order = np.argsort(np.where(mask, cols, np.iinfo(np.int64).max), axis=1, kind="stable")
        cols = np.take_along_axis(cols, order, axis=1)
        vals = np.take_along_axis(vals, order, axis=1)
        mask = np.take_along_axis(mask, order, axis=1)
        rp, c, v = _compress(cols, vals, mask)
        rowptrs.append(rp[1:] + base)
        base += rp[-1]
        colss.append(c)
        valss.append(v)
    return np.concatenate(rowptrs), np.concatenate(colss), np.concatenate(valss)


# Ground-state energies of the PBC ring (ARPACK, tol 1e-13).
HEISENBERG_RING_E0 = {8: -3.651093408937, 12: -5.387390917445, 16: -7.142296360617,
                      20: -8.904386529876, 24: -10.670014516537}


def hermitian_chain_csr(n):
    """sample_lanczos2.cpp:19-34: H(i,i+1) = -i, H(i+1,i) = +i; spectrum 2cos(k pi/(n+1))."""
    rows = np.arange(n)
    cols = np.stack([rows - 1, rows + 1], axis=1)
    mask = np.stack([rows > 0, rows < n - 1], axis=1)
    vals = np.broadcast_to(np.array([1j, -1j]), cols.shape)
    return _compress(cols, vals, mask)


def csr_bytes(n, nnz, scalar_bytes=8):
    """Algorithmic bytes of one CSR apply: values+indices once, row pointers,
    x once, y once."""
    ptr = 4 if nnz < (1 << 31) else 8
    return nnz * (scalar_bytes + 4) + (n + 1) * ptr + 2 * n * scalar_bytes
