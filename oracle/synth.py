"""Synthetic snapshot matrices -- numpy side.  TEST INFRASTRUCTURE.

X[i,j] = (1 + 0.3 sin(2 pi x_i)) + sum_{k<K} a_k phi_k(x_i) psi_k(t_j) + eps * h(i,j,seed)
  x_i = (i+1/2)/m, t_j = j/n, phi_k = sin(2 pi (k+1) x_i + theta_k), theta_k = 0.37 k,
  psi_k = cos / sin pairs at frequency ceil((k+1)/2), a_k = 10^(-6k/K), K = min(n,32),
  h = counter hash of (seed,i,j) mapped to U(-1/2,1/2), eps = 1e-8.
Any row slice [r0,r1) is reproducible without generating the rest, which is how the
row-sharded ranks build their shard.  ``bench.py`` has the same generator in torch for the
device side (same formulas; sin/cos differ by ulps, so parity tests always ship the numpy
matrix to the device instead of regenerating it there).
"""
import numpy as np

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def _hash01(seed, i, j):
    """splitmix64 finaliser of (seed, i, j) -> U(-1/2, 1/2) with 53 random bits."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * _M1) ^ (i.astype(np.uint64) * _M2)[:, None] ^ (j.astype(np.uint64) * _M3)[None, :]
        z = z + _M1
        z = (z ^ (z >> np.uint64(30))) * _M2
        z = (z ^ (z >> np.uint64(27))) * _M3
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


def snapshots(m, n, seed, r0=0, r1=None, eps=1e-8, nvars=1):
    """Rows [r0, r1) of the m x n synthetic snapshot matrix (fp64, C order)."""
    r1 = m if r1 is None else r1
    i = np.arange(r0, r1, dtype=np.int64)
    j = np.arange(n, dtype=np.int64)
    x = ((i // nvars) + 0.5) / (m // nvars)
    v = (i % nvars).astype(np.float64)
    t = j / n
    K = min(n, 32)
    X = np.repeat((1.0 + 0.3 * np.sin(2 * np.pi * x))[:, None], n, axis=1)
    for k in range(K):
        a = 10.0 ** (-6.0 * k / K) * (1.0 + 0.25 * v)
        f = (k + 2) // 2
        psi = np.cos(2 * np.pi * f * t) if k % 2 == 0 else np.sin(2 * np.pi * f * t)
        phi = np.sin(2 * np.pi * (k + 1) * x + 0.37 * k)
        X += (a * phi)[:, None] * psi[None, :]
    X += eps * _hash01(seed, i, j)
    return np.ascontiguousarray(X)


def random_matrix(m, n, seed, cond=None):
    """Dense random test matrix; with ``cond`` the singular values decay geometrically."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    if cond is not None:
        s = cond ** (-np.arange(n) / max(n - 1, 1))
        Vq, _ = np.linalg.qr(rng.standard_normal((n, n)))
        Uq, _ = np.linalg.qr(A)
        A = (Uq * s) @ Vq.T
    return np.ascontiguousarray(A)


def dmd_waves(m, n, seed, npairs=4, dt=0.1, noise=1e-6):
    """Snapshots of a linear system with `npairs` damped / growing travelling waves (distinct frequencies and decay
    rates, so the DMD eigenvalues are well separated conjugate pairs) + a steady offset + hashed noise."""
    i = np.arange(m, dtype=np.int64)
    j = np.arange(n, dtype=np.int64)
    x = (i + 0.5) / m
    t = j * dt
    X = np.repeat((1.0 + 0.3 * np.sin(2 * np.pi * x))[:, None], n, axis=1)
    for k in range(npairs):
        amp = 2.0 ** (-k)
        om = 2.0 + 1.7 * k
        sig = -0.05 * (k + 1) + (0.04 if k == 1 else 0.0)
        phase = 2 * np.pi * (k + 1) * x + 0.5 * k
        X += amp * np.exp(sig * t)[None, :] * np.cos(phase[:, None] - om * t[None, :])
    X += noise * _hash01(seed, i, j)
    return np.ascontiguousarray(X)
