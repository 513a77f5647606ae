"""Seeded synthetic inputs for the R_q hot path.

One host-side generator produces the key, messages and all randomness (r, y, d)
as flat arrays; the same arrays are fed to the CUDA engine and to the CPU oracle,
which is how "randomness drawn host-side from a seeded RNG and fed identically to
both implementations" is met without the reference's `rand` stream.

Distributions follow the reference's samplers:
  key blocks, x, g : uniform in [-(Q//2), Q//2]        commit.rs:41,53; tests/test.rs:95-97
  r                : uniform in [-b, b]                commit.rs:101, polynomial.rs:14-25
  y                : trunc(N(0, sigma)) per coeff      polynomial.rs:28-44, open.rs:88-94
  d                : min(kappa, N) entries +-1, shuffled  challenge_space.rs:12-33
"""
from __future__ import annotations

from math import isqrt

import numpy as np

Q_DEFAULT = 3515337053


def sigma(N, k=3, kappa=36, b=1):
    """params.rs:94-98"""
    return b * (11 * kappa) * isqrt(k * N)


class Synth:
    def __init__(self, seed, N=512, Q=Q_DEFAULT, n=1, k=3, l=1, kappa=36, b=1):
        self.rng = np.random.Generator(np.random.Philox(seed))
        self.N, self.Q, self.n, self.k, self.l, self.kappa, self.b = N, Q, n, k, l, kappa, b
        self.half = Q // 2

    def uniform_q(self, *shape):
        return self.rng.integers(-self.half, self.half + 1, size=shape + (self.N,), dtype=np.int64).astype(np.int32)

    def key(self):
        """(a1p [n][k-n][N], a2p [l][k-n-l][N]) random blocks of the commitment key."""
        return (self.uniform_q(self.n, self.k - self.n),
                self.uniform_q(self.l, self.k - self.n - self.l))

    def message(self, B, ragged=False):
        """x [B][l][N]; ragged=True zero-pads a random length in 1..=N like tests/test.rs:95-99."""
        x = self.uniform_q(B, self.l)
        if ragged:
            lens = self.rng.integers(1, self.N + 1, size=(B, self.l))
            mask = np.arange(self.N)[None, None, :] < lens[:, :, None]
            x = np.where(mask, x, 0).astype(np.int32)
        return x

    def scalar(self, *shape):
        """g [..][N] uniform full-size polynomials (prepare_scalar inputs)."""
        return self.uniform_q(*shape)

    def small(self, *shape):
        """r [..][k][N] uniform in [-b, b], int8."""
        return self.rng.integers(-self.b, self.b + 1, size=shape + (self.k, self.N), dtype=np.int64).astype(np.int8)

    def gaussian(self, *shape):
        """y [..][k][N]: truncation toward zero of N(0, sigma)."""
        s = sigma(self.N, self.k, self.kappa, self.b)
        v = self.rng.normal(0.0, float(s), size=shape + (self.k, self.N))
        return np.trunc(v).astype(np.int32)

    def challenge(self, B):
        """d [B][N]: min(kappa, N) entries +-1 at shuffled positions, int8."""
        N = self.N
        nnz = min(self.kappa, N)
        d = np.zeros((B, N), np.int8)
        signs = self.rng.integers(0, 2, size=(B, nnz)).astype(np.int8) * 2 - 1
        # random positions without replacement == shuffle of a vector with nnz leading non-zeros
        pos = np.argsort(self.rng.random((B, N)), axis=1)[:, :nnz]
        np.put_along_axis(d, pos, signs, axis=1)
        return d
