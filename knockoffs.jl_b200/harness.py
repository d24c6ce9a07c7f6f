"""Synthetic inputs of the BASELINE.json configs.  Harness code: seeded numpy generators
that follow the reference's simulators (src/utilities.jl:357-380 simulate_AR1 with fixed rho,
:483-508 simulate_block_covariance).  No reference data is read at run time."""
from __future__ import annotations

import numpy as np


def ar1_sigma(p: int, rho: float) -> np.ndarray:
    """Σ_ij = rho^|i-j| (simulate_AR1 with constant rho, src/utilities.jl:357-380)."""
    idx = np.arange(p)
    return np.asfortranarray(rho ** np.abs(idx[:, None] - idx[None, :]).astype(np.float64))


def block_sigma(p: int, block: int, rho: float, gamma: float) -> np.ndarray:
    """simulate_block_covariance (src/utilities.jl:483-508): 1 on the diagonal, rho within a block of
    `block` consecutive variables, gamma*rho between blocks."""
    g = np.arange(p) // block
    S = np.where(g[:, None] == g[None, :], rho, gamma * rho).astype(np.float64)
    np.fill_diagonal(S, 1.0)
    return np.asfortranarray(S)


def gaussian_X(n: int, Sigma: np.ndarray, seed: int, mu=None) -> np.ndarray:
    """X = Z0 chol(Σ)' (+ μ), Z0 ~ N(0,1)^{n x p}, numpy PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = np.linalg.cholesky(Sigma)
    X = rng.standard_normal((n, Sigma.shape[0])) @ L.T
    if mu is not None:
        X += np.asarray(mu)[None, :]
    return np.asfortranarray(X)


def dosage_X(n: int, p: int, seed: int, block: int = 50) -> np.ndarray:
    """Genotype-like 0/1/2 dosages with block LD (config 5): two haplotypes per row; within a block of
    `block` SNPs a haplotype copies its left neighbour's latent uniform with prob. 0.9 (AR-like LD)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = rng.uniform(0.05, 0.5, size=p)
    X = np.zeros((n, p), order="F")
    for _ in range(2):
        u = rng.uniform(size=(n, p))
        keep = rng.uniform(size=(n, p)) < 0.9
        for j in range(1, p):
            if j % block:
                u[:, j] = np.where(keep[:, j], u[:, j - 1], u[:, j])
        X += (u < f[None, :])
    return X


CONFIGS = {
    "c1": dict(desc="AR(1) rho=0.5 p=500 n=1000 :mvr m=1", p=500, n=1000, method="mvr", m=1, sigma=("ar1", 0.5)),
    "c2": dict(desc="AR(1) rho=0.5 p=2000 n=10000 :maxent m=1", p=2000, n=10000, method="maxent", m=1, sigma=("ar1", 0.5)),
    "c3": dict(desc="block-equicorrelated p=5000 (blocks of 10, rho=0.5, gamma=0.25) n=100000 :mvr/:sdp_ccd m=5",
               p=5000, n=100000, method="mvr", m=5, sigma=("block", 10, 0.5, 0.25)),
}


def config_sigma(name: str, p: int | None = None) -> np.ndarray:
    c = CONFIGS[name]
    p = p or c["p"]
    kind = c["sigma"]
    if kind[0] == "ar1":
        return ar1_sigma(p, kind[1])
    return block_sigma(p, kind[1], kind[2], kind[3])
