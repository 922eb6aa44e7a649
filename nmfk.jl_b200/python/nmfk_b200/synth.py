"""Synthetic nonnegative mixtures of the shapes BASELINE.json names:
X = W0 * H0 with W0 ~ U(0,1)^{n x k0}, H0 ~ U(0,1)^{k0 x m}, optional uniform noise, cast to T.
NumPy Philox streams so that the CPU oracle and the GPU path see identical bits."""
import numpy as np

CONFIGS = {
    # name: (n, m, k0, dtype, krange, nNMF)
    "C2": (1000, 200, 5, np.float64, list(range(2, 11)), 100),
    "C3": (10000, 10000, 16, np.float32, [16], 64),
    "C4": (100000, 2000, 8, np.float64, list(range(2, 33)), 256),
    "C5": (2000000, 1000, 24, np.float32, [24], 32),
}


def mixture(n, m, k0, seed=2015, dtype=np.float64, noise=0.0):
    rng = np.random.Generator(np.random.Philox(key=seed))
    W0 = rng.random((n, k0))
    H0 = rng.random((k0, m))
    X = W0 @ H0
    if noise:
        X = X + noise * rng.random((n, m))
    return np.asfortranarray(X.astype(dtype))


def readme_bss(seed=2015):
    """Readme.md:97-106: a,b,c ~ U(0,1)^15, X = [a+3c, 10a+b, b, 5b+c, a+2b+5c] (config C1)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    a, b, c = rng.random(15), rng.random(15), rng.random(15)
    return np.asfortranarray(np.stack([a + 3 * c, 10 * a + b, b, 5 * b + c, a + 2 * b + 5 * c], axis=1))


def philox_inits(seed0, R, n, k, m, dtype=np.float64):
    """Restart i (1-based) draws W (column-major) then H from Philox(key=seed0+i): the streams
    nmfk_batch_init_random generates on the device."""
    W = np.empty((R, n, k), dtype=dtype)
    H = np.empty((R, k, m), dtype=dtype)
    for i in range(1, R + 1):
        rng = np.random.Generator(np.random.Philox(key=seed0 + i))
        W[i - 1] = rng.random(n * k).reshape((n, k), order="F")
        H[i - 1] = rng.random(k * m).reshape((k, m), order="F")
    return W, H
