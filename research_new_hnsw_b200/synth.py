"""Synthetic data laws of the benchmark configurations; pure numpy, deterministic per seed."""
import numpy as np


def lowrank_data(n, d, seed, latent=16, noise=0.1, proj_seed=123, normalize=False):
    """'SIFT-shaped' / 'Deep-shaped' rows: z ~ N(0, I_latent), x = z A + noise * eps, A fixed by proj_seed.
    i.i.d. Gaussian data (the reference's own generator, build.cpp:124-138) cannot reach recall 0.95 in the
    ef 32-256 sweep, so the recall-driven metric uses this low-intrinsic-dimension law."""
    A = np.random.default_rng(proj_seed).standard_normal((latent, d), dtype=np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), np.float32)
    step = 1 << 18
    for s in range(0, n, step):
        e = min(n, s + step)
        z = rng.standard_normal((e - s, latent), dtype=np.float32)
        x = z @ A + np.float32(noise) * rng.standard_normal((e - s, d), dtype=np.float32)
        if normalize:
            x /= np.linalg.norm(x, axis=1, keepdims=True)
        out[s:e] = x
    return out
