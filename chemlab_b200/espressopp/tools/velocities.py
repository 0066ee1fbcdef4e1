"""tools.velocities.gaussian(T, N, mass): src/start_simulation.py:136-146."""
import numpy as np


def gaussian(T, N, mass, kb=1.0, seed=None):
    rng = np.random.default_rng(seed)
    m = np.asarray(mass, float)
    v = rng.normal(0.0, 1.0, (N, 3)) * np.sqrt(kb * T / m)[:, None]
    v -= (v * m[:, None]).sum(0) / m.sum()
    ek = 0.5 * (m[:, None] * v * v).sum()
    v *= np.sqrt(1.5 * N * kb * T / ek)
    return v[:, 0].tolist(), v[:, 1].tolist(), v[:, 2].tolist()
