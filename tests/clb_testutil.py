"""Shared builders for the parity tests: small seeded systems fed identically to oracle and engine."""
import numpy as np


def lj_table(rmax=3.0, dr=0.002, eps=1.0, sig=1.0, rc=2.5):
    """LJ 12-6 tabulated on the grid of examples/atrp_activator/table_MA_MA.pot (dr=0.002 from r=dr)."""
    r = dr * np.arange(1, int(round(rmax / dr)) + 1)
    sr6 = (sig / r) ** 6
    src6 = (sig / rc) ** 6
    e = 4 * eps * (sr6 * sr6 - sr6) - 4 * eps * (src6 * src6 - src6)
    f = 24 * eps * (2 * sr6 * sr6 - sr6) / r
    return r, e, f


def melt(n_side, rho=0.8442, seed=1, jitter=0.08, trimers=True):
    """n_side^3 beads on a jittered simple-cubic lattice, grouped as A-L-A trimers along x.

    Returns dict(pos, box, type, resid, bonds, angles). Types: 0 = A (reactive end), 1 = L (middle)."""
    rng = np.random.default_rng(seed)
    n = n_side ** 3
    L = (n / rho) ** (1.0 / 3.0)
    a = L / n_side
    g = np.arange(n_side)
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float64) * a + 0.5 * a
    pos += rng.uniform(-jitter, jitter, pos.shape) * a
    type_ = np.zeros(n, np.int32)
    resid = np.arange(n, dtype=np.int32)
    bonds, angles = [], []
    if trimers:
        ntri = n_side // 3
        idx = np.arange(n).reshape(n_side, n_side, n_side)
        rid = 0
        for k in range(n_side):
            for j in range(n_side):
                for t in range(ntri):
                    a0, a1, a2 = idx[k, j, 3 * t], idx[k, j, 3 * t + 1], idx[k, j, 3 * t + 2]
                    bonds += [(a0, a1), (a1, a2)]
                    angles.append((a0, a1, a2))
                    type_[a1] = 1
                    resid[[a0, a1, a2]] = rid
                    rid += 1
                for r in range(3 * ntri, n_side):
                    resid[idx[k, j, r]] = rid
                    rid += 1
    return dict(pos=pos, box=np.array([L, L, L]), type=type_, resid=resid,
                bonds=np.array(bonds, np.int64).reshape(-1, 2), angles=np.array(angles, np.int64).reshape(-1, 3))


def exclusions_from(bonds, angles):
    ex = [tuple(b) for b in bonds] + [(a[0], a[2]) for a in angles]
    return np.array(sorted(set((min(a, b), max(a, b)) for a, b in ex)), np.int64).reshape(-1, 2)


def rel_force_err(f, fref):
    """max_i |f_i - fref_i| / max(|fref_i|, rms|fref|): the norm used for the 1e-6 force bound."""
    d = np.linalg.norm(f - fref, axis=1)
    nr = np.linalg.norm(fref, axis=1)
    rms = np.sqrt((nr ** 2).mean())
    return float((d / np.maximum(nr, rms)).max())
