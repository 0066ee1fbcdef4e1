"""tools.decomp.nodeGrid / cellGrid / tuneSkin: src/start_simulation.py:152-163,716-721."""


def nodeGrid(n, *a, **k):
    """near-cubic factorisation of n ranks.  The engine itself uses a slab (1,1,n) decomposition."""
    best = (n, 1, 1)
    for x in range(1, n + 1):
        if n % x:
            continue
        for y in range(1, n // x + 1):
            if (n // x) % y:
                continue
            z = n // x // y
            if max(x, y, z) - min(x, y, z) < max(best) - min(best):
                best = (x, y, z)
    return tuple(sorted(best, reverse=True))


def cellGrid(box, node_grid, rc, skin, *a, **k):
    """int(L / (nodeGrid * (rc + skin))) per dimension (log of examples/atrp_lj/single:34-42)."""
    return tuple(max(1, int(box[d] / (node_grid[d] * (rc + skin)))) for d in range(3))


def tuneSkin(system, integrator, minSkin=0.01, maxSkin=1.5, precision=0.001, printInfo=True):
    return system.skin
