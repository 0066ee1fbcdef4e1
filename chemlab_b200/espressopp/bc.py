"""bc.OrthorhombicBC(rng, box): src/start_simulation.py:162."""


class OrthorhombicBC:
    def __init__(self, rng, boxL):
        self.rng = rng
        self.boxL = tuple(float(x) for x in boxL)
