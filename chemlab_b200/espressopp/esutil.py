"""esutil.RNG(seed): src/start_simulation.py:149.  The engine draws from counter-based Philox streams keyed by this seed."""
import random


class RNG:
    def __init__(self, seed=0):
        self.seed = int(seed)
        self._r = random.Random(self.seed)

    def __call__(self):
        return self._r.random()

    def normal(self):
        return self._r.gauss(0.0, 1.0)
