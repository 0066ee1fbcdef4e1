from . import decomp, velocities, convert, analyse  # noqa: F401
