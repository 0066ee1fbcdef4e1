from . import decomp, velocities, convert, analyse  # noqa: F401
from .analyse import info  # noqa: F401,E402
