from . import gromacs  # noqa: F401
