# the reference's callers do `from gridencoder import GridEncoder` (encoding.py:24)
from .grid import GridEncoder, grid_encode  # noqa: F401
