from .sphere_harmonics import SHEncoder, sh_encode  # noqa: F401  (encoding.py:20 does `from shencoder import SHEncoder`)
