"""b2nerf — host side of the B200-native NeRF hot path.

`lzzx-nerf_b200/` is put on sys.path (tests/conftest.py, bench.py, __graft_entry__.py do it) so that the
drop-in packages `gridencoder`, `raymarching`, `shencoder`, `freqencoder` shadow the reference's packages of the
same name and the reference's renderer.py / network.py / encoding.py run unchanged on top of them.
"""
from ._lib import lib, B2NError, declared_symbols, LIB_PATH  # noqa: F401
