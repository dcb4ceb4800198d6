# restores the exports the reference's callers rely on (renderer.py:12 uses `raymarching.march_rays...`; the
# reference snapshot ships an empty __init__.py, its stale .pyc shows `from .raymarching import *`)
from .raymarching import *  # noqa: F401,F403
