import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lzzx-nerf_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def scene_bitfield():
    from b2nerf import scene
    return scene.bitfield_from_grid(scene.density_grid())


def load_ref_ext(name):
    """Import one of the reference's own extensions built by oracle/build_ref_ext.sh (GPU parity oracle). None if absent."""
    import glob
    import importlib.util
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + ".*.so"))
    if not hits:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
