import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "hostsim"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rc():
    import raytracing_cuda
    return raytracing_cuda


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def hostsim():
    import hostsim_py
    hostsim_py.build()
    return hostsim_py


def load_scene(name, width=None, height=None):
    import raytracing_cuda as rc
    sc = rc.Scene.load_npz(os.path.join(GOLDEN, "scenes", name + ".npz"))
    if width is not None:
        sc.camera = sc.camera.with_raster_size(width, height)
    return sc


def bunny_mesh():
    import numpy as np
    import raytracing_cuda as rc
    z = np.load(os.path.join(GOLDEN, "scenes", "bunny_mesh.npz"))
    return rc.Mesh(z["vertices"], z["tris"], z["normals"], None)
