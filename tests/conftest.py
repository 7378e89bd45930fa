import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail loudly rather than silently pass on some fallback:
    # the tests themselves construct api.Context, which raises SqmpError(NODEVICE).
    pass


@pytest.fixture(scope="session")
def fixture_data():
    from squirrel_motion_planner_b200 import robot_desc as rd
    lt, fx = rd.load_fixture()
    spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
    return lt, fx, spheres


@pytest.fixture(scope="session")
def orc(fixture_data):
    import oracle
    lt, fx, spheres = fixture_data
    return oracle.Oracle(lt, spheres)


@pytest.fixture(scope="session")
def small_world():
    from squirrel_motion_planner_b200 import robot_desc as rd, synthetic as syn
    w = syn.box_world(g=96, n_boxes=20, seed=5, dims=(96, 80, 64))
    origin = w.origin + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET])
    return w, origin


@pytest.fixture(scope="session")
def ctx(fixture_data):
    """One CUDA context for the whole GPU session (raises, never falls back, if the library/GPU is missing)."""
    from squirrel_motion_planner_b200 import api, robot_desc as rd
    lt, fx, spheres = fixture_data
    c = api.Context(rd.compile_robot(lt, spheres), 0)
    yield c
    c.close()


def unpack(mask, n):
    return np.unpackbits(np.ascontiguousarray(mask).view(np.uint8), bitorder="little")[:n].astype(bool)
