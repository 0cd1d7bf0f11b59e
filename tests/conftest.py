import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EXAMPLE = os.path.join(GOLDEN, "example_infiles")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the gpu-marked tests instead of failing in mbsv_problem_create
    (the library has no CPU fallback).  `-m gpu` on such a box still fails loudly: that selection asks for the device."""
    if config.getoption("-m") and "gpu" in config.getoption("-m") and "not gpu" not in config.getoption("-m"):
        return
    try:
        import multibreak_sv_b200 as m
        have = m.load_library().mbsv_device_count() > 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the sampler has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def example_problem(oracle):
    """The shipped example (MultiBreak-SV-example/infiles), parsed by the oracle's restatement of the parser."""
    return oracle.load_files(os.path.join(EXAMPLE, "gasv.clusters"), os.path.join(EXAMPLE, "assignments.txt"),
                             os.path.join(EXAMPLE, "experiments.txt"))


@pytest.fixture(scope="session")
def mbsv():
    import multibreak_sv_b200 as m
    m.load_library()
    return m


def to_packed(mbsv, prob):
    """oracle.Problem -> multibreak_sv_b200.PackedProblem (same field names; the arrays are shared)."""
    return mbsv.PackedProblem(dict(prob.scalars), dict(prob.arrays))
