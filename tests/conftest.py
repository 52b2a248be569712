import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a GPU test must never silently pass on a box without a GPU
    try:
        import lbaudiodetective_b200 as lb
        have = lb.device_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- figures: what the parity tests measured (worst errors, mismatch rates), kept even when pytest runs without -s ----
_FIGURES = []


@pytest.fixture
def figure(request):
    """figure("text") records a measured number under the test's name; all of them are printed in the terminal summary (so the
    driver's pytest.log keeps them) and written to gpurun_out/pytest_figures.json when that directory exists."""
    def rec(text):
        _FIGURES.append((request.node.nodeid, str(text)))
        print(text)
    return rec


def pytest_terminal_summary(terminalreporter):
    if not _FIGURES:
        return
    terminalreporter.section("measured figures (parity tests)")
    for node, text in _FIGURES:
        terminalreporter.write_line("%s: %s" % (node.split("::", 1)[-1], text))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        try:
            json.dump([{"test": n, "figure": t} for n, t in _FIGURES], open(os.path.join(out, "pytest_figures.json"), "w"), indent=1)
        except OSError:
            pass


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("compiled reference (oracle/_ref) not available")
    return Ref()


@pytest.fixture(scope="session")
def checker():
    """Strongest oracle available: the compiled reference if built, else the C restatement."""
    from oracle.oracle import best
    return best()


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def config1():
    return np.load(os.path.join(GOLDEN, "config1.npz"))


@pytest.fixture(scope="session")
def compare_cases():
    z = np.load(os.path.join(GOLDEN, "compare_cases.npz"))
    return [{k: z["%s_%d" % (k, i)] for k in ("c1", "c2", "L", "range", "fp1", "fp2", "score")} for i in range(int(z["n"]))]


@pytest.fixture(scope="session")
def sweep():
    return np.load(os.path.join(GOLDEN, "sweep.npz"))


@pytest.fixture(scope="session")
def lb():
    import lbaudiodetective_b200 as m
    m.load_library()
    return m
