import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    return any(os.path.exists("/dev/nvidia%d" % i) for i in range(16))


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of erroring in their fixtures
    (on a GPU box nothing is skipped: a missing CUDA library must fail loudly there)."""
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this box (gpu-marked tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "golden_3bx32.npz"))


@pytest.fixture(scope="session")
def golden_weights_bin():
    return os.path.join(GOLDEN, "ref_3bx32.bin.txt")


@pytest.fixture(scope="session")
def golden_weights_txt():
    return os.path.join(GOLDEN, "ref_3bx32.txt")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def golden_blocks():
    """Second fixture: BottleneckBlock / NestedBottleneckBlock [-SE] tower, outputs of the
    UNMODIFIED compiled reference (tests/golden/make_golden.py blocks)."""
    import numpy as np
    return np.load(os.path.join(GOLDEN, "golden_btl_5bx32.npz"))


@pytest.fixture(scope="session")
def golden_blocks_weights():
    return os.path.join(GOLDEN, "ref_btl_5bx32.bin.txt")


@pytest.fixture(scope="session")
def golden_mixer():
    """Third fixture: MixerBlock[-SE] tower (depthwise 7x7 / 5x5 + FFN) with the RepLK policy head, outputs of the
    UNMODIFIED compiled reference (tests/golden/make_golden.py blocks)."""
    import numpy as np
    return np.load(os.path.join(GOLDEN, "golden_mix_4bx32.npz"))


@pytest.fixture(scope="session")
def golden_mixer_weights():
    return os.path.join(GOLDEN, "ref_mix_4bx32.bin.txt")
