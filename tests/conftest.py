import lzma
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = sorted(f[:-7] for f in os.listdir(GOLD) if f.endswith(".rec.xz"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_case(name, max_frames=None):
    from vp8b200 import recfile
    rec = recfile.parse(lzma.decompress(open(os.path.join(GOLD, name + ".rec.xz"), "rb").read()), max_frames)
    md5s = open(os.path.join(GOLD, name + ".md5")).read().split()
    return rec, md5s


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library, loaded for real: a GPU test must never pass on a fallback."""
    from vp8b200 import abi
    L = abi.lib()
    assert L.vp8b200_device_count() > 0, "no CUDA device visible to libvp8b200.so"
    return L
