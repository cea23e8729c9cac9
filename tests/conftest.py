import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def has_reference():
    return os.path.isfile("/root/reference/look2hear/models/mossformer2.py")


needs_reference = pytest.mark.skipif(not has_reference(), reason="reference tree only exists in the build container")
