import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'slow: long-running CPU test')
