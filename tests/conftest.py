import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='session')
def synth_frame():
    """Body + frame used by tests/golden/gen_golden.py (pose seed 7, |theta| <= 0.4)."""
    from avatarcap_b200 import synth
    body = synth.SynthBody()
    frame = synth.make_frame(body, synth.random_pose(7, 0.4))
    return body, frame
