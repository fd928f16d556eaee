import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def small_world():
    """A small synthetic gene set with its oracle indices (transcripts + BSJ reference)."""
    import oracle
    import synth
    synth.build()
    oracle.build()
    T = synth.Txome(11, 240, p_antisense=0.05, p_repeat=0.08, p_polya=0.02)
    tx = oracle.Index.from_text(T.tx_text)
    bsj = oracle.Index.from_text(T.bsj_text)
    return T, tx, bsj
