import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_terminal_summary(terminalreporter):
    """Name every tensor that passed through assert_parity's noise-floor branch (ill-conditioned in fp32: judged against
    the float64 evaluation instead of the fp32 reference), so that a regression cannot hide behind it."""
    from tests.helpers import NOISE_FLOOR_LOG
    if NOISE_FLOOR_LOG:
        terminalreporter.write_sep("-", f"assert_parity: {len(NOISE_FLOOR_LOG)} tensor(s) judged against float64 (noise-floor rule)")
        for line in NOISE_FLOOR_LOG:
            terminalreporter.write_line("  " + line)
