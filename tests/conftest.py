import pytest

from paif_testutil import GOLDEN_CASES, load_golden


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
