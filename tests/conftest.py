"""Shared fixtures.

``api`` is the module-like surface the transcribed reference tests run against:
  * "oracle" — oracle/lle_oracle.py (CPU restatement; runs everywhere)
  * "cuda"   — lle_b200 (the product: C-ABI + sm_100a kernels; needs a B200, marked ``gpu``)
so every known-answer test of the reference pins the oracle on CPU *and* checks the CUDA path
through the same assertions on the GPU box.
"""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _oracle_api():
    from oracle import lle_oracle

    return lle_oracle


def _cuda_api():
    import lle_b200

    return lle_b200


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def api(request):
    return _oracle_api() if request.param == "oracle" else _cuda_api()


@pytest.fixture(scope="session")
def layouts():
    with open(os.path.join(GOLDEN, "layouts.json")) as f:
        return json.load(f)


