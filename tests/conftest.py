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
    # a fresh checkout has no built library (it is git-ignored): build it in-tree before anything imports the package.
    # nvcc cross-compiles without a GPU; where there is no nvcc either, the imports below fail loudly, as they should.
    import importlib.util
    import shutil

    lib = os.path.join(ROOT, "lle_b200", "_native", "liblle_b200.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        spec = importlib.util.spec_from_file_location("_lle_b200_build", os.path.join(ROOT, "lle_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


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


