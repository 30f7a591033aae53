"""The C ABI used from a plain-C host program (examples/c_client.c): no Python, torch or CUDA headers on the host side —
the shape of the reference-side FFI binding."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exe():
    import __graft_entry__

    return __graft_entry__.build_c_client()


def test_c_client_builds_against_the_public_header_only():
    exe = _exe()
    assert os.access(exe, os.X_OK)
    needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True).stdout
    assert "liblle_b200.so" in needed and "libcudart" not in needed and "libtorch" not in needed


@pytest.mark.gpu
def test_c_client_runs():
    res = subprocess.run([_exe(), "4096", "150"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = res.stdout.strip().splitlines()
    assert lines[-1] == "ok" and "episodes finished" in lines[1] and "0 unexpected transitions" in lines[2]
    assert lines[-2].startswith("generated maps") and "64 maps x 16 envs" in lines[-2]


@pytest.mark.gpu
def test_quickstart_example_runs():
    import sys

    res = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "quickstart.py"), "1024"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().splitlines()[-1] == "ok" and "generated maps: 64 distinct" in res.stdout


@pytest.mark.gpu
def test_c_closed_loop_runs():
    """examples/c_closed_loop.c: the compiled-host closed loop bench.py reports as e2e - over parts of one batch (lle_vec_parts_*,
    modes s1 / s3 / s8) and over sub-batch vecs (lle_vec_pipeline_submit/_wait, 1 / 2 / 4 in flight); it checks itself that the host
    saw exactly the recorded done flags, that no env refused an action and that the device's final done flags match."""
    import json

    import __graft_entry__

    exe = __graft_entry__.build_c_client("c_closed_loop")
    res = subprocess.run([exe, "0", "8192", "40", "s1", "s3", "s8", "1", "2", "4"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert set(out["parts"]) == {"s1", "s3", "s8", "1", "2", "4"}
    for v in out["parts"].values():
        assert v["mismatches"] == 0 and v["env_errors"] == 0 and v["env_steps_per_s"] > 0
