"""AddressSanitizer + UndefinedBehaviorSanitizer over the product's host code that reads untrusted input (the map compiler and the
TOML reader) and over the per-world core of the tiny-map kernel (tests/host_shim/sanitize_main.cpp).  compute-sanitizer is closed on
the GPU pool this was developed on; this is the host-side half of that check.  CPU suite."""
import json
import os
import shutil
import subprocess

import pytest

from _util import GOLDEN, level_text

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "lle_b200", "csrc")

MALFORMED = [
    "", "\n\n", "S0", "S0 X\nS1", "S0 . X\n. .", "L0E S0 X", "S0 L9Q X", "S0 S0 X X", "G G G", "S0 X L0", "@ @\n@ @", "S0 . . X " * 40,
    "S1 X", "S0 X\x00Y", "L0N\nS0\nX", "width = 3\nheight = 2\n[[agents]]\nstart_positions = [{i = 9, j = 9}]\n", "world_string = '''S0 X'''\n[[lasers]]\n",
    "[agents]\nx = [[[[[[", "S0 " + "L0E " * 70 + "X",
]


def test_host_code_under_asan_ubsan(tmp_path, layouts):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "sanitize_main")
    srcs = [os.path.join(HERE, "host_shim", "sanitize_main.cpp"), os.path.join(HERE, "host_shim", "tiny_host.cpp"),
            os.path.join(CSRC, "map_compiler.cpp"), os.path.join(CSRC, "toml_config.cpp")]
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                            "-x", "c++"] + srcs + ["-o", exe], capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr.lower() and "cannot find" in build.stderr.lower():
        pytest.skip("the sanitizer runtimes are not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    files = []
    texts = [level_text(n) for n in range(1, 7)] + [t for _, t in sorted(layouts.items())][:12] + MALFORMED
    # seeded mutations of valid maps: replaced / deleted / duplicated tokens and characters
    import random

    rng = random.Random(20261018)
    alphabet = list("SLGX@.V0123456789NESW \n\t[]=,'\"-") + ["S0", "L0E", "L1S", "X", "G", "@", ".", "V", "\n", "S9", "L3W"]
    base = [level_text(n) for n in (1, 3, 4, 6)] + [t for _, t in sorted(layouts.items())][:6]
    for k in range(240):
        t = list(rng.choice(base))
        for _ in range(rng.randint(1, 6)):
            pos = rng.randrange(len(t) + 1)
            op = rng.random()
            if op < 0.4 and t:
                t[min(pos, len(t) - 1)] = rng.choice(alphabet)
            elif op < 0.7 and t:
                del t[min(pos, len(t) - 1)]
            else:
                t.insert(pos, rng.choice(alphabet))
        texts.append("".join(t))
    toml_dir = os.path.join(GOLDEN, "toml")
    if os.path.isdir(toml_dir):
        texts += [open(os.path.join(toml_dir, f)).read() for f in sorted(os.listdir(toml_dir))[:8]]
    for k, t in enumerate(texts):
        p = tmp_path / f"map{k}.txt"
        p.write_bytes(t.encode("utf-8", "replace") if isinstance(t, str) else t)
        files.append(str(p))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    run = subprocess.run([exe] + files, capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0, (run.stdout[-500:], run.stderr[-3000:])
    assert "runtime error" not in run.stderr and "AddressSanitizer" not in run.stderr, run.stderr[-3000:]
    out = json.loads(run.stdout.strip().splitlines()[-1])
    assert out["compiled"] >= 6 * 18 and out["rejected"] >= 6 * 20 and out["stepped"] >= 12, out
