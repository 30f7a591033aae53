"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/lle_b200.h declares, and its host-side map compiler agrees with the oracle on the whole corpus.
No kernel is launched here."""
import ctypes as C
import os
import re

import pytest

from _util import level_text
from oracle import lle_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _native():
    from lle_b200 import _native
    return _native


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lle_b200.h")).read()
    declared = re.findall(r"LLE_API\s+[\w\s\*]+?\b(lle_\w+)\s*\(", header)
    assert len(declared) >= 24
    L = C.CDLL(_native().LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/lle_b200.h but not exported"
    assert sorted(declared) == sorted(_native().SYMBOLS)
    assert b"sm_100a" in _native().lib().lle_version()


def test_no_cpu_fallback_without_a_device():
    import torch

    import lle_b200

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lle_b200.VecWorld(lle_b200.Map(level=1), 4)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lle_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "lle_oracle" not in text and "oracle/" not in text.replace("oracle/ is test", ""), f


def _map_facts_native(text=None, level=None):
    import lle_b200

    m = lle_b200.Map(text, level=level)
    return dict(dims=(m.height, m.width, m.n_agents, m.n_gems, m.n_sources), walls=m.walls, voids=m.voids, exits=m.exits,
                gems=m.gems, starts=m.starts, laser_cells=m.laser_cells,
                sources=[(s.pos, s.agent_id, int(s.direction), s.laser_id, s.beam_len) for s in m.sources()],
                lasers=[(pos, lid, colour, int(d)) for pos, lid, colour, d, _, _ in m.laser_tiles()])


def _map_facts_oracle(text):
    w = lo.World(text)
    return dict(dims=(w.height, w.width, w.n_agents, w.n_gems, w.n_sources), walls=w.wall_pos, voids=w.void_pos, exits=w.exit_pos,
                gems=[g.pos for g in w.gems], starts=w.start_pos, laser_cells=w.laser_pos,
                sources=[(s.pos, s.agent_id, int(s.direction), s.laser_id, s.beam_len) for s in w.laser_sources],
                lasers=[(l.pos, l.laser_id, l.agent_id, int(l.direction)) for l in w.lasers])


def test_map_compiler_matches_oracle_on_corpus(layouts):
    texts = [level_text(n) for n in range(1, 7)] + [t for _, t in sorted(layouts.items())]
    texts += ["S0 L1E X", ".   L0S S1\nS0   .   .\nL1E  X   X", "L0E . L0E . X S0",
              "\n".join([" .   .   . . . ."] + [f"S{k}  L{k}W  . . . X" for k in range(14)])]
    for text in texts:
        assert _map_facts_native(text) == _map_facts_oracle(text)
    for n in range(1, 7):
        assert _map_facts_native(level=n) == _map_facts_oracle(level_text(n))


def test_map_compiler_matches_oracle_on_random_maps():
    """Differential fuzzing of the host map compiler against the oracle's parser: 1,500 random maps (random sizes, agents,
    crossing beams, colours >= n_agents, starts on beams ...), some of them invalid on purpose — same facts or same error."""
    import random

    import lle_b200
    from _util import random_map

    rng = random.Random(20261018)
    n_ok = n_err = 0
    for _ in range(1500):
        text = random_map(rng, allow_invalid=True)
        try:
            expect = _map_facts_oracle(text)
        except lo.ParsingError as e:
            with pytest.raises(lle_b200.ParsingError, match=str(e).split(" ")[0].split("{")[0]):
                lle_b200.Map(text)
            n_err += 1
            continue
        assert _map_facts_native(text) == expect, text
        n_ok += 1
    assert n_ok > 1000 and n_err > 100


def test_map_errors():
    import lle_b200

    for text, name in [("", "EmptyWorld"), ("S0 S0 X X", "DuplicateStartTile"), ("S1 S0 X", "NotEnoughExitTiles"),
                       (". . G", "NoAgents"), ("X S0 .\n . .", "InconsistentDimensions"), ("S0 Q X", "InvalidTile"),
                       ("S0  S1 X . X\nL1N .  . . .", "AgentWithoutStart")]:
        with pytest.raises(lle_b200.ParsingError, match=name):
            lle_b200.Map(text)
    with pytest.raises(lle_b200.InvalidLevelError):
        lle_b200.Map(level=7)


def test_argument_validation_precedes_any_device_call():
    """Empty, oversized and inconsistent inputs are refused with LLE_INVALID_ARGUMENT (202) before a device is touched, so this
    runs without a GPU: zero worlds, zero maps, null pointers, a map index out of range, maps of different shapes, a reward
    dimension the reference does not have; generator batches beyond capacity, zero attempts, unknown label bits."""
    N = _native()
    L = N.lib()
    m1, m6, small = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert L.lle_map_level(1, C.byref(m1)) == 0 and L.lle_map_level(6, C.byref(m6)) == 0
    text = b"S0 . X"
    assert L.lle_map_parse(text, len(text), C.byref(small)) == 0
    opts = N.VecOptions()
    L.lle_vec_default_options(C.byref(opts))
    out = C.c_void_p()
    one = (C.c_void_p * 1)(m6)

    def create(maps, n_maps, map_of_env, n_envs, o=opts):
        return L.lle_vec_create(maps, n_maps, map_of_env, n_envs, C.byref(o) if o is not None else None, C.byref(out))

    assert create(one, 1, None, 0) == 202          # empty batch
    assert create(one, 1, None, -5) == 202
    assert create(one, 0, None, 16) == 202         # no maps
    assert create(None, 1, None, 16) == 202
    assert create(one, 1, None, 16, None) == 202   # no options
    bad = (C.c_int32 * 4)(0, 0, 1, 0)
    assert create(one, 1, bad, 4) == 202 and b"map_of_env" in L.lle_last_error()
    two = (C.c_void_p * 2)(m6, m1)
    assert create(two, 2, (C.c_int32 * 2)(0, 1), 2) == 202 and b"share" in L.lle_last_error()
    o2 = N.VecOptions()
    L.lle_vec_default_options(C.byref(o2))
    o2.reward_dim = 3
    assert create(one, 1, None, 16, o2) == 202
    assert not out.value
    for m in (m1, m6, small):
        L.lle_map_free(m)
    # generator
    g = N.GenOptions()
    L.lle_gen_default_options(C.byref(g))
    h = C.c_void_p()
    assert L.lle_gen_create(C.byref(g), 0, 0, C.byref(h)) == 202      # capacity < 1
    assert L.lle_gen_create(None, 0, 16, C.byref(h)) == 202
    g.width = 33
    assert L.lle_gen_create(C.byref(g), 0, 16, C.byref(h)) == 21      # LLE_LIMIT_EXCEEDED
    g.width, g.n_agents = 5, 33
    assert L.lle_gen_create(C.byref(g), 0, 16, C.byref(h)) == 21
    assert L.lle_gen_run(None, None, 0, 1, 1, 0, None) == 202
    assert L.lle_gen_attempt_seeds(1, -1, None) == 202 and L.lle_gen_attempt_seeds(1, 0, None) == 0
    assert L.lle_gen_cells_to_text(None, 1, 1, None, 0, None) == 202
    cells = (C.c_uint8 * 2)(0, 9)
    assert L.lle_gen_cells_to_text(cells, 1, 2, None, 0, None) == 202 and b"unknown cell code" in L.lle_last_error()


def test_reference_import_paths():
    """`import lle_b200 as lle` keeps the reference's submodule paths (python/lle/__init__.py:211-245) for what is on the path."""
    import lle_b200 as lle
    from lle_b200.exceptions import InvalidActionError, InvalidWorldStateError, ParsingError  # noqa: F401
    from lle_b200.generator import GeneratorBuilder, WorldGenerator, generate  # noqa: F401
    from lle_b200.observations import ObservationType
    from lle_b200.tiles import Direction, Gem, Laser, LaserSource  # noqa: F401
    from lle_b200.types import obs_spec
    from lle_b200.world import World  # noqa: F401

    assert issubclass(InvalidActionError, ValueError)  # pyexceptions.rs
    assert ObservationType.from_str("partial7x7") is ObservationType.PARTIAL_7x7
    assert obs_spec(ObservationType.AGENT0_PERSPECTIVE_LAYERED) == obs_spec("perspective")
    assert {t.value for t in ObservationType} >= {"layered", "flattened", "state", "normalized-state", "layered-padded"}
    for name in ("World", "WorldState", "Action", "EventType", "WorldEvent", "LLE", "ObservationType", "from_file", "from_str", "level",
                 "generate", "GeneratorBuilder", "Agent", "exceptions", "tiles", "observations", "generator", "world"):
        assert hasattr(lle, name), name
