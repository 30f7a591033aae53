"""The per-world core of the tiny-map step kernel (lle_b200/csrc/tiny_core.cuh, __host__ __device__), instantiated on the
host by tests/host_shim/tiny_host.cpp and driven like the kernel drives it (tickets of 32 worlds, E sub-tiles per emulated warp),
against the oracle: every output of every step, bit for bit.  CPU suite — the same comparisons run on the GPU through the
C ABI in tests/test_gpu_parity.py / test_gpu_fullsize.py, where the kernel itself is the subject."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from _parity import FIELDS
from _util import GOLDEN, build_tiny_host_shim, level_text
from oracle import lle_oracle as lo


_OBS_KINDS = {"layered": (0, 0), "partial3x3": (1, 3), "partial5x5": (1, 5), "partial7x7": (1, 7)}  # LLE_OBS_* of include/lle_b200.h


class TinyHost:
    def __init__(self, maps, map_of_env, n_envs, *, multi_objective=False, walkable_lasers=True, auto_reset=True, lle_semantics=True,
                 seed=0, env_id_base=0, E=8, n_warps=3, obs_type="layered"):
        self.lib = C.CDLL(build_tiny_host_shim())
        self.lib.tiny_host_create.restype = C.c_void_p
        self.lib.tiny_host_buffer.restype = C.c_void_p
        texts = (C.c_char_p * len(maps))(*[lo.prepare_map_text(m) for m in maps])
        moe = None if map_of_env is None else (C.c_int * n_envs)(*[int(m) for m in map_of_env])
        err = C.create_string_buffer(256)
        self.h = C.c_void_p(self.lib.tiny_host_create(texts, len(maps), moe, C.c_long(n_envs), 4 if multi_objective else 1, int(walkable_lasers),
                                                      int(auto_reset), int(lle_semantics), C.c_uint64(seed), C.c_uint64(env_id_base), E, n_warps,
                                                      *_OBS_KINDS[obs_type], err, 256))
        if not self.h:
            raise ValueError(err.value.decode())
        d = (C.c_long * 8)()
        self.lib.tiny_host_dims(self.h, d)
        A, G, H, W, ostr, R, Cc, S = list(d)
        self.n = n_envs

        def view(k, ctype, dtype, shape):
            ptr = C.cast(C.c_void_p(self.lib.tiny_host_buffer(self.h, k)), C.POINTER(ctype))
            return np.ctypeslib.as_array(ptr, shape=(int(np.prod(shape)),)).view(dtype).reshape(shape)

        self._obs_rows = view(0, C.c_float, np.float32, (n_envs, ostr))
        self.obs_shape = (Cc, H, W) if obs_type == "layered" else (A, 2 * A + 3, int(obs_type[-1]), int(obs_type[-1]))
        self.state = view(1, C.c_float, np.float32, (n_envs, S))
        self.avail = view(2, C.c_uint8, np.uint8, (n_envs, A, 5))
        self.reward = view(3, C.c_float, np.float32, (n_envs, R))
        self.done = view(4, C.c_uint8, np.uint8, (n_envs,))
        self.events = view(5, C.c_uint8, np.uint8, (n_envs, A))
        self.actions = view(6, C.c_int8, np.int8, (n_envs, A))
        self.err = view(7, C.c_uint8, np.uint8, (n_envs,))

    @property
    def obs(self):
        return self._obs_rows[:, : int(np.prod(self.obs_shape))].reshape((self.n,) + tuple(self.obs_shape))

    def step(self, actions=None):
        ptr = None
        if actions is not None:
            actions = np.ascontiguousarray(actions, dtype=np.int8)
            ptr = actions.ctypes.data_as(C.c_void_p)
        self.lib.tiny_host_step(self.h, ptr)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.tiny_host_free(self.h)
            self.h = None


def compare(tiny, ora, ctx, fields=FIELDS):
    for name in fields:
        a, b = np.asarray(getattr(tiny, name)), np.asarray(getattr(ora, name))
        assert a.shape == b.shape, f"{ctx}: '{name}' shapes {a.shape} vs {b.shape}"
        if not np.array_equal(a, b):
            first = np.argwhere(a != b)[0]
            raise AssertionError(f"{ctx}: '{name}' differs first at {first}: core={a[tuple(first)]} oracle={b[tuple(first)]}")


def run_pair(maps, map_of_env, n, steps, *, E=8, n_warps=3, **kw):
    okw = dict(multi_objective=kw.get("multi_objective", False), walkable_lasers=kw.get("walkable_lasers", True),
               auto_reset=kw.get("auto_reset", True), seed=kw.get("seed", 0), env_id_base=kw.get("env_id_base", 0))
    ora = lo.OracleVec(maps, map_of_env, n, obs_type=kw.get("obs_type", "layered"), **okw)
    tiny = TinyHost(maps, map_of_env, n, E=E, n_warps=n_warps, obs_type=kw.get("obs_type", "layered"), **okw)
    # after reset: state / availability come from the core's own reset; observations appear with the first step
    for t in range(steps):
        ora.step(None)
        tiny.step(None)
        compare(tiny, ora, f"step {t}")
    return tiny, ora


def eligible(text):
    """Maps the tiny path takes: <= 4 agents, a record of <= 8 words, beams <= 32 cells."""
    try:
        TinyHost([text], None, 1)
        return True
    except ValueError:
        return False


def test_levels():
    for level in range(1, 7):
        run_pair([level_text(level)], None, 100, 150, seed=level)


def test_layout_corpus(layouts):
    n = 0
    for k, (name, text) in enumerate(sorted(layouts.items())):
        if not eligible(text):
            continue
        run_pair([text], None, 70, 120, seed=100 + k, E=(4, 8, 16, 32)[k % 4], n_warps=1 + k % 3)
        n += 1
    assert n >= 20


def test_generated_5x5_maps_heterogeneous_batch():
    """BASELINE configs[2] shape: many distinct 5x5 maps in one batch, changing every few worlds (sub-tiles change map)."""
    with open(os.path.join(GOLDEN, "generated_5x5.json")) as f:
        maps = json.load(f)["maps"][:96]
    moe = [m for m in range(96) for _ in range(5)]
    run_pair(maps, moe, len(moe), 60, seed=21, E=8, n_warps=2)
    moe2 = [(e * 37) % 96 for e in range(500)]  # every world another map: every sub-tile is rebuilt
    run_pair(maps, moe2, 500, 40, seed=22, E=16, n_warps=1)


def test_options():
    run_pair([level_text(6)], None, 96, 150, multi_objective=True, seed=7)
    run_pair([level_text(5)], None, 96, 150, walkable_lasers=False, seed=8)
    run_pair([level_text(4)], None, 96, 100, auto_reset=False, seed=9)
    run_pair([level_text(3)], None, 33, 100, env_id_base=12345, seed=10, E=32)


def test_supplied_actions_with_invalid_ones():
    text = level_text(6)
    n = 200
    ora = lo.OracleVec([text], None, n, seed=3)
    tiny = TinyHost([text], None, n, seed=3)
    rng = np.random.default_rng(5)
    for t in range(120):
        acts = rng.integers(0, 5, size=(n, 4)).astype(np.int8)  # many are unavailable: those worlds must stay untouched
        if t % 7 == 0:
            acts[::13, 1] = 9
        ora.step(acts)
        tiny.step(acts)
        compare(tiny, ora, f"step {t}")
    assert int(np.asarray(ora.err).sum()) > 0


@pytest.mark.parametrize("obs_type,E", [("partial3x3", 4), ("partial5x5", 2), ("partial7x7", 1)])
def test_partial_observations(obs_type, E, layouts):
    """PartialGenerator (observations.py:312-369) through the per-world core: windows at the border, sources (-1), lit lasers of
    every colour, gems, exits and the other agents, on the levels and on the layout corpus."""
    for level in (1, 3, 6):
        run_pair([level_text(level)], None, 70, 100, seed=30 + level, E=E, obs_type=obs_type)
    n = 0
    for k, (name, text) in enumerate(sorted(layouts.items())):
        if not eligible(text) or k % 3:
            continue
        run_pair([text], None, 40, 60, seed=300 + k, E=E, n_warps=2, obs_type=obs_type)
        n += 1
    assert n >= 6
