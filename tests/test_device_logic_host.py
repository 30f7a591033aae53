"""CPU-side differential tests of the product's __host__ __device__ logic (step_core.cuh, vec_kernels.cuh,
map_compiler.cpp) against the oracle, via tests/host_shim.  These do not replace the GPU parity tests
(tests/test_gpu_parity.py): they let the bitmask formulation, the Philox action rule, the state packing and
the un-patch/patch observation renderer be validated bit-exactly where no GPU is available."""
import numpy as np
import pytest

from oracle import lle_oracle as lo
from _parity import assert_same
from _shim import ShimVec, compile_status
from _util import level_text


def run_pair(maps, map_of_env, n_envs, steps, **kw):
    okw = dict(multi_objective=kw.get("reward_dim", 1) == 4, walkable_lasers=kw.get("walkable_lasers", True),
               auto_reset=kw.get("auto_reset", True), seed=kw.get("seed", 0), env_id_base=kw.get("env_id_base", 0))
    ora = lo.OracleVec(maps, map_of_env, n_envs, **okw)
    shim = ShimVec(maps, map_of_env, n_envs, **kw)
    assert_same(shim, ora, shim.export_raw(), "after reset")
    deaths = 0
    for t in range(steps):
        ora.step(None, n_threads=1)
        shim.step(None)
        assert_same(shim, ora, shim.export_raw(), f"step {t}")
        deaths += int((np.asarray(ora.events) >= 3).sum())
    return deaths


def test_levels_philox_rollout():
    for level in range(1, 7):
        run_pair([level_text(level)], None, 96, 150, seed=level)


def test_layout_corpus_philox_rollout(layouts):
    total_deaths = 0
    for k, (name, text) in enumerate(sorted(layouts.items())):
        total_deaths += run_pair([text], None, 64, 120, seed=100 + k)
    assert total_deaths > 50  # the corpus exercises death cascades


def test_no_auto_reset_done_envs_refuse_to_step():
    ora = lo.OracleVec([level_text(6)], None, 64, auto_reset=False, seed=5)
    shim = ShimVec([level_text(6)], None, 64, auto_reset=False, seed=5)
    for t in range(200):
        ora.step(None, 1)
        shim.step(None)
        assert_same(shim, ora, shim.export_raw(), f"step {t}")
    assert (np.asarray(ora.err) == 2).any()  # done envs report LLE_ENV_DONE


def test_multi_objective_and_unwalkable_lasers():
    for level in (4, 5, 6):
        run_pair([level_text(level)], None, 64, 120, reward_dim=4, seed=7)
        run_pair([level_text(level)], None, 64, 120, walkable_lasers=False, seed=8)


def test_heterogeneous_maps_in_one_batch():
    maps = [level_text(2), level_text(3), level_text(4)]  # same (H, W, A, G), 0 / 1 / 2 beams
    moe = [(e * 7 + e // 5) % 3 for e in range(160)]
    run_pair(maps, moe, 160, 150, seed=11, env_id_base=1000)


def test_chunked_tiles():
    # force the chunked streaming path used by maps whose observation exceeds the tile budget
    for floats in (96, 500, 1000):
        ora = lo.OracleVec([level_text(6)], None, 40, seed=3)
        shim = ShimVec([level_text(6)], None, 40, seed=3, tile_override_floats=floats)
        assert shim.n_chunks > 1
        for t in range(60):
            ora.step(None, 1)
            shim.step(None)
            assert_same(shim, ora, None, f"chunk {floats} step {t}")


def test_supplied_actions_with_invalid_ones():
    rng = np.random.default_rng(0)
    text = level_text(5)
    ora = lo.OracleVec([text], None, 128, seed=1)
    shim = ShimVec([text], None, 128, seed=1)
    n_bad = 0
    for t in range(120):
        actions = rng.integers(0, 5, size=(128, ora.A)).astype(np.int8)  # many are unavailable
        if t % 10 == 0:
            actions[::17, 0] = 7  # not an Action at all
        ora.step(actions, 1)
        shim.step(actions)
        assert_same(shim, ora, shim.export_raw(), f"step {t}")
        n_bad += int((np.asarray(ora.err) == 1).sum())
    assert n_bad > 100


def test_many_agents_bucket_16():
    rows = [" .   .   . . . ."] + [f"S{k}  L{k}W  . . . X" for k in range(14)]
    run_pair(["\n".join(rows)], None, 32, 60, seed=2)


def test_set_state_fuzz(layouts):
    rng = np.random.default_rng(42)
    texts = [level_text(6), level_text(5), layouts["eight-agent-interdependent-8"]] + [t for _, t in sorted(layouts.items())][:10]
    seen = set()
    for text in texts:
        n = 96
        ora = lo.OracleVec([text], None, n, auto_reset=False, seed=9)
        shim = ShimVec([text], None, n, auto_reset=False, seed=9)
        world = lo.World(text)
        H, W, A, G = world.height, world.width, world.n_agents, world.n_gems
        for rnd in range(6):
            # a few random steps so that the "current state" being replaced varies (from a fresh reset: a
            # world that failed set_state with InvalidWorldState is inconsistent in the reference)
            ora.reset()
            shim.reset()
            for _ in range(3 + rnd):
                ora.step(None, 1)
                shim.step(None)
            pos = np.stack([rng.integers(-1, H + 1, size=(n, A)), rng.integers(-1, W + 1, size=(n, A))], axis=-1).astype(np.int32)
            pos[: n // 2] = np.clip(pos[: n // 2], 0, [H - 1, W - 1])
            gems = rng.integers(0, 2, size=(n, G)).astype(np.uint8)
            alive = (rng.random((n, A)) < 0.85).astype(np.uint8)
            shim.set_state(pos, gems, alive)
            # oracle: env by env through LLE.set_state semantics, mapping exceptions to the device error byte
            exp_err = np.zeros(n, np.uint8)
            for e in range(n):
                env = _oracle_env(ora, e)
                exp_err[e] = env(pos[e], gems[e], alive[e])
            ora_refresh(ora)
            seen.update(exp_err.tolist())
            assert np.array_equal(np.asarray(shim.err), exp_err), f"set_state err differs: {np.argwhere(np.asarray(shim.err) != exp_err)[:4]}"
            raw = shim.export_raw()
            for name in ("pos", "alive", "arrived", "slot", "collected"):
                assert np.array_equal(raw[name], np.asarray(getattr(ora, name))), name
            if ora.NB:
                assert np.array_equal(raw["beam_on"][:, : ora.NB], np.asarray(ora.beam_on)[:, : ora.NB])
            for name in ("obs", "state", "done"):
                assert np.array_equal(np.asarray(getattr(shim, name)), np.asarray(getattr(ora, name))), name
            # After InvalidWorldState (world.rs:588-594) the reference returns before refreshing its
            # availability cache (world.rs:595), which is then stale w.r.t. the modified world; the device
            # always derives availability from the state (documented deviation on this error path).
            ok = exp_err != 6
            assert np.array_equal(np.asarray(shim.avail)[ok], np.asarray(ora.avail)[ok]), "avail"
    assert {0, 3, 4, 5, 6} <= seen


# ---- helpers driving the oracle's per-env set_state through its C API -------------------------------
import ctypes as C  # noqa: E402


def _oracle_env(ora, e):
    L = lo.lib()

    def call(pos, gems, alive):
        na, ng = len(pos), len(gems)
        p = (C.c_long * (2 * na))(*[int(x) for x in np.asarray(pos).reshape(-1)])
        g = (C.c_uint8 * max(1, ng))(*[int(x) for x in gems])
        a = (C.c_uint8 * na)(*[int(x) for x in alive])
        st = L.lleo_vec_set_state_env(ora._h, C.c_long(e), p, na, g, ng, a)
        return {0: 0, 107: None, 105: 4, 104: 5}.get(st, st) if st != 107 else (3 if b"same position" in L.lleo_last_error() else 6)

    return call


def ora_refresh(ora):
    lo.lib().lleo_vec_refresh(ora._h)


def test_map_compiler_errors_match_oracle():
    cases = ["", "S0 S0 X X", "S1 S0 X", ". . G", "X S0 .\n . .", "S0 Q X", "S0  S1 X . X\nL1N .  . . .", "Sx X", "S0 L0 X", "S0 LxE X"]
    for text in cases:
        st, msg = compile_status(text)
        try:
            lo.World(text)
            expected = None
        except lo.ParsingError as ex:
            expected = str(ex)
        except lo.RustPanic:
            expected = "InvalidDirection"
        assert (st != 0) == (expected is not None), (text, st, msg, expected)
        if expected:
            assert msg.split(" ")[0].split("{")[0].strip() in expected or expected.split(":")[0] in msg, (text, msg, expected)


def test_synthetic_64x64_config5():
    from _util import synthetic_map

    text = synthetic_map(64, 64, 8, 16, seed=5)
    ora = lo.OracleVec([text], None, 32, seed=6)
    shim = ShimVec([text], None, 32, seed=6)
    assert shim.n_chunks > 1
    for t in range(25):
        ora.step(None, 2)
        shim.step(None)
        assert_same(shim, ora, shim.export_raw(), f"step {t}")
