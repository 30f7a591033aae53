"""Parity at the FULL sizes of BASELINE.json, on the workloads exactly as bench.py builds them (lle_b200/workloads.py).

An env's stream depends only on its global env id, the seed and the step count, so windows of a multi-million-env batch
can be replayed by small oracle batches (env_id_base = global id of the window's first env) and compared bit for bit,
every output, at many points of a long rollout.  Also: two GPUs against one (skipped with fewer than two devices).
"""
import numpy as np
import pytest
import torch

from oracle import lle_oracle as lo

pytestmark = pytest.mark.gpu

FIELDS = ("obs", "state", "avail", "reward", "done", "events", "actions", "err")
SEED = 2026


class Window:
    """Envs [s0, s0 + w) of one VecWorld, mirrored by an oracle batch over the same global env ids."""

    def __init__(self, part, texts, map_of_env, s0, w, seed):
        self.part, self.s0, self.w = part, s0, w
        base = int(part_base(part)) + s0
        if map_of_env is None:
            self.ora = lo.OracleVec([texts[0]], None, w, seed=seed, env_id_base=base, auto_reset=True)
        else:
            ids = sorted({int(m) for m in map_of_env[s0:s0 + w]})
            local = [ids.index(int(m)) for m in map_of_env[s0:s0 + w]]
            self.ora = lo.OracleVec([texts[i] for i in ids], local, w, seed=seed, env_id_base=base, auto_reset=True)

    def step(self):
        self.ora.step(None)

    def check(self, ctx):
        for name in FIELDS:
            a = getattr(self.part, name)[self.s0:self.s0 + self.w].cpu().numpy()
            b = np.asarray(getattr(self.ora, name))
            assert a.shape == b.shape and np.array_equal(a, b), f"{ctx}: window at env {self.s0}: '{name}' differs"


_BASES = {}


def part_base(part):
    return _BASES[id(part)]


def build(cfg, n=None, base=0, device=0):
    from lle_b200 import workloads

    wl = workloads.build(cfg, n, device=device, seed=SEED, env_id_base=base)
    b = base
    for p in wl.parts:
        _BASES[id(p)] = b
        b += p.n_envs
    return wl


def run_windows(wl, windows, steps, every):
    for w in windows:
        w.check("after reset")
    for t in range(steps):
        wl.step()
        for w in windows:
            w.step()
        if t % every == every - 1 or t == steps - 1:
            for p in wl.parts:
                p.synchronize()
            for w in windows:
                w.check(f"step {t}")
    for p in wl.parts:
        assert int(p.err.sum()) == 0


def spread(n, w, k, rng):
    return [0, n - w] + [int(x) for x in rng.integers(0, n - w, size=k)]


def test_config2_full_size_2000_steps():
    """BASELINE configs[1]: level 6 x 65,536, 2,048 steps, eight windows of 96 envs compared every 64 steps."""
    wl = build(2)
    part = wl.parts[0]
    rng = np.random.default_rng(1)
    windows = [Window(part, wl.texts, None, s0, 96, SEED) for s0 in spread(part.n_envs, 96, 6, rng)]
    run_windows(wl, windows, 2048, 64)


@pytest.mark.parametrize("obs_type", ["partial3x3", "partial5x5"])
def test_config2_full_size_partial_observations(obs_type):
    """Level 6 x 65,536 with PartialGenerator observations (thread-per-world kernel): windows replayed by the oracle."""
    import lle_b200
    from lle_b200.workloads import level_text

    n = 65536
    vec = lle_b200.VecWorld(level_text(6), n, seed=SEED, obs_type=obs_type)
    _BASES[id(vec)] = 0
    rng = np.random.default_rng(11)
    windows = []
    for s0 in spread(n, 48, 6, rng):
        w = Window.__new__(Window)
        w.part, w.s0, w.w = vec, s0, 48
        w.ora = lo.OracleVec([level_text(6)], None, 48, seed=SEED, env_id_base=s0, auto_reset=True, obs_type=obs_type)
        windows.append(w)
    for w in windows:
        w.check("after reset")
    for t in range(300):
        vec.step(None)
        for w in windows:
            w.step()
        if t % 60 == 59:
            vec.synchronize()
            for w in windows:
                w.check(f"step {t}")
    assert int(vec.err.sum()) == 0


def test_config2_full_size_closed_loop_over_parts():
    """Level 6 x 65,536 through lle_vec_parts_* (the loop bench.py times as e2e): the host feeds recorded actions part by part; the
    reward / done it receives for six windows of envs are compared with the oracle at every step, every output at the end."""
    import lle_b200
    from lle_b200.workloads import level_text

    n, steps, P = 65536, 120, 8
    text = level_text(6)
    src = lle_b200.VecWorld(text, n, seed=SEED)
    rec = torch.empty((steps, n, src.n_agents), dtype=torch.int8).pin_memory()
    for t in range(steps):
        src.step(None)
        rec[t].copy_(src.actions, non_blocking=True)
    src.synchronize()
    del src
    rec_np = rec.numpy()
    vec = lle_b200.VecWorld(text, n, seed=SEED)
    _BASES[id(vec)] = 0
    rng = np.random.default_rng(12)
    windows = []
    for s0 in spread(n, 40, 4, rng):
        w = Window.__new__(Window)
        w.part, w.s0, w.w = vec, s0, 40
        w.ora = lo.OracleVec([text], None, 40, seed=SEED, env_id_base=s0, auto_reset=True)
        windows.append(w)
    act = torch.empty((n, vec.n_agents), dtype=torch.int8).pin_memory()
    rew = torch.empty((n, vec.reward_dim), dtype=torch.float32).pin_memory()
    done = torch.empty((n,), dtype=torch.uint8).pin_memory()
    act_np, rew_np, done_np = act.numpy(), rew.numpy(), done.numpy()
    with vec.parts_loop(P, act, rew, done) as loop:
        slices = [loop.slice(k) for k in range(loop.n_parts)]
        loop.launch()
        for k, sl in enumerate(slices):
            act_np[sl] = rec_np[0][sl]
            loop.feed(k)
        loop.launch()
        for t in range(steps):
            for w in windows:
                w.ora.step(rec_np[t][w.s0:w.s0 + w.w])
            for k, sl in enumerate(slices):
                loop.wait(k)
                for w in windows:
                    if sl.start <= w.s0 and w.s0 + w.w <= sl.stop:  # the window lies in this part
                        assert np.array_equal(rew_np[w.s0:w.s0 + w.w], np.asarray(w.ora.reward)), f"reward, window {w.s0}, step {t}"
                        assert np.array_equal(done_np[w.s0:w.s0 + w.w], np.asarray(w.ora.done)), f"done, window {w.s0}, step {t}"
                if t + 1 < steps:
                    act_np[sl] = rec_np[t + 1][sl]
                    loop.feed(k)
            if t + 2 < steps:
                loop.launch()
    vec.synchronize()
    for w in windows:
        w.check("after the loop")
    assert int(vec.err.sum()) == 0


def test_config1_full_size():
    wl = build(1)
    part = wl.parts[0]
    rng = np.random.default_rng(2)
    windows = [Window(part, wl.texts, None, s0, 128, SEED) for s0 in spread(part.n_envs, 128, 4, rng)]
    run_windows(wl, windows, 600, 50)


def test_config3_full_size_device_generated_maps():
    """BASELINE configs[2]: 1,024 maps from the device generator x 1,024 envs each; windows straddle map boundaries."""
    wl = build(3)
    part = wl.parts[0]
    assert len(set(wl.texts)) == 1024 and part.n_envs == 1 << 20
    rng = np.random.default_rng(3)
    starts = spread(part.n_envs, 256, 4, rng) + [1024 * 7 - 100, 1024 * 512 - 128]
    windows = [Window(part, wl.texts, wl.map_of_env, s0, 256, SEED) for s0 in starts]
    run_windows(wl, windows, 60, 6)


def test_config4_full_size_mixed_levels():
    """BASELINE configs[3] per-GPU slice: 2,097,152 envs, the six levels mixed (one sub-batch per level), auto-reset."""
    wl = build(4)
    assert wl.n_envs == 1 << 21 and len(wl.parts) == 6
    rng = np.random.default_rng(4)
    windows = []
    for part, text in zip(wl.parts, wl.texts):
        for s0 in [0, part.n_envs - 64, int(rng.integers(0, part.n_envs - 64))]:
            windows.append(Window(part, [text], None, s0, 64, SEED))
    run_windows(wl, windows, 48, 8)


def test_config5_full_size_64x64():
    """BASELINE configs[4] at its per-GPU size: 262,144 envs of the 64x64 / 8 agents / 16 sources map (86 GB of observations)."""
    free_b, _ = torch.cuda.mem_get_info(0)
    n = 1 << 18
    if n * (4 * 20 * 64 * 64 + 4096) > 0.92 * free_b:
        pytest.skip("needs ~90 GB of free device memory")
    wl = build(5, n)
    part = wl.parts[0]
    windows = [Window(part, wl.texts, None, s0, 6, SEED) for s0 in (0, 131072 - 3, n - 6)]
    run_windows(wl, windows, 10, 5)
    obs = part.obs
    for lo_, hi_ in ((0, 4096), (n - 4096, n)):
        assert torch.equal(obs[lo_:hi_, :8].sum(dim=(2, 3)), torch.ones(hi_ - lo_, 8, device=obs.device))  # one cell per agent plane


def test_two_gpus_equal_one():
    """Range sharding over two real devices: the concatenated outputs of two ranks equal the single-device batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    n = 8192
    for cfg in (2, 3):
        n_cfg = n if cfg == 2 else 1 << 14
        one = build(cfg, n_cfg, base=0, device=0)
        two = [build(cfg, n_cfg // 2, base=r * (n_cfg // 2), device=r) for r in range(2)]
        if cfg == 3:  # the map of an env depends on its global id only
            assert np.array_equal(one.map_of_env, np.concatenate([t.map_of_env for t in two]))
        for t in range(40):
            one.step()
            for t2 in two:
                t2.step()
        for name in FIELDS:
            a = getattr(one.parts[0], name).cpu()
            b = torch.cat([getattr(t2.parts[0], name).cpu() for t2 in two])
            assert torch.equal(a, b), f"config {cfg}: '{name}' differs between one and two devices"
