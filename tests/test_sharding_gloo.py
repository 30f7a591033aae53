"""world_size-2 `gloo` test (CPU) of the multi-GPU host logic: contiguous env sharding, global env ids for the
action stream (results independent of the number of ranks) and the end-of-run stats reduction.  The engine in this
test is the oracle; on GPUs the same code drives VecWorld (bench.py)."""
import importlib.util
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import level_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# lle_b200/__init__ needs the CUDA extension; sharding.py is plain host logic and is loaded on its own here
_spec = importlib.util.spec_from_file_location("lle_b200_sharding", os.path.join(ROOT, "lle_b200", "sharding.py"))
sharding = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sharding)

N_TOTAL, STEPS, SEED = 150, 40, 17


def test_shard_range_partitions():
    for n, w in ((150, 2), (7, 3), (65536, 8), (5, 8)):
        ranges = [sharding.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world_size, port, out_dir):
    import sys

    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from oracle import lle_oracle as lo

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    begin, end = sharding.shard_range(N_TOTAL, rank, world_size)
    vec = lo.OracleVec([level_text(6)], None, end - begin, seed=SEED, env_id_base=begin)
    episodes = 0
    for _ in range(STEPS):
        vec.step(None, n_threads=1)
        episodes += int(vec.done.sum())
    stats = sharding.reduce_stats(torch.tensor([episodes, end - begin], dtype=torch.int64))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), obs=vec.obs, state=vec.state, actions=vec.actions, stats=stats.numpy(),
             episodes=episodes)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_one(tmp_path):
    from oracle import lle_oracle as lo

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    single = lo.OracleVec([level_text(6)], None, N_TOTAL, seed=SEED)
    episodes = 0
    for _ in range(STEPS):
        single.step(None, n_threads=2)
        episodes += int(single.done.sum())
    for name in ("obs", "state", "actions"):
        assert np.array_equal(np.concatenate([p[name] for p in parts]), getattr(single, name)), name
    for p in parts:  # every rank holds the reduced totals
        assert p["stats"].tolist() == [episodes, N_TOTAL]
    assert sum(int(p["episodes"]) for p in parts) == episodes
