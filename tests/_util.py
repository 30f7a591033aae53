import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def level_text(n: int) -> str:
    with open(os.path.join(GOLDEN, "levels", f"lvl{n}")) as f:
        return f.read()


def synthetic_map(h: int, w: int, n_agents: int, n_sources: int, seed: int = 0, n_gems: int | None = None) -> str:
    """The seeded synthetic map of BASELINE config 5 (lle_b200/workloads.py: a text generator, no device work)."""
    from lle_b200.workloads import synthetic_map as make

    return make(h, w, n_agents, n_sources, seed, n_gems)


def rollout_digest(x) -> list[int]:
    """Order-sensitive 64-bit checksums of one step's outputs (obs, state, avail, reward, done, events, actions)."""
    import zlib

    import numpy as np

    out = []
    for name in ("obs", "state", "avail", "reward", "done", "events", "actions"):
        a = np.ascontiguousarray(np.asarray(getattr(x, name)))
        out.append(zlib.crc32(a.tobytes()) ^ (zlib.adler32(a.tobytes()) << 32))
    return out


def random_map(rng, allow_invalid: bool = False) -> str:
    """A random v1 map for differential fuzzing: random size, agents, exits, walls, gems, voids and laser sources of any
    colour (including colours >= n_agents) and direction.  With allow_invalid, some maps break a parser rule on purpose
    (too few exits, a duplicated start, a start lost to a foreign beam ...), so that error paths are compared as well."""
    h, w = rng.randint(1, 7), rng.randint(2, 8)
    n_agents = rng.randint(1, min(5, h * w // 2))
    cells = [(i, j) for i in range(h) for j in range(w)]
    rng.shuffle(cells)
    grid = [["." for _ in range(w)] for _ in range(h)]
    take = iter(cells)
    try:
        for a in range(n_agents):
            i, j = next(take)
            grid[i][j] = f"S{a}"
        n_exits = n_agents + rng.randint(0, 2)
        if allow_invalid and rng.random() < 0.05:
            n_exits = max(0, n_agents - 1)
        for _ in range(n_exits):
            i, j = next(take)
            grid[i][j] = "X"
        for _ in range(rng.randint(0, 3)):
            i, j = next(take)
            grid[i][j] = f"L{rng.randint(0, n_agents + (1 if rng.random() < 0.2 else -1 if n_agents > 1 else 0))}{rng.choice('NESW')}"
        for (i, j) in take:
            r = rng.random()
            grid[i][j] = "@" if r < 0.12 else "G" if r < 0.22 else "V" if r < 0.26 else "."
    except StopIteration:
        pass
    if allow_invalid and rng.random() < 0.03:
        i, j = rng.choice(cells)
        grid[i][j] = "S0"
    return "\n".join(" ".join(row) for row in grid)


def build_gen_host_shim() -> str:
    """g++ build of tests/host_shim/gen_host.cpp: the generator core of the CUDA kernel (lle_b200/csrc/gen_core.cuh,
    __host__ __device__) instantiated on the host for the CPU test-suite.  Test infrastructure only."""
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "host_shim", "gen_host.cpp")
    out_dir = os.path.join(here, "host_shim", "_build")
    out = os.path.join(out_dir, "libgen_host.so")
    csrc = os.path.join(os.path.dirname(here), "lle_b200", "csrc")
    deps = [src, os.path.join(csrc, "gen_core.cuh"), os.path.join(csrc, "gen_config.hpp"), os.path.join(os.path.dirname(here), "include", "lle_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", out], check=True)
    return out


def build_tiny_host_shim() -> str:
    """g++ build of tests/host_shim/tiny_host.cpp: the per-world core of the tiny-map step kernel (lle_b200/csrc/tiny_core.cuh,
    __host__ __device__) plus the product's host map compiler, for the CPU test-suite.  Test infrastructure only."""
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), "lle_b200", "csrc")
    srcs = [os.path.join(here, "host_shim", "tiny_host.cpp"), os.path.join(csrc, "map_compiler.cpp"), os.path.join(csrc, "toml_config.cpp")]
    out_dir = os.path.join(here, "host_shim", "_build")
    out = os.path.join(out_dir, "libtiny_host.so")
    deps = srcs + [os.path.join(csrc, f) for f in ("tiny_core.cuh", "step_common.cuh", "static_map.h", "map_compiler.hpp", "toml_lite.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++"] + srcs + ["-o", out], check=True)
    return out
