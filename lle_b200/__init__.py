"""lle_b200 — B200-native batched implementation of the `World` step of yamoling/lle.

Public surface (mirrors the reference's `lle` package for the accelerated path):
    World, WorldState, Action, EventType, WorldEvent, LLE        single-world API (N = 1 on the device)
    VecWorld, VecLLE, Map, level, from_str, from_file            batched API
    generate, WorldGenerator, GeneratorBuilder                   lle.generate(...): layouts generated on the device
The package needs its CUDA extension (lle_b200/_native/liblle_b200.so, sm_100a) and a CUDA device;
there is no CPU fallback.
"""
from .types import (Action, Agent, Direction, EventType, Gem, InvalidActionError, InvalidLevelError, InvalidWorldStateError,
                    Laser, LaserSource, ParsingError, WorldEvent, WorldState)
import sys as _sys

from ._native import LIB_PATH, lib as _load_native

if "lle_b200.build" in getattr(_sys, "orig_argv", []) and "-m" in getattr(_sys, "orig_argv", []):
    # `python -m lle_b200.build` imports this package before it runs the builder: build first, then load
    from . import build as _build

    _build.build(force="--force" in _sys.argv, verbose=True)
_load_native()  # fail loudly at import time if the extension is not built

from .vec_world import Map, VecWorld  # noqa: E402
from .world import LLE, Step, World, decode_events  # noqa: E402
from .env import Builder, VecLLE, VecWorldGroup, from_file, from_str, level  # noqa: E402
from .generator import GeneratorBuilder, WorldGenerator, generate  # noqa: E402
from .observations import ObservationType  # noqa: E402
from . import exceptions, generator, observations, tiles, world  # noqa: E402,F401  (the reference's submodule paths)

__all__ = ["Action", "Agent", "Direction", "EventType", "Gem", "InvalidActionError", "InvalidLevelError",
           "InvalidWorldStateError", "Laser", "LaserSource", "ParsingError", "WorldEvent", "WorldState", "Map", "VecWorld",
           "World", "LLE", "Step", "VecLLE", "VecWorldGroup", "Builder", "level", "from_str", "from_file", "decode_events", "LIB_PATH",
           "generate", "WorldGenerator", "GeneratorBuilder", "ObservationType", "exceptions", "tiles", "observations", "generator", "world"]
__version__ = "0.1.0"
