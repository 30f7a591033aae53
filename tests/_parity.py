"""Shared parity checker: compares a batched implementation (device VecWorld or the host shim) with the
oracle, array by array, bit-exactly."""
import numpy as np

FIELDS = ("obs", "state", "avail", "reward", "done", "events", "actions", "err")


def assert_same(impl, ora, raw=None, ctx=""):
    for name in FIELDS:
        a, b = np.asarray(getattr(impl, name)), np.asarray(getattr(ora, name))
        if a.shape != b.shape or not np.array_equal(a, b):
            bad = np.argwhere(a != b) if a.shape == b.shape else None
            first = bad[0] if bad is not None and len(bad) else None
            env = int(first[0]) if first is not None else -1
            raise AssertionError(f"{ctx}: '{name}' differs (shapes {a.shape} vs {b.shape}); first at {first}; "
                                 f"impl={a[tuple(first)] if first is not None else None} oracle={b[tuple(first)] if first is not None else None}; "
                                 f"env {env} actions={np.asarray(ora.actions)[env] if env >= 0 else None}")
    if raw is not None:
        for name in ("pos", "alive", "arrived", "slot", "collected"):
            a, b = raw[name], np.asarray(getattr(ora, name))
            assert np.array_equal(a, b), f"{ctx}: raw '{name}' differs at {np.argwhere(a != b)[:3]}"
        if "extras" in raw:
            assert np.array_equal(raw["extras"], np.asarray(ora.extras)), f"{ctx}: LaserSubgoal extras differ at {np.argwhere(raw['extras'] != np.asarray(ora.extras))[:3]}"
        nb = ora.NB
        if nb:
            assert np.array_equal(raw["beam_on"][:, :nb], np.asarray(ora.beam_on)[:, :nb]), f"{ctx}: raw beam masks differ"
