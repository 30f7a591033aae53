"""ctypes wrapper of tests/host_shim (TEST HARNESS: the product's __host__ __device__ logic run on the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "host_shim")
_LIB = os.path.join(_DIR, "_build", "libshim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
        _lib.shim_create.restype = C.c_void_p
    return _lib


def compile_status(text: str):
    buf = C.create_string_buffer(512)
    st = lib().shim_compile(text.encode(), buf, 512)
    return st, buf.value.decode()


class ShimVec:
    def __init__(self, maps, map_of_env, n_envs, *, reward_dim=1, walkable_lasers=True, auto_reset=True, lle_semantics=True,
                 seed=0, env_id_base=0, tile_override_floats=0):
        texts = (C.c_char_p * len(maps))(*[m.encode() for m in maps])
        moe = None if map_of_env is None else (C.c_int * n_envs)(*[int(m) for m in map_of_env])
        err = C.create_string_buffer(512)
        self._h = C.c_void_p(lib().shim_create(texts, len(maps), moe, C.c_long(n_envs), reward_dim, int(walkable_lasers),
                                               int(auto_reset), int(lle_semantics), C.c_uint64(seed), C.c_uint64(env_id_base),
                                               tile_override_floats, err, 512))
        if not self._h:
            raise RuntimeError(err.value.decode())
        d = (C.c_long * 12)()
        lib().shim_dims(self._h, d)
        (self.N, self.A, self.G, self.C, self.H, self.W, self.R, self.S, self.NB, self.obs_stride, self.E, self.n_chunks) = list(d)
        ptrs = (C.c_void_p * 8)()
        lib().shim_buffers(self._h, ptrs)
        N, A = self.N, self.A

        def view(k, ctype, dtype, shape):
            n = int(np.prod(shape))
            arr = np.ctypeslib.as_array(C.cast(ptrs[k], C.POINTER(ctype)), shape=(n,))
            return arr.view(dtype).reshape(shape)

        self._obs_rows = view(0, C.c_float, np.float32, (N, self.obs_stride))
        self.state = view(1, C.c_float, np.float32, (N, self.S))
        self.avail = view(2, C.c_uint8, np.uint8, (N, A, 5))
        self.reward = view(3, C.c_float, np.float32, (N, self.R))
        self.done = view(4, C.c_uint8, np.uint8, (N,))
        self.events = view(5, C.c_uint8, np.uint8, (N, A))
        self.actions = view(6, C.c_int8, np.int8, (N, A))
        self.err = view(7, C.c_uint8, np.uint8, (N,))

    @property
    def obs(self):
        return self._obs_rows[:, : self.C * self.H * self.W].reshape(self.N, self.C, self.H, self.W)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().shim_free(self._h)
            self._h = None

    def reset(self, mask=None):
        ptr = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            ptr = mask.ctypes.data_as(C.POINTER(C.c_uint8))
        lib().shim_reset(self._h, ptr)

    def step(self, actions=None):
        ptr = None
        if actions is not None:
            actions = np.ascontiguousarray(actions, dtype=np.int8)
            ptr = actions.ctypes.data_as(C.POINTER(C.c_int8))
        lib().shim_step(self._h, ptr)

    def set_state(self, pos, gems, alive):
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        gems = np.ascontiguousarray(gems, dtype=np.uint8).reshape(self.N, max(self.G, 0))
        alive = np.ascontiguousarray(alive, dtype=np.uint8)
        gp = gems.ctypes.data_as(C.POINTER(C.c_uint8)) if self.G else None
        lib().shim_set_state(self._h, pos.ctypes.data_as(C.POINTER(C.c_int32)), gp, alive.ctypes.data_as(C.POINTER(C.c_uint8)))

    def set_step_count(self, t):
        lib().shim_set_step_count(self._h, C.c_uint64(t))

    def export_raw(self):
        N, A, NB = self.N, self.A, max(self.NB, 1)
        out = dict(pos=np.zeros((N, A, 2), np.int16), alive=np.zeros((N, A), np.uint8), arrived=np.zeros((N, A), np.uint8),
                   slot=np.zeros((N, A), np.uint8), beam_on=np.zeros((N, NB), np.uint64), collected=np.zeros(N, np.uint64),
                   counters=np.zeros((N, 3), np.uint8))
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        lib().shim_export_raw(self._h, P(out["pos"], C.c_int16), P(out["alive"], C.c_uint8), P(out["arrived"], C.c_uint8),
                              P(out["slot"], C.c_uint8), P(out["beam_on"], C.c_uint64), P(out["collected"], C.c_uint64),
                              P(out["counters"], C.c_uint8))
        return out
