// Small pieces shared by the step kernels (world_kernel.cuh, tiny_kernel.cuh): event / error codes, the fields of a
// per-cell beam entry, the Philox4x32-10 action stream and the action sampler.  __host__ __device__ so that the per-world
// core of the tiny-map kernel (tiny_core.cuh) can also be instantiated by the CPU test-suite (tests/host_shim).
#pragma once
#include <stdint.h>

#include "static_map.h"

#if defined(__CUDACC__)
#define LLE_HD __host__ __device__ __forceinline__
#define LLE_UNROLL _Pragma("unroll")
#else
#define LLE_HD inline
#define LLE_UNROLL
#endif

namespace lle {

// event codes of one agent in one pass (bits 0-1 of the exported event byte)
enum : uint32_t { EV_NONE = 0, EV_EXIT = 1, EV_GEM = 2, EV_DIED = 3 };

// per-env error codes (LLE_ENV_* of include/lle_b200.h)
enum : uint32_t {
    ERR_OK = 0,
    ERR_INVALID_ACTION = 1,      // RuntimeWorldError::InvalidAction (world.rs:444-453)
    ERR_DONE = 2,                // "Cannot step in a done environment" (env.py:166-167)
    ERR_STATE_DUPLICATE = 3,     // world.rs:529-534
    ERR_STATE_OUT_OF_WORLD = 4,  // world.rs:536-540
    ERR_STATE_NOT_WALKABLE = 5,  // world.rs:556-568
    ERR_STATE_MISMATCH = 6,      // world.rs:588-594
};

LLE_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
LLE_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
LLE_HD int ffs32(uint32_t x) {  // 1-based index of the lowest set bit, 0 for x == 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}

// beam-entry fields (static_map.h: LleCellBeams)
LLE_HD int be_b(uint32_t e) { return e & 63u; }
LLE_HD int be_k(uint32_t e) { return (e >> 6) & 63u; }
LLE_HD int be_colour(uint32_t e) { return (e >> 12) & 255u; }
LLE_HD int be_len(uint32_t e) { return (e >> 20) & 127u; }
LLE_HD bool be_enabled(uint32_t e) { return (e >> 27) & 1u; }
LLE_HD bool be_listed(uint32_t e) { return (e >> 28) & 1u; }

LLE_HD uint64_t len_mask(int len) { return len >= 64 ? ~0ull : ((1ull << len) - 1ull); }
LLE_HD uint32_t len_mask32(int len) { return len >= 32 ? ~0u : ((1u << len) - 1u); }

// Action deltas on a packed position (i<<8 | j), src/action.rs:18-26 (N=0, S=1, E=2, W=3, STAY=4)
LLE_HD int act_delta(int a) { return a == 0 ? -256 : a == 1 ? 256 : a == 2 ? 1 : a == 3 ? -1 : 0; }

// ---- Philox4x32-10 action stream -------------------------------------------------------------
LLE_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    LLE_UNROLL
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = umulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = umulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// k-th set bit of a 5-bit availability mask, k = mulhi(word, popcount)
LLE_HD uint32_t pick_action(uint32_t word, uint32_t mask) {
    uint32_t k = umulhi32(word, (uint32_t)popc32(mask & 31u));
    uint32_t m = mask & 31u;
    for (uint32_t n = 0; n < k; ++n) m &= m - 1;  // drop the k lowest set bits
    return m ? (uint32_t)(ffs32(m) - 1) : 4u;
}

}  // namespace lle
