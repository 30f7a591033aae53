// The fused step kernel of lle_b200 and the __host__ __device__ building blocks it is made of.
//
// One kernel launch advances N independent worlds by one joint action (or resets them, or forces a
// state) and writes every per-step output of the reference's `LLE.step`:
//   world.rs:435-475 (step) . world.rs:343-363 (availability) . reward_strategy.py:58-109 .
//   env.py:253-254 (done) . pyworld_state.rs:79-101 (state vector) . observations.py:254-266 (layered).
//
// Execution model:
//   * work unit = 32 consecutive environments, fetched dynamically by a warp (atomic ticket), so SMs
//     self-balance and the grid is sized from the SM count, not from N;
//   * phase 1, one THREAD per environment (`unit_logic`): the record (a handful of 32-bit words,
//     word-major in HBM => one coalesced 128 B transaction per word per warp) is unpacked into
//     registers, the transition of step_core.cuh runs in registers, the record and the small outputs
//     are written back;
//   * phase 2, one WARP per environment: the layered observation is not recomputed cell by cell.
//     Each warp keeps two observation tiles in shared memory, initialised from the map's static plane;
//     for the next environment it un-patches the few cells that depended on the previous occupant's
//     state (`tile_unpatch`), patches the new ones (`tile_patch`: agents, lit laser cells, uncollected
//     gems) and hands the tile to the TMA engine with ONE `cp.async.bulk.global.shared::cta` (SASS
//     UBLKCP) per tile.  The LSU never touches the 7.5 KB of an observation; the two tiles double-buffer
//     patching against the drain.
//
// The HD functions are also compiled for the host by tests/host_shim (a test-only harness that checks
// this logic against the oracle on CPU); the product has no CPU path.
#pragma once
#include "step_core.cuh"

namespace lle {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

enum Mode : int { MODE_STEP = 0, MODE_RESET = 1, MODE_SET_STATE = 2 };

struct KParams {
    const uint8_t* const* blobs;  // device array [n_maps] of map blobs
    const int32_t* map_of_env;    // device [N_pad] or nullptr
    uint32_t* words;              // [n_words][N_pad]
    LleStateLayout L;
    int64_t N, N_pad;
    int32_t A, G, NBmax, C, H, W, S, R, HW;
    float* obs;
    int64_t obs_stride;  // floats per env
    float* state;
    uint8_t* avail;
    float* reward;
    uint8_t* done;
    uint8_t* events;
    int8_t* actions;
    uint8_t* err;
    const int8_t* actions_in;
    const uint8_t* reset_mask;
    const int32_t* ss_pos;
    const uint8_t* ss_gems;
    const uint8_t* ss_alive;
    uint64_t seed, env_id_base, t;
    int32_t mode, auto_reset, lle_semantics, walkable, write_obs;
    // observation tiling
    int32_t E;             // environments per tile (n_chunks == 1)
    int32_t n_chunks;      // > 1: one environment's block is streamed in chunks (E == 1)
    int32_t chunk_floats;  // floats per chunk, multiple of 4
    int32_t tile_floats;   // floats per smem buffer
    int32_t warp_smem_bytes;
    uint32_t* sched;       // [0] next unit ticket, [1] warps finished
    uint32_t n_units, n_warps_total;
};

// Observation descriptor of one environment: the only state the layered tensor depends on.
//   words [0, PW)      packed positions (two per word)
//   words [PW, PW+2)   collected mask
//   words [PW+2 + 2b]  on-mask of beam b (lo, hi)
template <int AMAX, int NBMAX>
struct Desc {
    static constexpr int PW = (AMAX + 1) / 2;
    static constexpr int WORDS = PW + 2 + 2 * NBMAX;
};

LLE_HD bool desc_bit(const uint32_t* d, int stride, int PW, const LlePatch& pe) {
    if (pe.src == 0xFF) {  // gem: lit while NOT collected (observations.py:260-263)
        uint32_t w = d[(PW + (pe.bit >> 5)) * stride];
        return !((w >> (pe.bit & 31)) & 1u);
    }
    uint32_t w = d[(PW + 2 + 2 * pe.src + (pe.bit >> 5)) * stride];  // laser: lit while its beam bit is on (:256-259)
    return (w >> (pe.bit & 31)) & 1u;
}
LLE_HD uint32_t desc_pos(const uint32_t* d, int stride, int a) {
    uint32_t w = d[(a >> 1) * stride];
    return (a & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// Writes env's descriptor, word w at out[w*stride].
template <int AMAX, int NBMAX>
LLE_HD void desc_write(const Env<AMAX, NBMAX>& e, uint32_t* out, int stride) {
    using D = Desc<AMAX, NBMAX>;
#pragma unroll
    for (int w = 0; w < D::PW; ++w) {
        uint32_t v = e.pos[2 * w];
        if (2 * w + 1 < AMAX) v |= (uint32_t)e.pos[2 * w + 1] << 16;
        out[w * stride] = v;
    }
    out[(D::PW + 0) * stride] = (uint32_t)e.collected;
    out[(D::PW + 1) * stride] = (uint32_t)(e.collected >> 32);
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int b = 0; b < NBMAX; ++b) {
        out[(D::PW + 2 + 2 * b) * stride] = (uint32_t)e.on[b];
        out[(D::PW + 3 + 2 * b) * stride] = (uint32_t)(e.on[b] >> 32);
    }
}

// ---- phase 1: the whole transition of one environment ------------------------------------------------
template <int AMAX, int NBMAX>
LLE_HD void unit_logic(const KParams& p, int64_t env, const MapView& mv, Env<AMAX, NBMAX>& e) {
    const int A = p.A, NB = mv.hdr->NB;
    env_load(e, p.L, A, NB, [&](int w) { return p.words[(int64_t)w * p.N_pad + env]; });

    uint8_t ev[AMAX];
    uint8_t act[AMAX];
    float reward[4] = {0.f, 0.f, 0.f, 0.f};
    uint8_t err = ERR_OK;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) {
        ev[a] = 0;
        act[a] = 4;
    }
    bool touch_transition = true;  // whether reward/done/events/err/actions are (re)written

    if (p.mode == MODE_STEP) {
        uint32_t av[AMAX];
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a) av[a] = a < A ? env_available(mv, e, a) : 16u;
        if (p.actions_in) {
            if (env < p.N) {  // padding worlds (env >= N) just STAY
#pragma unroll(AMAX <= 8 ? 16 : 1)
                for (int a = 0; a < AMAX; ++a)
                    if (a < A) act[a] = (uint8_t)p.actions_in[env * A + a];
            }
        } else {
#pragma unroll(AMAX <= 8 ? 16 : 1)
            for (int q = 0; q < (AMAX + 3) / 4; ++q) {
                if (q * 4 < A) {
                    uint32_t r[4];
                    philox4x32_10((uint32_t)(p.env_id_base + (uint64_t)env), (uint32_t)p.t, (uint32_t)q, (uint32_t)(p.t >> 32),
                                  (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (q * 4 + k < AMAX && q * 4 + k < A) act[q * 4 + k] = pick_action(r[k], av[q * 4 + k]);
                }
            }
        }
        if (p.lle_semantics && e.done) {
            err = ERR_DONE;
        } else {
            bool bad = false;
#pragma unroll(AMAX <= 8 ? 16 : 1)
            for (int a = 0; a < AMAX; ++a)
                if (a < A && (act[a] > 4 || !((av[a] >> act[a]) & 1u))) bad = true;
            if (bad) {
                err = ERR_INVALID_ACTION;
            } else {
                StepResult r = env_step(mv, e, act, ev);
                env_reward(e, r, A, p.R, reward);
            }
        }
    } else if (p.mode == MODE_RESET) {
        if (!p.reset_mask || env >= p.N || p.reset_mask[env]) env_reset(mv, e);
        else touch_transition = false;
    } else if (env >= p.N) {
        touch_transition = false;  // padding world: nothing to force
    } else {                       // MODE_SET_STATE
        int32_t si[AMAX], sj[AMAX];
        uint32_t sa = 0;
        uint64_t sg = 0;
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a) {
            si[a] = sj[a] = 0;
            if (a < A) {
                si[a] = p.ss_pos[(env * A + a) * 2];
                sj[a] = p.ss_pos[(env * A + a) * 2 + 1];
                if (p.ss_alive[env * A + a]) sa |= 1u << a;
            }
        }
        for (int g = 0; g < p.G; ++g)
            if (p.ss_gems[env * p.G + g]) sg |= 1ull << g;
        err = env_set_state(mv, e, si, sj, sg, sa, ev, p.lle_semantics != 0);
    }

    if (touch_transition) {
        for (int k = 0; k < p.R; ++k) p.reward[env * p.R + k] = reward[k];
        p.done[env] = (uint8_t)e.done;
        p.err[env] = err;
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a) {
            if (a < A) {
                p.events[env * A + a] = ev[a];
                p.actions[env * A + a] = (int8_t)act[a];
            }
        }
    }
    // auto-reset: the transition above is reported; observation/state/avail below are those of the
    // freshly reset world (SURVEY §8d "Auto-reset")
    if (p.mode == MODE_STEP && p.auto_reset && e.done && err == ERR_OK) env_reset(mv, e);

    env_store(e, p.L, A, NB, [&](int w, uint32_t v) { p.words[(int64_t)w * p.N_pad + env] = v; });
    env_state_vector(mv, e, [&](int k, float v) { p.state[env * p.S + k] = v; });
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) {
        if (a < A) {
            uint32_t m = env_available(mv, e, a);
            if (!p.walkable) m = env_available_no_walk(mv, e, a, m);
#pragma unroll
            for (int k = 0; k < 5; ++k) p.avail[(env * A + a) * 5 + k] = (uint8_t)((m >> k) & 1u);
        }
    }
}

// ---- phase 2 building blocks; `lane` in [0, 32) ------------------------------------------------------
// `sub` holds floats [lo, hi) of one environment's (C,H,W) block.

// (Re)build from the map's static plane (observations.py:216-237); pad floats beyond C*H*W are zero.
LLE_HD void tile_rebuild(float* sub, const uint8_t* blob, int lo, int hi, int lane) {
    const LleMapHeader* hdr = reinterpret_cast<const LleMapHeader*>(blob);
    const float* stat = reinterpret_cast<const float*>(blob + hdr->static_off);
    for (int f = lane * 4; f < hi - lo; f += 128) {
        const int g = lo + f;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = g + k < hdr->obs_floats ? stat[g + k] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) sub[f + k] = v[k];
    }
}

// Un-patch what the previous occupant `old` (stride 1) had lit and the new one `cur` has not.
LLE_HD void tile_unpatch(float* sub, const uint8_t* blob, const uint32_t* old, const uint32_t* cur, int cur_stride, int PW,
                         int A, int HW, int W, int lo, int hi, int lane) {
    const LleMapHeader* hdr = reinterpret_cast<const LleMapHeader*>(blob);
    const LlePatch* patches = reinterpret_cast<const LlePatch*>(blob + hdr->patch_off);
    if (lane < A) {
        uint32_t op = desc_pos(old, 1, lane);
        int idx = lane * HW + (int)(op >> 8) * W + (int)(op & 0xFF);
        if (idx >= lo && idx < hi) sub[idx - lo] = 0.0f;  // agent planes have no static content
    }
    for (int k = lane; k < hdr->n_patch; k += 32) {
        const LlePatch pe = patches[k];
        if ((int)pe.idx >= lo && (int)pe.idx < hi && desc_bit(old, 1, PW, pe) && !desc_bit(cur, cur_stride, PW, pe))
            sub[pe.idx - lo] = (float)pe.stat;
    }
}

// Patch: lit laser cells and uncollected gems, then the agents (observations.py:256-265).  Every lit
// entry is rewritten so that entries aliasing one cell (crossing beams of one colour, colours >=
// n_agents) stay correct whatever was un-patched before.  Must run after tile_unpatch of ALL lanes.
LLE_HD void tile_patch(float* sub, const uint8_t* blob, const uint32_t* cur, int cur_stride, int PW, int A, int HW, int W,
                       int lo, int hi, int lane) {
    const LleMapHeader* hdr = reinterpret_cast<const LleMapHeader*>(blob);
    const LlePatch* patches = reinterpret_cast<const LlePatch*>(blob + hdr->patch_off);
    for (int k = lane; k < hdr->n_patch; k += 32) {
        const LlePatch pe = patches[k];
        if ((int)pe.idx >= lo && (int)pe.idx < hi && desc_bit(cur, cur_stride, PW, pe)) sub[pe.idx - lo] = 1.0f;
    }
    if (lane < A) {
        uint32_t np = desc_pos(cur, cur_stride, lane);
        int idx = lane * HW + (int)(np >> 8) * W + (int)(np & 0xFF);
        if (idx >= lo && idx < hi) sub[idx - lo] = 1.0f;
    }
}

#if defined(__CUDACC__)
// ---- PTX wrappers (TMA 1-D bulk store through the async proxy) ----------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int AMAX, int NBMAX>
__global__ void __launch_bounds__(kThreads) lle_fused_kernel(const KParams p) {
    using D = Desc<AMAX, NBMAX>;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* wbase = smem_raw + (size_t)warp * p.warp_smem_bytes;
    float* tiles = reinterpret_cast<float*>(wbase);                                   // [2][tile_floats]
    uint32_t* descs = reinterpret_cast<uint32_t*>(tiles + 2 * (size_t)p.tile_floats);  // [WORDS][32] of this unit
    uint32_t* applied = descs + D::WORDS * 32;                                        // [2][E][WORDS]
    int32_t* tags = reinterpret_cast<int32_t*>(applied + 2 * p.E * D::WORDS);         // [2][E] map id, then [2] chunk id
    for (int k = lane; k < 2 * p.E + 2; k += 32) tags[k] = -1;
    __syncwarp();
    int buf = 0;

    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(&p.sched[0], 1u);
        unit = __shfl_sync(kFull, unit, 0);
        if (unit >= p.n_units) break;
        const int64_t env = (int64_t)unit * 32 + lane;  // N_pad is a multiple of 32: every lane owns a world

        // ------------------------------------------------------------------ phase 1: thread per env
        const int map_id = p.map_of_env ? __ldg(p.map_of_env + env) : 0;
        const MapView mv = MapView::make(p.blobs[map_id]);
        Env<AMAX, NBMAX> e;
        unit_logic(p, env, mv, e);
        if (!p.write_obs) continue;

        // ------------------------------------------------------------------ phase 2: warp per env
        desc_write(e, descs + lane, 32);  // the 32 descriptors of this unit, word-major (conflict-free)
        __syncwarp();
        const int tiles_per_unit = p.n_chunks > 1 ? 32 : 32 / p.E;
        for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
            const int lo = chunk * p.chunk_floats;  // float range [lo, hi) of one env's block
            const int hi = min(lo + p.chunk_floats, (int)p.obs_stride);
            for (int tix = 0; tix < tiles_per_unit; ++tix) {
                float* tile = tiles + (size_t)buf * p.tile_floats;
                // The bulk store issued two tiles ago read from this buffer: wait until the TMA engine
                // has finished reading it (the other buffer's group may stay in flight).
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
                const int n_sub = p.n_chunks > 1 ? 1 : p.E;
                for (int s = 0; s < n_sub; ++s) {
                    const int l = p.n_chunks > 1 ? tix : tix * p.E + s;  // lane whose world goes into sub-tile s
                    const int mid = __shfl_sync(kFull, map_id, l);
                    const uint8_t* blob = p.blobs[mid];
                    float* sub = tile + (size_t)s * p.obs_stride;
                    uint32_t* old = applied + ((size_t)buf * p.E + s) * D::WORDS;
                    const uint32_t* cur = descs + l;
                    const bool same = tags[buf * p.E + s] == mid && tags[2 * p.E + buf] == chunk;
                    if (!same) tile_rebuild(sub, blob, lo, hi, lane);
                    else tile_unpatch(sub, blob, old, cur, 32, D::PW, p.A, p.HW, p.W, lo, hi, lane);
                    __syncwarp();
                    tile_patch(sub, blob, cur, 32, D::PW, p.A, p.HW, p.W, lo, hi, lane);
                    for (int w = lane; w < D::WORDS; w += 32) old[w] = cur[w * 32];
                    if (lane == 0) tags[buf * p.E + s] = mid;
                }
                if (lane == 0) tags[2 * p.E + buf] = chunk;
                fence_proxy_async_smem();  // generic-proxy writes above -> visible to the async proxy
                __syncwarp();
                if (lane == 0) {
                    const int64_t first_env = (int64_t)unit * 32 + (p.n_chunks > 1 ? tix : tix * p.E);
                    float* dst = p.obs + first_env * p.obs_stride + lo;
                    const uint32_t bytes = (uint32_t)((p.n_chunks > 1 ? (hi - lo) : p.E * (int)p.obs_stride) * 4);
                    bulk_store(dst, tile, bytes);
                    bulk_commit();
                }
                buf ^= 1;
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        bulk_wait_all();
        __threadfence();
        uint32_t finished = atomicAdd(&p.sched[1], 1u);
        if (finished == p.n_warps_total - 1) {  // last warp out re-arms the ticket counter for the next launch
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
        }
    }
}
#endif  // __CUDACC__

// buckets the kernel is instantiated for: (max agents, max laser sources)
struct Bucket {
    int amax, nbmax;
};
constexpr Bucket kBuckets[] = {{4, 4}, {8, 8}, {8, 16}, {16, 16}};
constexpr int kNumBuckets = 4;

}  // namespace lle
