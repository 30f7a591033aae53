// Per-environment transition function of the batched `World`, written once as __host__ __device__
// code on a register-resident bitmask state.  One CUDA thread owns one environment; the loops over
// agents and beams are compile-time bounded (AMAX, NBMAX) so the whole record lives in registers.
//
// This is NOT a translation of the reference's object graph (Vec<Vec<Tile>>, Rc<LaserBeam>,
// RefCell<Vec<bool>>): a beam is one u64 mask, a tile slot is one bit per agent, and the nested
// `Laser(Laser(tile))` dispatch becomes a scan over the beams crossing a cell.  The *observable*
// semantics are the reference's, cited per function (paths relative to the reference repository).
#pragma once
#include <stdint.h>

#include "static_map.h"

#if defined(__CUDACC__)
#define LLE_HD __host__ __device__ __forceinline__
#else
#define LLE_HD inline
#endif

namespace lle {

// event codes of one agent in one pass (also bits 0-1 of the exported event byte)
enum : uint32_t { EV_NONE = 0, EV_EXIT = 1, EV_GEM = 2, EV_DIED = 3 };

// per-env error codes written to the `err` buffer
enum : uint8_t {
    ERR_OK = 0,
    ERR_INVALID_ACTION = 1,        // RuntimeWorldError::InvalidAction        (world.rs:444-453)
    ERR_DONE = 2,                  // "Cannot step in a done environment"      (env.py:166-167)
    ERR_STATE_DUPLICATE = 3,       // InvalidWorldState "two agents at the same position" (world.rs:529-534)
    ERR_STATE_OUT_OF_WORLD = 4,    // OutOfWorldPosition                       (world.rs:536-540)
    ERR_STATE_NOT_WALKABLE = 5,    // InvalidAgentPosition                     (world.rs:556-568)
    ERR_STATE_MISMATCH = 6,        // InvalidWorldState "the given state is invalid" (world.rs:588-594)
};

struct MapView {
    const LleMapHeader* hdr;
    const uint16_t* tiles;
    const LleBeam* beams;
    LLE_HD static MapView make(const uint8_t* blob) {
        MapView m;
        m.hdr = reinterpret_cast<const LleMapHeader*>(blob);
        m.tiles = reinterpret_cast<const uint16_t*>(blob + m.hdr->tiles_off);
        m.beams = reinterpret_cast<const LleBeam*>(blob + m.hdr->beams_off);
        return m;
    }
};

LLE_HD uint64_t len_mask(int len) { return len >= 64 ? ~0ull : ((1ull << len) - 1ull); }

// offset of packed position p on beam b, or -1 (replaces the per-tile `Laser.offset`, laser.rs:91)
LLE_HD int beam_offset(const LleBeam& b, int i, int j) {
    int d_i = i - (int)b.first_i, d_j = j - (int)b.first_j;
    int k = d_i * b.di + d_j * b.dj;
    int perp = d_i * b.dj - d_j * b.di;
    return (perp == 0 && k >= 0 && k < (int)b.len) ? k : -1;
}

// Action deltas, src/action.rs:18-26 (N=0, S=1, E=2, W=3, STAY=4)
LLE_HD int act_di(int a) { return a == 0 ? -1 : (a == 1 ? 1 : 0); }
LLE_HD int act_dj(int a) { return a == 2 ? 1 : (a == 3 ? -1 : 0); }

template <int AMAX, int NBMAX>
struct Env {
    uint16_t pos[AMAX];   // packed (i<<8 | j)
    uint64_t on[NBMAX];   // LaserBeam.beam as a bitmask
    uint64_t collected;   // Gem.collected per gem (wrapped gems included)
    uint32_t alive, arrived, slot;
    uint32_t n_arrived, n_deads, done;  // RewardStrategy.n_arrived / n_deads (reward_strategy.py:31-38), LLE.done
};

struct StepResult {
    uint32_t n_gem, n_exit, n_died;
    bool any_event;
};

// ---- record <-> registers ------------------------------------------------------------------------
template <int AMAX, int NBMAX, class Load>
LLE_HD void env_load(Env<AMAX, NBMAX>& e, const LleStateLayout& L, int A, int NB, Load&& ld) {
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int w = 0; w < (AMAX + 1) / 2; ++w) {
        if (w < L.w_flags) {
            uint32_t v = ld(w);
            e.pos[2 * w] = (uint16_t)(v & 0xFFFFu);
            if (2 * w + 1 < AMAX) e.pos[2 * w + 1] = (uint16_t)(v >> 16);
        }
    }
    if (!L.wide_flags) {
        uint32_t f = ld(L.w_flags);
        e.alive = f & 0xFFu;
        e.arrived = (f >> 8) & 0xFFu;
        e.slot = (f >> 16) & 0xFFu;
        e.n_arrived = (f >> 24) & 0xFu;
        e.n_deads = (f >> 28) & 0x7u;
        e.done = f >> 31;
    } else {
        e.alive = ld(L.w_flags);
        e.arrived = ld(L.w_flags + 1);
        e.slot = ld(L.w_flags + 2);
        uint32_t m = ld(L.w_flags + 3);
        e.n_arrived = m & 0xFFu;
        e.n_deads = (m >> 8) & 0xFFu;
        e.done = (m >> 16) & 1u;
    }
    e.collected = 0;
    if (L.gem_words >= 1) e.collected = ld(L.w_gems);
    if (L.gem_words == 2) e.collected |= (uint64_t)ld(L.w_gems + 1) << 32;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int b = 0; b < NBMAX; ++b) {
        e.on[b] = 0;
        if (b < NB) {
            e.on[b] = ld(L.w_on + b * L.on_words);
            if (L.on_words == 2) e.on[b] |= (uint64_t)ld(L.w_on + b * 2 + 1) << 32;
        }
    }
    (void)A;
}

template <int AMAX, int NBMAX, class Store>
LLE_HD void env_store(const Env<AMAX, NBMAX>& e, const LleStateLayout& L, int A, int NB, Store&& st) {
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int w = 0; w < (AMAX + 1) / 2; ++w) {
        if (w < L.w_flags) {
            uint32_t v = e.pos[2 * w];
            if (2 * w + 1 < AMAX && 2 * w + 1 < A) v |= (uint32_t)e.pos[2 * w + 1] << 16;
            st(w, v);
        }
    }
    if (!L.wide_flags) {
        uint32_t nd = e.n_deads > 7u ? 7u : e.n_deads;  // only "> 0" is ever observed (env.py:253-254)
        st(L.w_flags, e.alive | (e.arrived << 8) | (e.slot << 16) | (e.n_arrived << 24) | (nd << 28) | (e.done << 31));
    } else {
        st(L.w_flags, e.alive);
        st(L.w_flags + 1, e.arrived);
        st(L.w_flags + 2, e.slot);
        st(L.w_flags + 3, (e.n_arrived & 0xFFu) | ((e.n_deads > 255u ? 255u : e.n_deads) << 8) | (e.done << 16));
    }
    if (L.gem_words >= 1) st(L.w_gems, (uint32_t)e.collected);
    if (L.gem_words == 2) st(L.w_gems + 1, (uint32_t)(e.collected >> 32));
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int b = 0; b < NBMAX; ++b) {
        if (b < NB) {
            st(L.w_on + b * L.on_words, (uint32_t)e.on[b]);
            if (L.on_words == 2) st(L.w_on + b * 2 + 1, (uint32_t)(e.on[b] >> 32));
        }
    }
}

// ---- tile protocol --------------------------------------------------------------------------------
// Tile::leave for an alive agent standing on packed position p (tile.rs:52-61, laser.rs:199-202 +
// :157-162 + :50-55): every beam through the cell whose bit there is off is re-armed from that
// offset on (unless the source is disabled); the base tile's slot is cleared.
template <int AMAX, int NBMAX>
LLE_HD void tile_leave(const MapView& m, Env<AMAX, NBMAX>& e, int a, int NB) {
    int i = e.pos[a] >> 8, j = e.pos[a] & 0xFF;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int b = 0; b < NBMAX; ++b) {
        if (b < NB) {
            const LleBeam bm = m.beams[b];
            int k = beam_offset(bm, i, j);
            if (k >= 0 && bm.enabled && !((e.on[b] >> k) & 1ull)) e.on[b] |= (~0ull << k) & len_mask(bm.len);
        }
    }
    e.slot &= ~(1u << a);
}

// Tile::pre_enter on packed position p (tile.rs:21-27, laser.rs:173-182): an alive agent cuts the
// enabled beams of its own colour from its offset on.
template <int AMAX, int NBMAX>
LLE_HD void tile_pre_enter(const MapView& m, Env<AMAX, NBMAX>& e, int a, uint16_t p, int NB) {
    if (!((e.alive >> a) & 1u)) return;
    int i = p >> 8, j = p & 0xFF;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int b = 0; b < NBMAX; ++b) {
        if (b < NB) {
            const LleBeam bm = m.beams[b];
            if (bm.enabled && bm.colour == a) {
                int k = beam_offset(bm, i, j);
                if (k >= 0) e.on[b] &= (1ull << k) - 1ull;
            }
        }
    }
}

// Tile::enter on packed position p (tile.rs:29-50, laser.rs:184-197, gem.rs:26-35, void.rs:13-22).
// An on beam of another colour stops the agent before the wrapped tile: alive -> dies; dead -> nothing;
// in both cases the base tile is NOT entered.  Otherwise the base tile takes the agent.
template <int AMAX, int NBMAX>
LLE_HD uint32_t tile_enter(const MapView& m, Env<AMAX, NBMAX>& e, int a, uint16_t p, int NB) {
    int i = p >> 8, j = p & 0xFF;
    bool lethal = false;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int b = 0; b < NBMAX; ++b) {
        if (b < NB) {
            const LleBeam bm = m.beams[b];
            int k = beam_offset(bm, i, j);
            if (k >= 0 && ((e.on[b] >> k) & 1ull) && bm.colour != a) lethal = true;
        }
    }
    const uint32_t bit = 1u << a;
    if (lethal) {
        if (e.alive & bit) {
            e.alive &= ~bit;
            return EV_DIED;
        }
        return EV_NONE;
    }
    uint16_t tile = m.tiles[i * m.hdr->W + j];
    e.slot |= bit;
    switch (tile & 7u) {
        case LLE_T_EXIT:
            if (!(e.arrived & bit)) {
                e.arrived |= bit;
                return EV_EXIT;
            }
            return EV_NONE;
        case LLE_T_GEM: {
            uint64_t g = 1ull << (tile >> 8);
            if (!(e.collected & g)) {
                e.collected |= g;
                return EV_GEM;
            }
            return EV_NONE;
        }
        case LLE_T_VOID:
            if (e.alive & bit) {
                e.alive &= ~bit;
                return EV_DIED;
            }
            return EV_NONE;
        default: return EV_NONE;
    }
}

// All tiles reset (tile.rs:75-84; laser.rs:168-171 => an enabled beam ends fully on, a disabled one
// stays fully off; gem.rs:21-24).  Agents are NOT touched.
template <int AMAX, int NBMAX>
LLE_HD void tiles_reset(const MapView& m, Env<AMAX, NBMAX>& e, int NB) {
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int b = 0; b < NBMAX; ++b) {
        if (b < NB) {
            const LleBeam bm = m.beams[b];
            e.on[b] = bm.enabled ? len_mask(bm.len) : 0ull;
        }
    }
    e.collected = 0;
    e.slot = 0;
}

// World::reset (world.rs:411-432) with one start per agent (RNG-free, utils/mod.rs:63) followed by
// RewardStrategy.reset / LLE.reset bookkeeping (env.py:191-203).
template <int AMAX, int NBMAX>
LLE_HD void env_reset(const MapView& m, Env<AMAX, NBMAX>& e) {
    const int A = m.hdr->A, NB = m.hdr->NB;
    tiles_reset(m, e, NB);
    e.alive = A >= 32 ? ~0u : ((1u << A) - 1u);
    e.arrived = 0;
    e.n_arrived = 0;
    e.n_deads = 0;
    e.done = 0;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) e.pos[a] = a < A ? m.hdr->start[a] : (uint16_t)0;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a)
        if (a < A) tile_pre_enter(m, e, a, e.pos[a], NB);
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a)
        if (a < A) (void)tile_enter(m, e, a, e.pos[a], NB);  // events are dropped (world.rs:428-430)
}

// World::compute_available_actions (world.rs:343-363) as a 5-bit mask indexed by Action value.
template <int AMAX, int NBMAX>
LLE_HD uint32_t env_available(const MapView& m, const Env<AMAX, NBMAX>& e, int a) {
    uint32_t mask = 1u << 4;  // Stay
    const uint32_t bit = 1u << a;
    if (!(e.alive & bit) || (e.arrived & bit)) return mask;
    const int H = m.hdr->H, W = m.hdr->W, A = m.hdr->A;
    const int i = e.pos[a] >> 8, j = e.pos[a] & 0xFF;
#pragma unroll
    for (int act = 0; act < 4; ++act) {
        int ti = i + act_di(act), tj = j + act_dj(act);
        if (ti < 0 || tj < 0 || ti >= H || tj >= W) continue;
        if ((m.tiles[ti * W + tj] & 7u) == LLE_T_WALL) continue;  // Wall or LaserSource (tile.rs:63-73)
        uint16_t tp = (uint16_t)((ti << 8) | tj);
        bool occupied = false;
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int o = 0; o < AMAX; ++o)
            if (o < A && ((e.slot >> o) & 1u) && e.pos[o] == tp) occupied = true;  // Tile::is_occupied
        if (!occupied) mask |= 1u << act;
    }
    return mask;
}

// LLE.available_actions with walkable_lasers == False (env.py:153-163): additionally drop every listed
// action whose target holds an *on*, *listed* (world.rs:159-172) laser of another colour.
template <int AMAX, int NBMAX>
LLE_HD uint32_t env_available_no_walk(const MapView& m, const Env<AMAX, NBMAX>& e, int a, uint32_t mask) {
    const int NB = m.hdr->NB;
    const int i = e.pos[a] >> 8, j = e.pos[a] & 0xFF;
    uint32_t out = 0;
#pragma unroll
    for (int act = 0; act < 5; ++act) {
        if (!((mask >> act) & 1u)) continue;
        int ti = i + act_di(act), tj = j + act_dj(act);
        bool blocked = false;
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int b = 0; b < NBMAX; ++b) {
            if (b < NB) {
                const LleBeam bm = m.beams[b];
                int k = beam_offset(bm, ti, tj);
                if (k >= 0 && ((bm.vis >> k) & 1ull) && ((e.on[b] >> k) & 1ull) && bm.colour != a) blocked = true;
            }
        }
        if (!blocked) out |= 1u << act;
    }
    return out;
}

// World::step (world.rs:435-475) after validation: new positions, vertex conflicts
// (world.rs:365-378 + utils/mod.rs:18-36), then move_agents passes (world.rs:477-505) repeated while
// an agent died (world.rs:468-472).  `ev[a]` receives the exported event byte:
// bits 0-1 = pass-1 event, bits 2-7 = pass (>= 2) in which the agent died.
template <int AMAX, int NBMAX>
LLE_HD StepResult env_step(const MapView& m, Env<AMAX, NBMAX>& e, const uint8_t* act, uint8_t* ev) {
    const int A = m.hdr->A, NB = m.hdr->NB;
    uint16_t np[AMAX];
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) {
        np[a] = 0;
        if (a < A) {
            int i = (e.pos[a] >> 8) + act_di(act[a]), j = (e.pos[a] & 0xFF) + act_dj(act[a]);
            np[a] = (uint16_t)((i << 8) | j);
        }
        ev[a] = 0;
    }
    // vertex conflicts: every agent whose target is shared goes back to where it stands; repeat.
    for (;;) {
        uint32_t dup = 0;
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a)
#pragma unroll(AMAX <= 8 ? 16 : 1)
            for (int o = a + 1; o < AMAX; ++o)
                if (o < A && np[a] == np[o]) dup |= (1u << a) | (1u << o);
        if (!dup) break;
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a)
            if (a < A && ((dup >> a) & 1u)) np[a] = e.pos[a];
    }
    StepResult r{0, 0, 0, false};
    uint32_t pass = 1;
    bool died;
    do {
        died = false;
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a)
            if (a < A && ((e.alive >> a) & 1u)) tile_leave(m, e, a, NB);
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a)
            if (a < A) tile_pre_enter(m, e, a, np[a], NB);
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a) {
            if (a < A) {
                uint32_t code = tile_enter(m, e, a, np[a], NB);
                if (code != EV_NONE) {
                    r.any_event = true;
                    if (code == EV_DIED) { died = true; r.n_died++; }
                    else if (code == EV_GEM) r.n_gem++;
                    else r.n_exit++;
                    // passes >= 2 can only emit deaths; keep the first pass's code in bits 0-1
                    if (pass == 1) ev[a] |= (uint8_t)code;
                    else ev[a] |= (uint8_t)((pass > 63 ? 63 : pass) << 2);
                }
            }
        }
        if (pass == 1) {
#pragma unroll(AMAX <= 8 ? 16 : 1)
            for (int a = 0; a < AMAX; ++a)
                if (a < A) e.pos[a] = np[a];
        }
        ++pass;
    } while (died);
    return r;
}

// SingleObjective / MultiObjective.compute_reward (reward_strategy.py:58-75, :90-109) and
// LLE.compute_done (env.py:253-254).  reward has reward_dim (1 or 4) entries.
template <int AMAX, int NBMAX>
LLE_HD void env_reward(Env<AMAX, NBMAX>& e, const StepResult& r, int A, int reward_dim, float* reward) {
    e.n_arrived += r.n_exit;
    e.n_deads += r.n_died;
    if (reward_dim == 1) {
        float v = (float)r.n_gem + (float)r.n_exit - (float)r.n_died;
        if (e.n_arrived == (uint32_t)A) v += 1.0f;  // REWARD_DONE; the death override is dead code (:71-72)
        reward[0] = v;
    } else {
        reward[0] = (float)r.n_gem;
        reward[1] = (float)r.n_exit;
        reward[2] = -(float)r.n_died;
        reward[3] = 0.0f;
        if (r.n_died) {
            reward[0] = 0.0f;
            reward[1] = 0.0f;
        } else if (e.n_arrived == (uint32_t)A) {
            reward[3] = 1.0f;
        }
    }
    e.done = (e.n_arrived == (uint32_t)A || e.n_deads > 0) ? 1u : 0u;
}

// Body of World::set_state after the argument checks (world.rs:541-594): tiles reset, forced gem
// collection, pre_enter with the agents' *current* alive flags, then reset+enter per agent and the
// final equality check.  Returns ERR_OK, ERR_STATE_NOT_WALKABLE (nothing entered yet) or
// ERR_STATE_MISMATCH (world left modified, as in the reference).
template <int AMAX, int NBMAX>
LLE_HD uint8_t set_state_apply(const MapView& m, Env<AMAX, NBMAX>& e, const int32_t* si, const int32_t* sj, uint64_t sg,
                               uint32_t sa, uint8_t* ev, StepResult& r) {
    const int A = m.hdr->A, NB = m.hdr->NB, W = m.hdr->W;
    tiles_reset(m, e, NB);
    e.collected = sg & m.hdr->gem_toplevel;  // only top-level Gem tiles can be force-collected (:550-554)
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) {
        if (a < A) {
            if ((m.tiles[si[a] * W + sj[a]] & 7u) == LLE_T_WALL) return ERR_STATE_NOT_WALKABLE;  // :556-568
            tile_pre_enter(m, e, a, (uint16_t)((si[a] << 8) | sj[a]), NB);
        }
    }
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a)
        if (a < A) e.pos[a] = (uint16_t)((si[a] << 8) | sj[a]);  // :571
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) {
        if (a < A) {
            e.alive |= 1u << a;  // agent.reset() (:578)
            e.arrived &= ~(1u << a);
            uint32_t code = tile_enter(m, e, a, e.pos[a], NB);
            if (code != EV_NONE) {
                ev[a] = (uint8_t)code;
                if (code == EV_DIED) r.n_died++;
                else if (code == EV_GEM) r.n_gem++;
                else r.n_exit++;
            }
            if (!((sa >> a) & 1u)) e.alive &= ~(1u << a);  // forced death, no event (:583-585)
        }
    }
    const uint32_t amask = A >= 32 ? ~0u : ((1u << A) - 1u);
    if (e.collected != sg || (e.alive & amask) != (sa & amask)) return ERR_STATE_MISMATCH;  // :588-594
    return ERR_OK;
}

// World::set_state (world.rs:515-597), sizes already checked on the host.  `si/sj` are the requested
// positions, `sg` the collected mask, `sa` the alive mask.  Returns an ERR_* code; `ev[a]` receives the
// event codes.  `lle_level` additionally applies LLE.set_state's bookkeeping (env.py:208-216).
template <int AMAX, int NBMAX>
LLE_HD uint8_t env_set_state(const MapView& m, Env<AMAX, NBMAX>& e, const int32_t* si, const int32_t* sj, uint64_t sg,
                             uint32_t sa, uint8_t* ev, bool lle_level) {
    const int A = m.hdr->A, H = m.hdr->H, W = m.hdr->W;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) ev[a] = 0;
    if (lle_level) {  // reward_strategy.reset() precedes world.set_state (env.py:213)
        e.n_arrived = 0;
        e.n_deads = 0;
    }
    // :529-534 duplicates, then :536-540 bounds (in that order)
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a)
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int o = a + 1; o < AMAX; ++o)
            if (o < A && si[a] == si[o] && sj[a] == sj[o]) return ERR_STATE_DUPLICATE;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a)
        if (a < A && (si[a] < 0 || sj[a] < 0 || si[a] >= H || sj[a] >= W)) return ERR_STATE_OUT_OF_WORLD;
    // current_state = self.get_state() (:541)
    int32_t ci[AMAX], cj[AMAX];
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) {
        ci[a] = e.pos[a] >> 8;
        cj[a] = e.pos[a] & 0xFF;
    }
    const uint64_t cg = e.collected;
    const uint32_t ca = e.alive;
    StepResult r{0, 0, 0, false};
    uint8_t code = set_state_apply(m, e, si, sj, sg, sa, ev, r);
    if (code == ERR_STATE_NOT_WALKABLE) {
        // self.set_state(&current_state).unwrap() (:563): the previous state is re-derived, events dropped
        uint8_t scratch[AMAX];
        StepResult r2{0, 0, 0, false};
#pragma unroll(AMAX <= 8 ? 16 : 1)
        for (int a = 0; a < AMAX; ++a) scratch[a] = 0;
        (void)set_state_apply(m, e, ci, cj, cg, ca, scratch, r2);
        return code;
    }
    if (code != ERR_OK) return code;
    if (lle_level) {
        float scratch[4];
        env_reward(e, r, A, 1, scratch);  // compute_reward(events); done = compute_done() (env.py:215-216)
    }
    return ERR_OK;
}

// ---- exported per-env vectors ---------------------------------------------------------------------
// PyWorldState::as_array (src/bindings/world/pyworld_state.rs:79-101): [i0,j0,...,gems...,alive...]
template <int AMAX, int NBMAX, class Put>
LLE_HD void env_state_vector(const MapView& m, const Env<AMAX, NBMAX>& e, Put&& put) {
    const int A = m.hdr->A, G = m.hdr->G;
#pragma unroll(AMAX <= 8 ? 16 : 1)
    for (int a = 0; a < AMAX; ++a) {
        if (a < A) {
            put(2 * a, (float)(e.pos[a] >> 8));
            put(2 * a + 1, (float)(e.pos[a] & 0xFF));
            put(2 * A + G + a, ((e.alive >> a) & 1u) ? 1.0f : 0.0f);
        }
    }
    for (int g = 0; g < G; ++g) put(2 * A + g, ((e.collected >> g) & 1ull) ? 1.0f : 0.0f);
}

// ---- Philox4x32-10 action stream -------------------------------------------------------
LLE_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

LLE_HD int popc5(uint32_t m) {
    m &= 31u;
    return (int)((m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u) + ((m >> 4) & 1u));
}

// k-th set bit of a 5-bit availability mask, k = mulhi(word, popcount)
LLE_HD uint8_t pick_action(uint32_t word, uint32_t mask) {
    uint32_t n = (uint32_t)popc5(mask);
    uint32_t k = (uint32_t)(((uint64_t)word * n) >> 32);
    uint8_t res = 4;
    bool found = false;
#pragma unroll
    for (uint32_t b = 0; b < 5; ++b) {
        if ((mask >> b) & 1u) {
            if (k == 0 && !found) { res = (uint8_t)b; found = true; }
            --k;
        }
    }
    return res;
}

}  // namespace lle
