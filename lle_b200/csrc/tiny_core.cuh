// Per-world core of the tiny-map step kernel (tiny_kernel.cuh): one thread owns one whole world.
//
// `TinyWorld<A_, Rec>` holds the world's positions and flags in registers and reaches the rest of its record (beam masks,
// gem mask, availability cache; layout: static_map.h) through `Rec`, an accessor `uint32_t& operator()(int word)` — a
// lane-private column of shared memory in the kernel, a plain array in the host instantiation.  The agents are walked by
// compile-time loops in the reference's own order (world.rs:435-505): leave all, pre_enter all, enter all, repeated while
// somebody died — sequential like the reference, so no order-independence argument is needed here.
//
// __host__ __device__: tests/host_shim/tiny_host.cpp instantiates the same code with g++ so that the CPU suite can compare
// it with the oracle bit for bit without a GPU; the product only ever runs it inside lle_tiny_step_kernel.
#pragma once
#include "step_common.cuh"

namespace lle {

struct TinyLayout {  // the part of LleStateLayout a tiny record uses (n_words <= 8, one word per beam, <= 32 gems)
    int w_flags, w_avail, w_gems, w_on, stride;
    bool has_gems;
};

// ---- PartialGenerator.observe (observations.py:331-350): agent a sees a (2A+3, size, size) window centred on itself; channels:
// the agents, WALL, one laser channel per agent colour, GEM, EXIT.  `sub` is the world's ZERO-FILLED block of A windows.
// One task = (agent a, window cell r): look the map cell up once and write its few non-zero channels.  `recw(word)` reads a word of
// the world's record (static_map.h; tiny records: one word per beam, one gem word); pa = packed position of agent a.
template <class RecW>
LLE_HD void partial_cell_task(float* sub, int A, int a, int r, int sz, uint32_t pa, int H, int W, const uint32_t* cellinfo,
                              const LleCellBeams* cellbeams, int w_gems, int w_on, RecW recw) {
    const int s2 = sz * sz, ctr = sz >> 1, Cp = 2 * A + 3;
    const int wi = (int)(((float)r + 0.5f) * (1.0f / (float)sz)), wj = r - wi * sz;  // exact for these small integers
    const int i = (int)(pa >> 8) + wi - ctr, j = (int)(pa & 0xFFu) + wj - ctr;
    if (i < 0 || j < 0 || i >= H || j >= W) return;  // outside the map: every layer stays 0 (:325-329)
    const int c = i * W + j;
    const uint32_t info = cellinfo[c];
    if (!(info & (7u | (1u << 24)))) return;  // plain floor without a beam
    float* o = sub + a * Cp * s2 + r;
    const uint32_t kind = info & 7u;
    if (kind == LLE_T_WALL) {
        if (info & (1u << 25)) o[A * s2] = 1.0f;  // World::walls(): the walls and the v1 sources, not the TOML [[lasers]] sources
        if (info & 128u) {
            const int ch = A + 1 + (int)((info >> 16) & 255u);  // LASER_0 + source.agent_id, filled with -1 (:348-350)
            if (ch < Cp) o[ch * s2] = -1.0f;
        }
    } else if (kind == LLE_T_EXIT) {
        o[(2 * A + 2) * s2] = 1.0f;
    } else if (kind == LLE_T_GEM) {
        if (!((recw(w_gems) >> ((info >> 8) & 31u)) & 1u)) o[(2 * A + 1) * s2] = 1.0f;
    }
    if (info & (1u << 24)) {
        const LleCellBeams cb = cellbeams[c];
        LLE_UNROLL
        for (int n = 0; n < 4; ++n) {  // lit lasers listed by World::lasers (:352-360)
            const uint32_t e = cb.e[n];
            if (e == LLE_NO_BEAM) break;
            if (be_listed(e) && ((recw(w_on + be_b(e)) >> be_k(e)) & 1u)) {
                const int ch = A + 1 + be_colour(e);
                if (ch < Cp) o[ch * s2] = 1.0f;
            }
        }
    }
}
// agent layers (:335-336): agent a2 as seen from agent a, one task per ordered pair
LLE_HD void partial_agent_task(float* sub, int A, int a, int a2, int sz, uint32_t pa, uint32_t pb) {
    const int ctr = sz >> 1;
    const int di = (int)(pb >> 8) - (int)(pa >> 8) + ctr, dj = (int)(pb & 0xFFu) - (int)(pa & 0xFFu) + ctr;
    if (di >= 0 && dj >= 0 && di < sz && dj < sz) sub[(a * (2 * A + 3) + a2) * sz * sz + di * sz + dj] = 1.0f;
}

template <int A_, class Rec>
struct TinyWorld {
    Rec rec;
    TinyLayout L;
    int W;
    const uint8_t* blob;
    const LleMapHeader* hdr;
    const uint32_t* cellinfo;
    const LleCellBeams* cellbeams;
    const LlePatch* list;  // the map's render list (static_map.h: static_list_off): n_static static floats, then n_patch dynamic cells
    int n_static, n_patch;
    uint32_t pos[A_];
    uint32_t alive, arrived, slot, n_arrived, n_deads, done;

    LLE_HD void bind(const uint8_t* b) {
        blob = b;
        hdr = reinterpret_cast<const LleMapHeader*>(b);
        cellinfo = reinterpret_cast<const uint32_t*>(b + hdr->cellinfo_off);
        cellbeams = reinterpret_cast<const LleCellBeams*>(b + hdr->cellbeams_off);
        list = reinterpret_cast<const LlePatch*>(b + hdr->static_list_off);
        n_static = hdr->n_static;
        n_patch = hdr->n_patch;
    }
    LLE_HD uint32_t cellof(uint32_t q) const { return (q >> 8) * (uint32_t)W + (q & 0xFFu); }
    LLE_HD uint32_t& on_word(int b) { return rec(L.w_on + b); }

    LLE_HD void unpack() {
        LLE_UNROLL
        for (int a = 0; a < A_; ++a) {
            const uint32_t w = rec(a >> 1);
            pos[a] = (a & 1) ? (w >> 16) : (w & 0xFFFFu);
        }
        const uint32_t f = rec(L.w_flags);
        alive = f & 0xFFu; arrived = (f >> 8) & 0xFFu; slot = (f >> 16) & 0xFFu;
        n_arrived = (f >> 24) & 0xFu; n_deads = (f >> 28) & 0x7u; done = f >> 31;
    }
    LLE_HD void pack(uint32_t avail_cache) {
        LLE_UNROLL
        for (int a = 0; a < A_; a += 2) rec(a >> 1) = pos[a] | ((a + 1 < A_ ? pos[(a + 1 < A_) ? a + 1 : a] : 0u) << 16);
        rec(L.w_flags) = alive | (arrived << 8) | (slot << 16) | (n_arrived << 24) | ((n_deads > 7u ? 7u : n_deads) << 28) | (done << 31);
        rec(L.w_avail) = avail_cache;
    }

    // Tile::leave of an alive agent standing on q (tile.rs:52-61, laser.rs:199-202 + :157-162 + :50-55)
    LLE_HD void leave(uint32_t q, uint32_t info) {
        if (!(info & (1u << 24))) return;
        const LleCellBeams cb = cellbeams[cellof(q)];
        LLE_UNROLL
        for (int n = 0; n < 4; ++n) {
            const uint32_t e = cb.e[n];
            if (e == LLE_NO_BEAM) break;
            if (be_enabled(e)) {
                uint32_t& w = on_word(be_b(e));
                if (!((w >> be_k(e)) & 1u)) w |= (~0u << be_k(e)) & len_mask32(be_len(e));
            }
        }
    }
    // Tile::pre_enter (tile.rs:21-27, laser.rs:173-182): an alive agent cuts the enabled beams of its own colour from its offset on
    LLE_HD void pre_enter(int a, uint32_t q, uint32_t info) {
        if (!((alive >> a) & 1u) || !(info & (1u << 24))) return;
        const LleCellBeams cb = cellbeams[cellof(q)];
        LLE_UNROLL
        for (int n = 0; n < 4; ++n) {
            const uint32_t e = cb.e[n];
            if (e == LLE_NO_BEAM) break;
            if (be_enabled(e) && be_colour(e) == a) on_word(be_b(e)) &= (1u << be_k(e)) - 1u;
        }
    }
    // Tile::enter (tile.rs:29-50, laser.rs:184-197, gem.rs:26-35, void.rs:13-22): an on beam of another colour stops the agent
    // before the wrapped tile (alive -> dies, dead -> nothing); otherwise the base tile takes it.  Returns the event code.
    LLE_HD uint32_t enter(int a, uint32_t q, uint32_t info) {
        if (info & (1u << 24)) {
            const LleCellBeams cb = cellbeams[cellof(q)];
            bool lethal = false;
            LLE_UNROLL
            for (int n = 0; n < 4; ++n) {
                const uint32_t e = cb.e[n];
                if (e == LLE_NO_BEAM) break;
                if (be_colour(e) != a && ((on_word(be_b(e)) >> be_k(e)) & 1u)) lethal = true;
            }
            if (lethal) {
                if ((alive >> a) & 1u) { alive &= ~(1u << a); return EV_DIED; }
                return EV_NONE;
            }
        }
        slot |= 1u << a;
        const uint32_t kind = info & 7u;
        if (kind == LLE_T_EXIT) {
            if (!((arrived >> a) & 1u)) { arrived |= 1u << a; return EV_EXIT; }
        } else if (kind == LLE_T_GEM) {
            const uint32_t g = (info >> 8) & 63u;
            uint32_t& gw = rec(L.w_gems);
            if (!((gw >> g) & 1u)) { gw |= 1u << g; return EV_GEM; }
        } else if (kind == LLE_T_VOID) {
            if ((alive >> a) & 1u) { alive &= ~(1u << a); return EV_DIED; }
        }
        return EV_NONE;
    }
    // World::reset (world.rs:411-432) with one start per agent (no random number is drawn, utils/mod.rs:63) followed by
    // RewardStrategy.reset / LLE.reset bookkeeping (env.py:191-203)
    LLE_HD void reset() {
        const int nb = hdr->NB;
        const LleBeam* beams = reinterpret_cast<const LleBeam*>(blob + hdr->beams_off);
        for (int b = 0; b < nb; ++b) on_word(b) = beams[b].enabled ? len_mask32(beams[b].len) : 0u;
        if (L.has_gems) rec(L.w_gems) = 0u;
        slot = 0; alive = (1u << A_) - 1u; arrived = 0; n_arrived = 0; n_deads = 0; done = 0;
        uint32_t info[A_];
        LLE_UNROLL
        for (int a = 0; a < A_; ++a) {
            pos[a] = hdr->start[a];
            info[a] = cellinfo[cellof(pos[a])];
        }
        LLE_UNROLL
        for (int a = 0; a < A_; ++a) pre_enter(a, pos[a], info[a]);
        LLE_UNROLL
        for (int a = 0; a < A_; ++a) (void)enter(a, pos[a], info[a]);  // events are dropped (world.rs:428-430)
    }
    // World::step after validation (world.rs:454-472) + reward bookkeeping (reward_strategy.py:58-109, env.py:253-254).
    // ev[a]: bits 0-1 the event of the first pass, bits 2-7 the pass (>= 2) in which the agent died.
    LLE_HD void step(const uint32_t (&act)[A_], uint32_t (&ev)[A_], uint32_t& n_gem, uint32_t& n_exit, uint32_t& n_died) {
        uint32_t np[A_];
        LLE_UNROLL
        for (int a = 0; a < A_; ++a) np[a] = pos[a] + (uint32_t)act_delta((int)act[a]);
        // vertex conflicts (world.rs:365-378 + utils/mod.rs:18-36): every agent whose target is shared goes back, until none is
        if (A_ > 1) {
            for (;;) {
                uint32_t dup = 0;
                LLE_UNROLL
                for (int a = 0; a < A_; ++a)
                    LLE_UNROLL
                    for (int o = 0; o < A_; ++o)
                        if (o != a && np[o] == np[a]) dup |= 1u << a;
                if (!dup) break;
                LLE_UNROLL
                for (int a = 0; a < A_; ++a)
                    if ((dup >> a) & 1u) np[a] = pos[a];
            }
        }
        uint32_t info_new[A_];
        LLE_UNROLL
        for (int a = 0; a < A_; ++a) info_new[a] = cellinfo[cellof(np[a])];
        n_gem = n_exit = n_died = 0;
        // move_agents (world.rs:477-505), repeated while an agent died (:468-472)
        for (uint32_t pass = 1;; ++pass) {
            LLE_UNROLL
            for (int a = 0; a < A_; ++a) {  // the alive agents leave the tile they stand on (pass 1: the old cell, later: the new one)
                if (!((alive >> a) & 1u)) continue;
                leave(pos[a], pass == 1 ? cellinfo[cellof(pos[a])] : info_new[a]);
                slot &= ~(1u << a);
            }
            LLE_UNROLL
            for (int a = 0; a < A_; ++a) pre_enter(a, np[a], info_new[a]);
            bool died = false;
            LLE_UNROLL
            for (int a = 0; a < A_; ++a) {
                const uint32_t code = enter(a, np[a], info_new[a]);
                if (pass == 1) ev[a] = code;
                else if (code != EV_NONE) ev[a] |= (pass > 63u ? 63u : pass) << 2;  // passes >= 2 can only emit deaths
                n_died += code == EV_DIED;
                n_gem += code == EV_GEM;
                n_exit += code == EV_EXIT;
                died = died || code == EV_DIED;
            }
            if (pass == 1) {
                LLE_UNROLL
                for (int a = 0; a < A_; ++a) pos[a] = np[a];
            }
            if (!died) break;
        }
        n_arrived += n_exit;
        n_deads += n_died;
        done = (n_arrived == (uint32_t)A_ || n_deads > 0) ? 1u : 0u;
    }
    // World::compute_available_actions (world.rs:343-363) of agent a as a 5-bit mask indexed by Action value
    LLE_HD uint32_t available(int a) const {
        uint32_t nbr = 0;
        if (((alive >> a) & 1u) && !((arrived >> a) & 1u)) {
            nbr = (cellinfo[cellof(pos[a])] >> 3) & 15u;
            LLE_UNROLL
            for (int o = 0; o < A_; ++o) {
                if (o == a || !((slot >> o) & 1u)) continue;  // Tile::is_occupied (tile.rs:97-99)
                const int d = (int)pos[o] - (int)pos[a];
                if (d == -256) nbr &= ~1u;
                else if (d == 256) nbr &= ~2u;
                else if (d == 1) nbr &= ~4u;
                else if (d == -1) nbr &= ~8u;
            }
        }
        return 16u | nbr;  // Stay is always listed
    }
    // LLE.available_actions with walkable_lasers == False (env.py:153-163): drop every listed action (STAY included) whose
    // target holds an on, listed (world.rs:159-172) laser of another colour
    LLE_HD uint32_t available_no_walk(int a, uint32_t mask) {
        uint32_t out = 0;
        LLE_UNROLL
        for (int k = 0; k < 5; ++k) {
            if (!((mask >> k) & 1u)) continue;
            const uint32_t c = cellof(pos[a] + (uint32_t)act_delta(k));
            bool blocked = false;
            if (cellinfo[c] & (1u << 24)) {
                const LleCellBeams cb = cellbeams[c];
                LLE_UNROLL
                for (int n = 0; n < 4; ++n) {
                    const uint32_t e = cb.e[n];
                    if (e != LLE_NO_BEAM && be_listed(e) && be_colour(e) != a && ((on_word(be_b(e)) >> be_k(e)) & 1u)) blocked = true;
                }
            }
            if (!blocked) out |= 1u << k;
        }
        return out;
    }
    // reward_strategy.py:58-75 (SingleObjective; the death override is dead code, :71-72) / :90-109 (MultiObjective)
    LLE_HD void reward(bool paid, int R, uint32_t n_gem, uint32_t n_exit, uint32_t n_died, float (&rw)[4]) const {
        const bool all_in = n_arrived == (uint32_t)A_;
        if (R == 1) {
            rw[0] = paid ? (float)n_gem + (float)n_exit - (float)n_died + (all_in ? 1.0f : 0.0f) : 0.0f;
            rw[1] = rw[2] = rw[3] = 0.0f;
        } else {
            rw[0] = (paid && !n_died) ? (float)n_gem : 0.0f;
            rw[1] = (paid && !n_died) ? (float)n_exit : 0.0f;
            rw[2] = paid ? -(float)n_died : 0.0f;
            rw[3] = (paid && !n_died && all_in) ? 1.0f : 0.0f;
        }
    }

    // ---- layered observation (observations.py:254-266) of this world in `sub`, a ZERO-FILLED block of obs_stride floats.
    // Write order of the reference: the static layers (walls, voids, exits = 1, sources = -1; :216-237), the laser cells whose
    // beam bit is on and the gems that are NOT collected (= 1; :256-263), then the agents (:264-265).  HW = H*W.
    LLE_HD void render(float* sub, int HW) {
        for (int k = 0; k < n_static; ++k) {
            const LlePatch pe = list[k];
            sub[pe.idx] = (float)pe.stat;
        }
        for (int k = n_static; k < n_static + n_patch; ++k) {
            const LlePatch pe = list[k];
            const uint32_t w = rec(pe.src == 0xFF ? L.w_gems : L.w_on + pe.src);
            if ((((w >> pe.bit) & 1u) != 0) != (pe.src == 0xFF)) sub[pe.idx] = 1.0f;
        }
        LLE_UNROLL
        for (int a = 0; a < A_; ++a) sub[a * HW + (int)(pos[a] >> 8) * W + (int)(pos[a] & 0xFFu)] = 1.0f;
    }
    // ---- partial observation (observations.py:312-369) of this world in `sub`, a ZERO-FILLED block; H = rows of the map
    LLE_HD void render_partial(float* sub, int sz, int H) {
        for (int a = 0; a < A_; ++a)
            for (int r = 0; r < sz * sz; ++r)
                partial_cell_task(sub, A_, a, r, sz, pos[a], H, W, cellinfo, cellbeams, L.w_gems, L.w_on, [&](int word) { return rec(word); });
        for (int a = 0; a < A_; ++a)
            for (int a2 = 0; a2 < A_; ++a2) partial_agent_task(sub, A_, a, a2, sz, pos[a], pos[a2]);
    }
};

}  // namespace lle
