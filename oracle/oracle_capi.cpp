// ORACLE — TEST INFRASTRUCTURE ONLY (see lle_oracle.hpp header).
//
// C entry points (ctypes) over the oracle engine:
//   lleo_world_*  one `World`           (mirrors the reference's PyWorld surface, pyworld.rs:144-626)
//   lleo_env_*    one `LLE`             (python/lle/env/env.py)
//   lleo_vec_*    N independent `LLE`s  stepped in lockstep with the SAME output layout as the
//                 device buffers of lle_b200, plus the Philox action rule of SURVEY §8(d), so a
//                 parity test is a plain array comparison.  Also the CPU baseline timed by bench.py.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>

#include "lle_oracle.hpp"

using namespace lle_oracle;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const ParseError& e) {
        return fail((int)e.kind, e.what());
    } catch (const RuntimeWorldError& e) {
        static const char* names[] = {"InvalidAction", "InvalidNumberOfGems", "InvalidNumberOfAgents",
                                      "InvalidAgentPosition", "OutOfWorldPosition", "InvalidNumberOfActions",
                                      "InvalidWorldState", "TileNotWalkable", "Panic"};
        char buf[96];
        std::snprintf(buf, sizeof buf, " [%ld,%ld,%ld]", e.a, e.b, e.c);
        return fail((int)e.kind, std::string(names[(int)e.kind - 101]) + ": " + e.what() + buf);
    } catch (const std::out_of_range& e) {
        return fail(201, e.what());  // numpy IndexError
    } catch (const std::invalid_argument& e) {
        return fail(202, e.what());  // python ValueError
    } catch (const std::exception& e) {
        return fail(299, e.what());
    }
}

using lle_oracle::Philox;  // restated in lle_oracle.hpp (the start sampler of World::reset needs it too)

// Action rule: counter = (env_id, step_lo, agent/4, step_hi), key = (seed_lo, seed_hi);
// agent a takes word a%4; action = k-th set bit of the availability mask in Action.value order,
// k = mulhi(word, popcount(mask)).
inline uint8_t sample_action(uint64_t seed, uint32_t env_id, uint64_t step, uint32_t agent, uint32_t mask5) {
    uint32_t ctr[4] = {env_id, (uint32_t)step, agent / 4, (uint32_t)(step >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    Philox::run(ctr, key, out);
    uint32_t n = (uint32_t)__builtin_popcount(mask5 & 31u);
    uint32_t k = (uint32_t)(((uint64_t)out[agent % 4] * n) >> 32);
    for (uint32_t b = 0; b < 5; ++b) {
        if (mask5 >> b & 1) {
            if (k == 0) return (uint8_t)b;
            --k;
        }
    }
    return 4;
}

struct WorldHandle {
    World w;
    std::unique_ptr<Layered> layered;  // built lazily so channel-index errors surface at observe()
};

int dir_code(Direction d) { return (int)d; }

// event byte shared with the device ABI (include/lle_b200.h): bits 0-1 = pass-1 event
// (0 none, 1 exit, 2 gem, 3 died); bits 2-7 = pass (>=2) in which the agent died, 0 = none.
void encode_events(const std::vector<WorldEvent>& events, const std::vector<int>& pass, uint8_t* out, size_t A) {
    std::memset(out, 0, A);
    for (size_t k = 0; k < events.size(); ++k) {
        const auto& e = events[k];
        if (pass[k] == 1) {
            uint8_t code = e.type == EventType::AgentExit ? 1 : e.type == EventType::GemCollected ? 2 : 3;
            out[e.agent_id] |= code;
        } else {
            out[e.agent_id] |= (uint8_t)(pass[k] << 2);
        }
    }
}

struct Vec {
    std::vector<std::unique_ptr<Env>> envs;
    std::vector<int> map_of_env;
    size_t N = 0, A = 0, G = 0, C = 0, H = 0, W = 0, R = 1, S = 0, NB = 0;
    size_t OF = 0;  // floats of one env's observation block (ObsGen::floats)
    bool auto_reset = false;
    uint64_t seed = 0, env_id_base = 0, t = 0;
    uint32_t reset_epoch = 1;  // explicit resets so far (the construction reset is the first)
    void arm_rng(size_t e, uint64_t step) {  // counter words of the start sampler for a reset of env e at step `step`
        World& w = envs[e]->world;
        w.rng_seed = seed; w.rng_env = (uint32_t)(env_id_base + e); w.rng_t = (uint32_t)step; w.rng_epoch = reset_epoch;
    }
    std::vector<float> obs, state, reward, extras;  // extras: LaserSubgoal flags [N, A, JE]
    size_t JE = 0;
    std::vector<uint8_t> avail, done, events, err;
    std::vector<int8_t> actions;
    // raw engine state for white-box parity checks
    std::vector<int16_t> pos;       // [N,A,2]
    std::vector<uint8_t> alive, arrived, slot;  // [N,A]
    std::vector<uint64_t> beam_on;  // [N,NB] bit k = beam[k]
    std::vector<uint64_t> collected;  // [N]

    void export_env(size_t e) {
        Env& env = *envs[e];
        env.observe(&obs[e * OF]);
        env.state(&state[e * S]);
        env.available_actions(&avail[e * A * 5]);
        if (JE) env.compute_extras(&extras[e * A * JE]);
        const World& w = env.world;
        for (size_t a = 0; a < A; ++a) {
            pos[(e * A + a) * 2] = (int16_t)w.agents_positions[a].i;
            pos[(e * A + a) * 2 + 1] = (int16_t)w.agents_positions[a].j;
            alive[e * A + a] = w.agents[a].is_alive();
            arrived[e * A + a] = w.agents[a].has_arrived();
            const Tile* t = w.at(w.agents_positions[a]);
            auto occ = t->agent();
            slot[e * A + a] = occ.has_value() && *occ == a;
        }
        for (size_t b = 0; b < w.laser_source_positions.size(); ++b) {
            auto beam = w.source_beam(b);
            uint64_t m = 0;
            for (size_t k = 0; k < beam->beam.size() && k < 64; ++k) m |= (uint64_t)beam->beam[k] << k;
            beam_on[e * NB + b] = m;
        }
        uint64_t cm = 0;
        auto gems = w.gems();
        for (size_t g = 0; g < gems.size() && g < 64; ++g) cm |= (uint64_t)gems[g]->collected << g;
        collected[e] = cm;
    }

    void step_env(size_t e, const int8_t* given) {
        Env& env = *envs[e];
        err[e] = 0;
        std::vector<Action> acts(A);
        for (size_t a = 0; a < A; ++a) {
            uint8_t v;
            if (given) {
                v = (uint8_t)given[e * A + a];
            } else {
                uint32_t mask = 0;
                for (Action x : env.world.available_actions[a]) mask |= 1u << (uint32_t)x;
                v = sample_action(seed, (uint32_t)(env_id_base + e), t, (uint32_t)a, mask);
            }
            actions[e * A + a] = (int8_t)v;
            acts[a] = (Action)v;
        }
        std::fill(&reward[e * R], &reward[e * R] + R, 0.0f);
        std::memset(&events[e * A], 0, A);
        bool bad = false;
        for (size_t a = 0; a < A; ++a)
            if ((uint8_t)actions[e * A + a] > 4) bad = true;
        if (env.done) {
            err[e] = 2;  // env.py:166-167 "Cannot step in a done environment"
        } else if (bad) {
            err[e] = 1;
        } else {
            try {
                auto ev = env.step(acts, &reward[e * R]);
                encode_events(ev, env.world.last_event_pass, &events[e * A], A);
            } catch (const RuntimeWorldError& ex) {
                if (ex.kind != RuntimeErrorKind::InvalidAction) throw;
                err[e] = 1;  // world.rs:444-453: the world is left untouched
            }
        }
        done[e] = env.done;
        if (env.done && auto_reset && err[e] == 0) {
            arm_rng(e, t);
            env.reset();
        }
        export_env(e);
    }
};

}  // namespace

extern "C" {

const char* lleo_last_error() { return g_last_error.c_str(); }

void lleo_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { Philox::run(ctr, key, out); }
int lleo_sample_action(uint64_t seed, uint32_t env_id, uint64_t step, uint32_t agent, uint32_t mask5) {
    return sample_action(seed, env_id, step, agent, mask5);
}

// parse_v1 (parser_v1.rs:132-175) of a world string, written out in the config-text form (lle_oracle.hpp:
// parse_config_text) — oracle/toml_config.py completes a v2 document with it (toml_config.rs:36-95).
int lleo_v1_config_text(const char* text, char* out, long cap) {
    return guarded([&] {
        WorldConfig cfg = parse_v1(text);
        std::string os;  // std::to_string, not iostream formatting (locale-independent)
        auto num = [&](size_t v) { os += " " + std::to_string(v); };
        auto positions = [&](const char* name, const std::vector<Position>& v) {
            os += name;
            num(v.size());
            for (const auto& p : v) { num(p.i); num(p.j); }
            os += "\n";
        };
        os += "%LLE-CONFIG\nsize";
        num(cfg.height); num(cfg.width);
        os += "\n";
        positions("gems", cfg.gems);
        positions("voids", cfg.voids);
        positions("exits", cfg.exits);
        positions("walls", cfg.walls);
        os += "agents"; num(cfg.random_starts.size()); os += "\n";
        for (const auto& s : cfg.random_starts) positions("starts", s);
        os += "lasers"; num(cfg.lasers.size()); os += "\n";
        for (const auto& l : cfg.lasers) {
            os += "laser";
            num(l.first.i); num(l.first.j); num(l.second.agent_id); num((size_t)l.second.direction); num(l.second.laser_id);
            os += "\n";
        }
        const std::string& str = os;
        if ((long)str.size() + 1 > cap) throw std::invalid_argument("config text buffer too small");
        std::memcpy(out, str.c_str(), str.size() + 1);
    });
}

// ---------------------------------------------------------------- World
void* lleo_world_new(const char* text, int* status) {
    WorldHandle* h = nullptr;
    *status = guarded([&] {
        auto tmp = std::make_unique<WorldHandle>();
        tmp->w = parse_world(text);
        h = tmp.release();
    });
    return h;
}
void lleo_world_free(void* p) { delete (WorldHandle*)p; }

// out: H, W, A, G, n_sources, n_exits, n_walls, n_voids
void lleo_world_dims(void* p, int* out) {
    const World& w = ((WorldHandle*)p)->w;
    out[0] = (int)w.height; out[1] = (int)w.width; out[2] = (int)w.n_agents(); out[3] = (int)w.n_gems();
    out[4] = (int)w.laser_source_positions.size(); out[5] = (int)w.exits.size();
    out[6] = (int)w.wall_positions.size(); out[7] = (int)w.void_positions.size();
}
// counter words of the start sampler, kept like the device's N = 1 facade keeps them: the construction reset is explicit
// reset number 1, every later reset() adds one, every step() call (failed ones included) advances the step count
int lleo_world_reset(void* p) {
    World& w = ((WorldHandle*)p)->w;
    w.rng_epoch++;
    return guarded([&] { w.reset(); });
}
void lleo_world_seed(void* p, uint64_t seed) { ((WorldHandle*)p)->w.rng_seed = seed; }
void lleo_env_seed(void* p, uint64_t seed) { ((Env*)p)->world.rng_seed = seed; }
// out: per agent the candidate count, then the (i, j) pairs, agent after agent; returns the number of ints written
int lleo_world_random_starts(void* p, long* out, int cap) {
    const World& w = ((WorldHandle*)p)->w;
    int n = 0;
    for (const auto& c : w.random_start_positions) {
        if (n < cap) out[n] = (long)c.size();
        ++n;
        for (const auto& pos : c) {
            if (n + 1 < cap) { out[n] = (long)pos.i; out[n + 1] = (long)pos.j; }
            n += 2;
        }
    }
    return n;
}

// events_out: pairs (type, agent_id); passes_out (optional): pass index per event
int lleo_world_step(void* p, const uint8_t* actions, int n, int* events_out, int* passes_out, int* n_events) {
    World& w = ((WorldHandle*)p)->w;
    *n_events = 0;
    w.rng_t++;  // after the call: the step count the next reset would see
    return guarded([&] {
        std::vector<Action> acts;
        for (int a = 0; a < n; ++a) acts.push_back((Action)actions[a]);
        auto ev = w.step(acts);
        for (size_t k = 0; k < ev.size(); ++k) {
            events_out[2 * k] = (int)ev[k].type;
            events_out[2 * k + 1] = (int)ev[k].agent_id;
            if (passes_out) passes_out[k] = w.last_event_pass[k];
        }
        *n_events = (int)ev.size();
    });
}
// mask[a*5 + action.value]; order[a*5 + k] = k-th listed action (world.rs:349-351 order), -1 padded
void lleo_world_available(void* p, uint8_t* mask, int8_t* order) {
    const World& w = ((WorldHandle*)p)->w;
    size_t A = w.n_agents();
    std::memset(mask, 0, A * 5);
    if (order) std::memset(order, -1, A * 5);
    for (size_t a = 0; a < A; ++a) {
        size_t k = 0;
        for (Action x : w.available_actions[a]) {
            mask[a * 5 + (size_t)x] = 1;
            if (order) order[a * 5 + k++] = (int8_t)x;
        }
    }
}
void lleo_world_get_state(void* p, int* pos, uint8_t* gems, uint8_t* alive) {
    WorldState s = ((WorldHandle*)p)->w.get_state();
    for (size_t a = 0; a < s.agents_positions.size(); ++a) {
        pos[2 * a] = (int)s.agents_positions[a].i;
        pos[2 * a + 1] = (int)s.agents_positions[a].j;
        alive[a] = s.agents_alive[a];
    }
    for (size_t g = 0; g < s.gems_collected.size(); ++g) gems[g] = s.gems_collected[g];
}
int lleo_world_set_state(void* p, const long* pos, int n_agents, const uint8_t* gems, int n_gems,
                         const uint8_t* alive, int* events_out, int* n_events) {
    World& w = ((WorldHandle*)p)->w;
    *n_events = 0;
    return guarded([&] {
        WorldState s;
        for (int a = 0; a < n_agents; ++a) {
            s.agents_positions.push_back(Position{(size_t)pos[2 * a], (size_t)pos[2 * a + 1]});
            s.agents_alive.push_back(alive[a] != 0);
        }
        for (int g = 0; g < n_gems; ++g) s.gems_collected.push_back(gems[g] != 0);
        auto ev = w.set_state(s);
        for (size_t k = 0; k < ev.size(); ++k) {
            events_out[2 * k] = (int)ev[k].type;
            events_out[2 * k + 1] = (int)ev[k].agent_id;
        }
        *n_events = (int)ev.size();
    });
}
void lleo_world_agents(void* p, uint8_t* dead, uint8_t* arrived) {
    const World& w = ((WorldHandle*)p)->w;
    for (size_t a = 0; a < w.n_agents(); ++a) {
        dead[a] = w.agents[a].is_dead();
        arrived[a] = w.agents[a].has_arrived();
    }
}
// per laser: i, j, laser_id, agent_id, direction, is_on, is_enabled  -> returns count (<= max)
int lleo_world_lasers(void* p, int* out, int max) {
    std::vector<LaserView> ls;
    try {
        ls = ((WorldHandle*)p)->w.lasers();
    } catch (const RuntimeWorldError&) {
        return -1;  // `unreachable!()` in World::lasers (world.rs:159-172): an inconsistent world; the Python side raises
    }
    int n = 0;
    for (const auto& l : ls) {
        if (n >= max) break;
        int* o = out + 7 * n++;
        o[0] = (int)l.pos.i; o[1] = (int)l.pos.j; o[2] = (int)l.laser_id; o[3] = (int)l.agent_id;
        o[4] = dir_code(l.direction); o[5] = l.is_on; o[6] = l.is_enabled;
    }
    return (int)ls.size();
}
// per source: i, j, agent_id, direction, is_enabled, laser_id, beam_len
int lleo_world_sources(void* p, int* out, int max) {
    const World& w = ((WorldHandle*)p)->w;
    int n = 0;
    for (size_t s = 0; s < w.laser_source_positions.size() && n < max; ++s, ++n) {
        auto b = w.source_beam(s);
        int* o = out + 7 * n;
        o[0] = (int)w.laser_source_positions[s].i; o[1] = (int)w.laser_source_positions[s].j;
        o[2] = (int)b->agent_id; o[3] = dir_code(b->direction); o[4] = b->enabled; o[5] = (int)b->laser_id;
        o[6] = (int)b->beam.size();
    }
    return (int)w.laser_source_positions.size();
}
// kind: 0 walls, 1 voids, 2 exits, 3 gems, 4 starts (actual), 5 laser tiles, 6 agents_positions
int lleo_world_positions(void* p, int kind, int* out, int max) {
    const World& w = ((WorldHandle*)p)->w;
    const std::vector<Position>* v = nullptr;
    switch (kind) {
        case 0: v = &w.wall_positions; break;
        case 1: v = &w.void_positions; break;
        case 2: v = &w.exits; break;
        case 3: v = &w.gems_positions; break;
        case 4: v = &w.start_positions; break;
        case 5: v = &w.lasers_positions; break;
        default: v = &w.agents_positions; break;
    }
    for (size_t k = 0; k < v->size() && (int)k < max; ++k) {
        out[2 * k] = (int)(*v)[k].i;
        out[2 * k + 1] = (int)(*v)[k].j;
    }
    return (int)v->size();
}
// World::gems (world.rs:129-139) reaches `unreachable!()` when a gem position holds no gem tile (a TOML document that lists a
// gem on a wall / exit / void cell): returns 1 there instead of the flags, and the Python side raises
int lleo_world_gem_flags(void* p, uint8_t* out) {
    auto gems = ((WorldHandle*)p)->w.gems();
    for (size_t g = 0; g < gems.size(); ++g) {
        if (!gems[g]) return 1;
        out[g] = gems[g]->collected;
    }
    return 0;
}
// PyGem::collect (src/bindings/tiles/pygem.rs:51-65): only a top-level Gem tile; returns 0, or 1 when the tile is not a gem
int lleo_world_gem_collect(void* p, long i, long j) {
    World& w = ((WorldHandle*)p)->w;
    if (i < 0 || j < 0 || (size_t)i >= w.height || (size_t)j >= w.width) return 2;
    Tile& t = w.grid[(size_t)i][(size_t)j];
    if (t.kind != Tile::Gem) return 1;
    t.collected = true;  // Gem::collect (gem.rs:17-19)
    return 0;
}
// PyGem::agent / PyLaser::agent (pygem.rs:67-76, pylaser.rs:73-81): the agent standing on the tile, -1 if none
int lleo_world_tile_agent(void* p, long i, long j) {
    World& w = ((WorldHandle*)p)->w;
    if (i < 0 || j < 0 || (size_t)i >= w.height || (size_t)j >= w.width) return -1;
    auto a = w.grid[(size_t)i][(size_t)j].agent();
    return a.has_value() ? (int)*a : -1;
}
int lleo_world_n_gems_collected(void* p) { return (int)((WorldHandle*)p)->w.n_gems_collected(); }
void lleo_world_source_set_enabled(void* p, int idx, int enabled) {
    auto b = ((WorldHandle*)p)->w.source_beam((size_t)idx);
    if (enabled) b->enable(); else b->disable();
}
void lleo_world_source_set_agent_id(void* p, int idx, int agent_id) {
    ((WorldHandle*)p)->w.source_beam((size_t)idx)->agent_id = (size_t)agent_id;
}
int lleo_world_set_exits(void* p, const long* ij, int n) {
    WorldHandle* h = (WorldHandle*)p;
    return guarded([&] {
        std::vector<Position> pos;
        for (int k = 0; k < n; ++k) pos.push_back(Position{(size_t)ij[2 * k], (size_t)ij[2 * k + 1]});
        h->w.set_exit_positions(pos);
    });
}
void lleo_world_beam_bits(void* p, int idx, uint8_t* out) {
    auto b = ((WorldHandle*)p)->w.source_beam((size_t)idx);
    for (size_t k = 0; k < b->beam.size(); ++k) out[k] = b->beam[k];
}
void lleo_world_occupied(void* p, uint8_t* out) {
    const World& w = ((WorldHandle*)p)->w;
    for (size_t i = 0; i < w.height; ++i)
        for (size_t j = 0; j < w.width; ++j) out[i * w.width + j] = w.grid[i][j].is_occupied();
}
int lleo_world_observe_layered(void* p, float* out) {
    WorldHandle* h = (WorldHandle*)p;
    return guarded([&] {
        Layered l(h->w);  // python builds the generator from the live world (observations.py:199-214)
        l.observe(h->w, out);
    });
}
void lleo_world_state_array(void* p, float* out) { state_as_array(((WorldHandle*)p)->w.get_state(), out); }

// ---------------------------------------------------------------- Env (LLE)
void* lleo_env_new(const char* text, int multi_objective, int walkable_lasers, int* status) {
    Env* e = nullptr;
    *status = guarded([&] { e = new Env(parse_world(text), multi_objective != 0, walkable_lasers != 0); });
    return e;
}
void lleo_env_free(void* p) { delete (Env*)p; }
void* lleo_env_world(void* p) { return nullptr; (void)p; }
void lleo_env_dims(void* p, int* out) {
    const Env& e = *(Env*)p;
    out[0] = (int)e.world.height; out[1] = (int)e.world.width; out[2] = (int)e.world.n_agents();
    out[3] = (int)e.world.n_gems(); out[4] = (int)(2 * e.world.n_agents() + 4); out[5] = (int)e.reward_dim();
}
// Builder.obs_type (builder.py:42-49): kind 0 layered (param = padding), 1 partial (param = size), 2 perspective,
// 3 state (param 1 = normalised).  out6: block_shape (see ObsGen::block_shape).
int lleo_env_set_obs(void* p, int kind, int param, long* out6) {
    Env& e = *(Env*)p;
    return guarded([&] {
        e.set_obs((ObsKind)kind, param);
        e.obs.block_shape(out6);
    });
}
long lleo_env_obs_floats(void* p) { return (long)((Env*)p)->obs.floats(); }
// Builder.state_type (builder.py:51-58): any observation generator as the state; out6 like lleo_env_set_obs
int lleo_env_set_state_type(void* p, int kind, int param, long* out6) {
    Env& e = *(Env*)p;
    return guarded([&] {
        e.set_state_type((ObsKind)kind, param);
        e.state_gen->block_shape(out6);
    });
}
long lleo_env_state_obs_floats(void* p) { return (long)((Env*)p)->state_gen->floats(); }
int lleo_env_state_observation(void* p, float* out) { return guarded([&] { ((Env*)p)->state_observation(out); }); }
// a generator built from the live world, as the reference tests do (`PartialGenerator(world, 3).observe()`)
int lleo_world_observe(void* p, int kind, int param, float* out, long cap, long* out6) {
    WorldHandle* h = (WorldHandle*)p;
    return guarded([&] {
        ObsGen g(h->w, (ObsKind)kind, param);
        g.block_shape(out6);
        if (out) {
            if ((long)g.floats() > cap) throw std::invalid_argument("observation buffer too small");
            g.observe(h->w, out);
        }
    });
}
int lleo_env_reset(void* p) {
    Env& e = *(Env*)p;
    e.world.rng_epoch++;
    return guarded([&] { e.reset(); });
}
int lleo_env_step(void* p, const uint8_t* actions, int n, float* reward, uint8_t* done, int* events_out,
                  int* n_events) {
    Env& e = *(Env*)p;
    *n_events = 0;
    e.world.rng_t++;
    return guarded([&] {
        std::vector<Action> acts;
        for (int a = 0; a < n; ++a) acts.push_back((Action)actions[a]);
        auto ev = e.step(acts, reward);
        *done = e.done;
        for (size_t k = 0; k < ev.size(); ++k) {
            events_out[2 * k] = (int)ev[k].type;
            events_out[2 * k + 1] = (int)ev[k].agent_id;
        }
        *n_events = (int)ev.size();
    });
}
int lleo_env_set_state(void* p, const long* pos, int n_agents, const uint8_t* gems, int n_gems,
                       const uint8_t* alive) {
    Env& e = *(Env*)p;
    return guarded([&] {
        WorldState s;
        for (int a = 0; a < n_agents; ++a) {
            s.agents_positions.push_back(Position{(size_t)pos[2 * a], (size_t)pos[2 * a + 1]});
            s.agents_alive.push_back(alive[a] != 0);
        }
        for (int g = 0; g < n_gems; ++g) s.gems_collected.push_back(gems[g] != 0);
        e.set_state(s);
    });
}
int lleo_env_enable_extras(void* p, int n, const int* src) {
    Env& e = *(Env*)p;
    return guarded([&] {
        std::vector<size_t> s;
        if (n < 0) for (size_t k = 0; k < e.world.laser_source_positions.size(); ++k) s.push_back(k);
        else for (int k = 0; k < n; ++k) s.push_back((size_t)src[k]);
        e.enable_extras(s);
    });
}
int lleo_env_enable_pbrs(void* p, double gamma, double value, int n, const int* src) {
    Env& e = *(Env*)p;
    return guarded([&] {
        std::vector<size_t> s;
        if (n < 0) for (size_t k = 0; k < e.world.laser_source_positions.size(); ++k) s.push_back(k);
        else for (int k = 0; k < n; ++k) s.push_back((size_t)src[k]);
        e.enable_pbrs(gamma, value, s);
    });
}
int lleo_env_extras_dim(void* p) { return ((Env*)p)->has_extras ? (int)((Env*)p)->extras.pos_to_reward.size() : 0; }
void lleo_env_extras(void* p, float* out) { ((Env*)p)->compute_extras(out); }
int lleo_env_reward_dim(void* p) { return (int)((Env*)p)->reward_dim(); }
int lleo_env_done(void* p) { return ((Env*)p)->done; }
int lleo_env_n_arrived(void* p) { return (int)((Env*)p)->n_arrived; }
void lleo_env_available(void* p, uint8_t* out) { ((Env*)p)->available_actions(out); }
int lleo_env_observe(void* p, float* out) { return guarded([&] { ((Env*)p)->observe(out); }); }
void lleo_env_state(void* p, float* out) { ((Env*)p)->state(out); }
// Step.info of LLE.step (env.py:174-188): World::n_gems_collected and, per agent, Agent.is_dead / has_arrived
int lleo_env_info(void* p, uint8_t* dead, uint8_t* arrived) {
    const World& w = ((Env*)p)->world;
    for (size_t a = 0; a < w.n_agents(); ++a) {
        dead[a] = w.agents[a].is_dead();
        arrived[a] = w.agents[a].has_arrived();
    }
    return (int)w.n_gems_collected();
}

// ---------------------------------------------------------------- Vec (N envs, device-layout outputs)
void* lleo_vec_new(const char** texts, int n_maps, const int* map_of_env, int n_envs, int multi_objective,
                   int walkable_lasers, int auto_reset, uint64_t seed, uint64_t env_id_base, int* status) {
    Vec* v = nullptr;
    *status = guarded([&] {
        auto tmp = std::make_unique<Vec>();
        tmp->N = (size_t)n_envs;
        tmp->auto_reset = auto_reset != 0;
        tmp->seed = seed;
        tmp->env_id_base = env_id_base;
        for (int e = 0; e < n_envs; ++e) {
            int m = map_of_env ? map_of_env[e] : 0;
            if (m < 0 || m >= n_maps) throw std::invalid_argument("map index out of range");
            tmp->map_of_env.push_back(m);
            tmp->envs.push_back(std::make_unique<Env>(parse_world(texts[m]), multi_objective != 0,
                                                      walkable_lasers != 0));
        }
        const Env& e0 = *tmp->envs[0];
        tmp->A = e0.world.n_agents(); tmp->G = e0.world.n_gems(); tmp->C = 2 * tmp->A + 4;
        tmp->H = e0.world.height; tmp->W = e0.world.width; tmp->R = e0.reward_dim();
        tmp->S = 3 * tmp->A + tmp->G;
        for (const auto& e : tmp->envs) {
            if (e->world.n_agents() != tmp->A || e->world.n_gems() != tmp->G || e->world.height != tmp->H ||
                e->world.width != tmp->W)
                throw std::invalid_argument("all maps of a vec must share (H, W, n_agents, n_gems)");
            tmp->NB = std::max(tmp->NB, e->world.laser_source_positions.size());
        }
        size_t N = tmp->N, A = tmp->A;
        tmp->OF = e0.obs.floats();
        tmp->obs.assign(N * tmp->OF, 0.f);
        tmp->state.assign(N * tmp->S, 0.f);
        tmp->reward.assign(N * tmp->R, 0.f);
        tmp->avail.assign(N * A * 5, 0);
        tmp->done.assign(N, 0);
        tmp->events.assign(N * A, 0);
        tmp->err.assign(N, 0);
        tmp->actions.assign(N * A, 4);
        tmp->pos.assign(N * A * 2, 0);
        tmp->alive.assign(N * A, 0);
        tmp->arrived.assign(N * A, 0);
        tmp->slot.assign(N * A, 0);
        tmp->beam_on.assign(N * std::max<size_t>(tmp->NB, 1), 0);
        tmp->collected.assign(N, 0);
        for (size_t e = 0; e < N; ++e) {
            tmp->arm_rng(e, 0);
            tmp->envs[e]->reset(false);  // the construction reset is World::new's, not LLE.reset
            tmp->export_env(e);
        }
        v = tmp.release();
    });
    return v;
}
// Builder.add_extras("laser_subgoal") / Builder.pbrs(...) for every env of the vec (python/lle/env/builder.py:77-150).
// Source lists are indices into World::sources() order; n < 0 means "all sources".  Resets every env afterwards.
int lleo_vec_configure(void* p, int n_extras, const int* extras_src, int pbrs, double gamma, double value, int n_pbrs,
                       const int* pbrs_src) {
    Vec& v = *(Vec*)p;
    return guarded([&] {
        auto pick = [](const World& w, int n, const int* src) {
            std::vector<size_t> out;
            if (n < 0) for (size_t k = 0; k < w.laser_source_positions.size(); ++k) out.push_back(k);
            else for (int k = 0; k < n; ++k) out.push_back((size_t)src[k]);
            return out;
        };
        size_t je = 0;
        for (auto& e : v.envs) {
            if (n_extras != 0) {
                auto src = pick(e->world, n_extras, extras_src);
                e->enable_extras(src);
                je = std::max(je, src.size());
            }
            if (pbrs) e->enable_pbrs(gamma, value, pick(e->world, n_pbrs, pbrs_src));
        }
        v.JE = je;
        v.extras.assign(v.N * v.A * std::max<size_t>(je, 1), 0.f);
        v.R = v.envs[0]->reward_dim();
        v.reward.assign(v.N * v.R, 0.f);
        for (size_t e = 0; e < v.N; ++e) {
            v.arm_rng(e, v.t);
            v.envs[e]->reset(false);
            v.export_env(e);
        }
    });
}
// LLE(randomize_lasers=True) for every env of the vec (takes effect from the next LLE-level reset on)
void lleo_vec_set_randomize_lasers(void* p, int enabled) {
    for (auto& e : ((Vec*)p)->envs) e->randomize_lasers = enabled != 0;
}
// the env's world through the World getters above: a WorldHandle view is not needed, only the two listings
int lleo_env_sources(void* p, int* out, int max) {
    const World& w = ((Env*)p)->world;
    int n = 0;
    for (size_t s = 0; s < w.laser_source_positions.size() && n < max; ++s, ++n) {
        auto b = w.source_beam(s);
        int* o = out + 7 * n;
        o[0] = (int)w.laser_source_positions[s].i; o[1] = (int)w.laser_source_positions[s].j;
        o[2] = (int)b->agent_id; o[3] = dir_code(b->direction); o[4] = b->enabled; o[5] = (int)b->laser_id;
        o[6] = (int)b->beam.size();
    }
    return (int)w.laser_source_positions.size();
}
int lleo_env_lasers(void* p, int* out, int max) {
    auto ls = ((Env*)p)->world.lasers();
    int n = 0;
    for (const auto& l : ls) {
        if (n >= max) break;
        int* o = out + 7 * n++;
        o[0] = (int)l.pos.i; o[1] = (int)l.pos.j; o[2] = (int)l.laser_id; o[3] = (int)l.agent_id;
        o[4] = dir_code(l.direction); o[5] = l.is_on; o[6] = l.is_enabled;
    }
    return (int)ls.size();
}
void lleo_env_set_randomize_lasers(void* p, int enabled) { ((Env*)p)->randomize_lasers = enabled != 0; }
// Observation type of every env of the vec (see lleo_env_set_obs); re-exports all envs.  The obs buffer is reallocated:
// fetch the pointers again with lleo_vec_buffers.
int lleo_vec_set_obs(void* p, int kind, int param, long* out6) {
    Vec& v = *(Vec*)p;
    return guarded([&] {
        for (auto& e : v.envs) e->set_obs((ObsKind)kind, param);
        v.envs[0]->obs.block_shape(out6);
        v.OF = v.envs[0]->obs.floats();
        v.obs.assign(v.N * v.OF, 0.f);
        for (size_t e = 0; e < v.N; ++e) v.export_env(e);
    });
}
// LaserBeam::set_agent_id / enable / disable (laser.rs:69-84) on source `idx` of every env whose map is `map_index`
// (agent_id < 0 / enabled < 0: keep).  Outputs are re-exported; call lleo_vec_reset to refresh the cached static
// observation layers the way LLE.reset does (ObservationGenerator.reset, observations.py:128-137).
int lleo_vec_set_source(void* p, int map_index, int idx, int agent_id, int enabled) {
    Vec& v = *(Vec*)p;
    return guarded([&] {
        for (size_t e = 0; e < v.N; ++e) {
            if (v.map_of_env[e] != map_index) continue;
            auto b = v.envs[e]->world.source_beam((size_t)idx);
            if (agent_id >= 0) b->agent_id = (size_t)agent_id;
            if (enabled == 0 && b->enabled) b->disable();
            else if (enabled > 0 && !b->enabled) b->enable();
            v.envs[e]->obs.setup(v.envs[e]->world);
            v.export_env(e);
        }
    });
}
// World::set_exit_positions (world.rs:195-234) on every env whose map is `map_index`; static observation layers refreshed.
int lleo_vec_set_exits(void* p, int map_index, const long* ij, int n) {
    Vec& v = *(Vec*)p;
    return guarded([&] {
        std::vector<Position> pos;
        for (int k = 0; k < n; ++k) pos.push_back(Position{(size_t)ij[2 * k], (size_t)ij[2 * k + 1]});
        for (size_t e = 0; e < v.N; ++e) {
            if (v.map_of_env[e] != map_index) continue;
            v.envs[e]->world.set_exit_positions(pos);
            v.envs[e]->obs.setup(v.envs[e]->world);
            v.export_env(e);
        }
    });
}
void* lleo_vec_extras(void* p) { return ((Vec*)p)->extras.data(); }
long lleo_vec_extras_dim(void* p) { return (long)((Vec*)p)->JE; }
long lleo_vec_reward_dim(void* p) { return (long)((Vec*)p)->R; }
void* lleo_vec_reward(void* p) { return ((Vec*)p)->reward.data(); }
void lleo_vec_free(void* p) { delete (Vec*)p; }
// out: N, A, G, C, H, W, R, S, NB
void lleo_vec_dims(void* p, long* out) {
    Vec& v = *(Vec*)p;
    long d[9] = {(long)v.N, (long)v.A, (long)v.G, (long)v.C, (long)v.H, (long)v.W, (long)v.R, (long)v.S, (long)v.NB};
    std::memcpy(out, d, sizeof d);
}
// order: obs, state, avail, reward, done, events, actions, err, pos, alive, arrived, slot, beam_on, collected
void lleo_vec_buffers(void* p, void** out) {
    Vec& v = *(Vec*)p;
    out[0] = v.obs.data(); out[1] = v.state.data(); out[2] = v.avail.data(); out[3] = v.reward.data();
    out[4] = v.done.data(); out[5] = v.events.data(); out[6] = v.actions.data(); out[7] = v.err.data();
    out[8] = v.pos.data(); out[9] = v.alive.data(); out[10] = v.arrived.data(); out[11] = v.slot.data();
    out[12] = v.beam_on.data(); out[13] = v.collected.data();
}
int lleo_vec_reset_masked(void* p, const uint8_t* mask);
int lleo_vec_reset(void* p) { return lleo_vec_reset_masked(p, nullptr); }
// lle_vec_reset with a mask: only the envs whose byte is non-zero are reset (all when mask is null)
int lleo_vec_reset_masked(void* p, const uint8_t* mask) {
    Vec& v = *(Vec*)p;
    v.reset_epoch++;
    return guarded([&] {
        for (size_t e = 0; e < v.N; ++e) {
            if (mask && !mask[e]) continue;
            v.arm_rng(e, v.t);
            v.envs[e]->reset();
            v.done[e] = 0; v.err[e] = 0;
            std::fill(&v.reward[e * v.R], &v.reward[e * v.R] + v.R, 0.f);
            std::memset(&v.events[e * v.A], 0, v.A);
            std::memset(&v.actions[e * v.A], 4, v.A);  // no action has been taken in the new episode (device ABI: STAY)
            v.export_env(e);
        }
    });
}
// LLE.set_state on one env of the vec (events dropped); returns the oracle status code.
int lleo_vec_set_state_env(void* p, long e, const long* pos, int n_agents, const uint8_t* gems, int n_gems,
                           const uint8_t* alive) {
    Vec& v = *(Vec*)p;
    return guarded([&] {
        WorldState s;
        for (int a = 0; a < n_agents; ++a) {
            // negative coordinates wrap to huge usize values in the reference's (usize, usize) positions
            s.agents_positions.push_back(Position{(size_t)pos[2 * a], (size_t)pos[2 * a + 1]});
            s.agents_alive.push_back(alive[a] != 0);
        }
        for (int g = 0; g < n_gems; ++g) s.gems_collected.push_back(gems[g] != 0);
        v.envs[(size_t)e]->set_state(s);
    });
}
// Re-export every env's outputs (after out-of-band mutations such as lleo_vec_set_state_env).
void lleo_vec_refresh(void* p) {
    Vec& v = *(Vec*)p;
    for (size_t e = 0; e < v.N; ++e) {
        v.done[e] = v.envs[e]->done;
        v.export_env(e);
    }
}
uint64_t lleo_vec_step_count(void* p) { return ((Vec*)p)->t; }
void lleo_vec_set_step_count(void* p, uint64_t t) { ((Vec*)p)->t = t; }

// One lockstep step of all N envs.  actions == NULL => Philox sampling.  n_threads <= 0 => hardware
// concurrency.  Envs are range-partitioned over threads (one World is only ever touched by one thread).
int lleo_vec_step(void* p, const int8_t* actions, int n_threads) {
    Vec& v = *(Vec*)p;
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::min<size_t>((size_t)T, v.N);
    std::atomic<int> status{0};
    auto work = [&](size_t lo, size_t hi) {
        int s = guarded([&] {
            for (size_t e = lo; e < hi; ++e) v.step_env(e, actions);
        });
        if (s) status = s;
    };
    if (T <= 1) {
        work(0, v.N);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < T; ++k) th.emplace_back(work, v.N * k / T, v.N * (k + 1) / T);
        for (auto& x : th) x.join();
    }
    v.t += 1;
    return status;
}

// CPU baseline: each thread owns a contiguous env range and advances it `steps` times with Philox
// actions, auto-reset, and all outputs (layered obs, state, avail, reward, done, events) written each
// step, exactly as lleo_vec_step does.  Returns wall seconds of the timed region; *threads_used set.
double lleo_vec_rollout(void* p, int steps, int n_threads, int* threads_used) {
    Vec& v = *(Vec*)p;
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::min<size_t>((size_t)T, v.N);
    if (threads_used) *threads_used = T;
    uint64_t t0 = v.t;
    auto work = [&](size_t lo, size_t hi) {
        // each thread keeps its own step counter; per-env results depend only on (env, t)
        for (int s = 0; s < steps; ++s) {
            uint64_t t = t0 + (uint64_t)s;
            for (size_t e = lo; e < hi; ++e) {
                Env& env = *v.envs[e];
                std::vector<Action> acts(v.A);
                for (size_t a = 0; a < v.A; ++a) {
                    uint32_t mask = 0;
                    for (Action x : env.world.available_actions[a]) mask |= 1u << (uint32_t)x;
                    uint8_t act = sample_action(v.seed, (uint32_t)(v.env_id_base + e), t, (uint32_t)a, mask);
                    v.actions[e * v.A + a] = (int8_t)act;
                    acts[a] = (Action)act;
                }
                auto ev = env.step(acts, &v.reward[e * v.R]);
                encode_events(ev, env.world.last_event_pass, &v.events[e * v.A], v.A);
                v.done[e] = env.done;
                if (env.done) {
                    v.arm_rng(e, t);
                    env.reset();
                }
                env.observe(&v.obs[e * v.OF]);
                env.state(&v.state[e * v.S]);
                env.available_actions(&v.avail[e * v.A * 5]);
                if (v.JE) env.compute_extras(&v.extras[e * v.A * v.JE]);
            }
        }
    };
    auto start = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int k = 0; k < T; ++k) th.emplace_back(work, v.N * k / T, v.N * (k + 1) / T);
    for (auto& x : th) x.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    v.t += (uint64_t)steps;
    return secs;
}

}  // extern "C"
