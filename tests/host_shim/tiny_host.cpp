// TEST INFRASTRUCTURE: a host (g++) instantiation of the per-world core of the tiny-map step kernel
// (lle_b200/csrc/tiny_core.cuh is __host__ __device__), driven the way lle_tiny_step_kernel drives it — tickets of 32
// consecutive worlds, one world per "lane", E zero-filled sub-tiles per emulated warp — so that
// `-m "not gpu"` tests can compare the very code the kernel runs per thread with the oracle, bit for bit, without a GPU.
// The maps are compiled by the product's own host map compiler.  The product never loads this file.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../lle_b200/csrc/map_compiler.hpp"
#include "../../lle_b200/csrc/tiny_core.cuh"

namespace {

struct ArrayRec {
    uint32_t* base;
    uint32_t& operator()(int word) const { return base[word]; }
};
struct PitchRec {
    uint32_t* base;
    int pitch;
    uint32_t& operator()(int word) const { return base[word * pitch]; }
};

struct Warp {
    std::vector<float> tile;  // [E][ostr]
};

struct Shim {
    std::vector<lle::CompiledMap> maps;
    std::vector<int> map_of_env;
    int64_t N = 0, N_pad = 0;
    int A = 0, G = 0, H = 0, W = 0, S = 0, R = 1, E = 8;
    int obs_kind = LLE_OBS_LAYERED, obs_param = 0;
    LleStateLayout L;
    int64_t ostr = 0;
    int walkable = 1, auto_reset = 1, lle_semantics = 1;
    uint64_t seed = 0, env_id_base = 0, t = 0;
    std::vector<uint32_t> records;
    std::vector<float> obs, state, reward;
    std::vector<uint8_t> avail, done, events, err;
    std::vector<int8_t> actions;
    std::vector<Warp> warps;
    std::string error;
};

template <int A_>
void reset_all(Shim& s) {
    for (int64_t env = 0; env < s.N_pad; ++env) {
        lle::TinyWorld<A_, ArrayRec> w;
        w.rec = ArrayRec{s.records.data() + env * s.L.stride};
        w.L = lle::TinyLayout{s.L.w_flags, s.L.w_avail, s.L.w_gems, s.L.w_on, s.L.stride, s.L.gem_words != 0};
        w.W = s.W;
        w.bind(s.maps[(size_t)s.map_of_env[(size_t)env]].blob.data());
        w.reset();
        uint32_t cache = 0;
        for (int a = 0; a < A_; ++a) cache |= w.available(a) << (8 * a);
        w.pack(cache);
    }
}

// one step of every world, written like the body of lle_tiny_step_kernel (tiny_kernel.cuh)
template <int A_>
void step_all(Shim& s, const int8_t* actions_in) {
    const int stride = s.L.stride;
    const int64_t n_tickets = s.N_pad / 32;
    for (int64_t ticket = 0; ticket < n_tickets; ++ticket) {
        Warp& wp = s.warps[(size_t)(ticket % (int64_t)s.warps.size())];
        std::vector<uint32_t> srec((size_t)stride * 32);  // [stride][32] columns
        std::vector<lle::TinyWorld<A_, PitchRec>> lanes(32);
        int map_ids[32];
        for (int lane = 0; lane < 32; ++lane) {
            const int64_t env = ticket * 32 + lane;
            const bool real = env < s.N;
            auto& w = lanes[(size_t)lane];
            w.rec = PitchRec{srec.data() + lane, 32};
            w.L = lle::TinyLayout{s.L.w_flags, s.L.w_avail, s.L.w_gems, s.L.w_on, stride, s.L.gem_words != 0};
            w.W = s.W;
            for (int k = 0; k < stride; ++k) w.rec(k) = s.records[(size_t)(env * stride + k)];
            map_ids[lane] = s.map_of_env[(size_t)env];
            w.bind(s.maps[(size_t)map_ids[lane]].blob.data());
            w.unpack();
            const uint32_t av_cache = w.rec(s.L.w_avail);
            uint32_t act[A_], ev[A_];
            bool bad = false;
            uint32_t r[4] = {0, 0, 0, 0};
            if (!actions_in)
                lle::philox4x32_10((uint32_t)(s.env_id_base + (uint64_t)env), (uint32_t)s.t, 0u, (uint32_t)(s.t >> 32), (uint32_t)s.seed,
                                   (uint32_t)(s.seed >> 32), r);
            for (int a = 0; a < A_; ++a) {
                const uint32_t av = (av_cache >> (8 * a)) & 0xFFu;
                act[a] = 4u;
                if (actions_in) {
                    if (real) act[a] = (uint32_t)(uint8_t)actions_in[env * A_ + a];
                } else {
                    act[a] = lle::pick_action(r[a], av);
                }
                if (act[a] > 4u || !((av >> act[a]) & 1u)) bad = true;
                ev[a] = 0;
            }
            uint32_t err = lle::ERR_OK;
            if (s.lle_semantics && w.done) err = lle::ERR_DONE;
            else if (bad) err = lle::ERR_INVALID_ACTION;
            const bool paid = err == lle::ERR_OK;
            uint32_t n_gem = 0, n_exit = 0, n_died = 0;
            if (paid) w.step(act, ev, n_gem, n_exit, n_died);
            float rw[4];
            w.reward(paid, s.R, n_gem, n_exit, n_died, rw);
            if (real) {
                for (int k = 0; k < s.R; ++k) s.reward[(size_t)(env * s.R + k)] = rw[k];
                s.done[(size_t)env] = (uint8_t)w.done;
                s.err[(size_t)env] = (uint8_t)err;
                for (int a = 0; a < A_; ++a) {
                    s.events[(size_t)(env * A_ + a)] = (uint8_t)ev[a];
                    s.actions[(size_t)(env * A_ + a)] = (int8_t)act[a];
                }
            }
            if (s.auto_reset && w.done && err == lle::ERR_OK) w.reset();
            uint32_t cache = 0;
            for (int a = 0; a < A_; ++a) {
                uint32_t mask = w.available(a);
                cache |= mask << (8 * a);
                if (!s.walkable) mask = w.available_no_walk(a, mask);
                if (real)
                    for (int k = 0; k < 5; ++k) s.avail[(size_t)((env * A_ + a) * 5 + k)] = (uint8_t)((mask >> k) & 1u);
            }
            w.pack(cache);
            for (int k = 0; k < stride; ++k) s.records[(size_t)(env * stride + k)] = w.rec(k);
            if (real) {
                float* st = s.state.data() + env * s.S;
                for (int a = 0; a < A_; ++a) {
                    st[2 * a] = (float)(w.pos[a] >> 8);
                    st[2 * a + 1] = (float)(w.pos[a] & 0xFFu);
                    st[2 * A_ + s.G + a] = ((w.alive >> a) & 1u) ? 1.0f : 0.0f;
                }
                if (s.G) {
                    const uint32_t coll = w.rec(s.L.w_gems);
                    for (int g = 0; g < s.G; ++g) st[2 * A_ + g] = ((coll >> g) & 1u) ? 1.0f : 0.0f;
                }
            }
        }
        // observation: E lanes at a time patch their sub-tile, then the tile "leaves" (memcpy in the place of the bulk store)
        const int E = s.E;
        for (int r = 0; r < 32 / E; ++r) {
            for (int lane = r * E; lane < (r + 1) * E; ++lane) {
                auto& w = lanes[(size_t)lane];
                const int sidx = lane - r * E;
                float* sub = wp.tile.data() + (size_t)sidx * s.ostr;
                for (int64_t f = 0; f < s.ostr; ++f) sub[f] = 0.0f;
                if (s.obs_kind == LLE_OBS_PARTIAL) w.render_partial(sub, s.obs_param, s.H);
                else w.render(sub, s.H * s.W);
            }
            for (int sidx = 0; sidx < E; ++sidx) {
                const int64_t env = ticket * 32 + (int64_t)r * E + sidx;
                if (env < s.N) std::memcpy(s.obs.data() + env * s.ostr, wp.tile.data() + (size_t)sidx * s.ostr, (size_t)s.ostr * 4);
            }
        }
    }
    s.t++;
}

}  // namespace

extern "C" {

void* tiny_host_create(const char** texts, int n_maps, const int* map_of_env, long n_envs, int reward_dim, int walkable, int auto_reset,
                       int lle_semantics, uint64_t seed, uint64_t env_id_base, int E, int n_warps, int obs_kind, int obs_param, char* err,
                       int errlen) {
    auto s = std::make_unique<Shim>();
    auto fail = [&](const std::string& why) -> void* {
        if (err && errlen > 0) {
            std::strncpy(err, why.c_str(), (size_t)errlen - 1);
            err[errlen - 1] = 0;
        }
        return nullptr;
    };
    try {
        lle::ObsSpec spec;
        spec.kind = obs_kind;
        spec.param = obs_param;
        if (obs_kind != LLE_OBS_LAYERED && obs_kind != LLE_OBS_PARTIAL) return fail("the tiny path renders layered and partial observations");
        for (int k = 0; k < n_maps; ++k) s->maps.push_back(lle::compile_map(texts[k], spec));
    } catch (const std::exception& e) {
        return fail(e.what());
    }
    const auto& m0 = s->maps[0];
    s->A = m0.A; s->G = m0.G; s->H = m0.H; s->W = m0.W; s->S = 3 * m0.A + m0.G; s->R = reward_dim;
    int nb = 0, max_len = 0;
    for (const auto& m : s->maps) {
        if (m.A != s->A || m.G != s->G || m.H != s->H || m.W != s->W) return fail("maps differ in shape");
        nb = std::max(nb, m.NB);
        max_len = std::max(max_len, m.max_beam_len);
        if (m.header().random_starts) return fail("random starts are not on the tiny path");
    }
    s->L = lle_state_layout(s->A, s->G, nb, max_len);
    if (s->A > 4 || s->L.n_words > 8 || s->L.on_words != 1 || s->L.gem_words > 1) return fail("not a tiny record");
    if (E < 1 || E > 32 || (E & (E - 1))) return fail("E must be a power of two <= 32");
    s->E = E;
    s->obs_kind = obs_kind;
    s->obs_param = obs_param;
    s->N = n_envs;
    s->N_pad = (n_envs + 31) / 32 * 32;
    s->ostr = ((int64_t)m0.header().obs_floats + 3) / 4 * 4;
    s->walkable = walkable; s->auto_reset = auto_reset; s->lle_semantics = lle_semantics; s->seed = seed; s->env_id_base = env_id_base;
    s->map_of_env.assign((size_t)s->N_pad, 0);
    for (long e = 0; e < n_envs; ++e) s->map_of_env[(size_t)e] = map_of_env ? map_of_env[e] : 0;
    for (int64_t e = n_envs; e < s->N_pad; ++e) s->map_of_env[(size_t)e] = map_of_env ? map_of_env[n_envs - 1] : 0;
    s->records.assign((size_t)(s->N_pad * s->L.stride), 0);
    s->obs.assign((size_t)(s->N * s->ostr), 0.f);
    s->state.assign((size_t)(s->N * s->S), 0.f);
    s->reward.assign((size_t)(s->N * s->R), 0.f);
    s->avail.assign((size_t)(s->N * s->A * 5), 0);
    s->done.assign((size_t)s->N, 0);
    s->events.assign((size_t)(s->N * s->A), 0);
    s->err.assign((size_t)s->N, 0);
    s->actions.assign((size_t)(s->N * s->A), 0);
    s->warps.resize((size_t)std::max(1, n_warps));
    for (auto& w : s->warps) {
        w.tile.assign((size_t)(E * s->ostr), 0.f);
    }
    switch (s->A) {
        case 1: reset_all<1>(*s); break;
        case 2: reset_all<2>(*s); break;
        case 3: reset_all<3>(*s); break;
        default: reset_all<4>(*s); break;
    }
    return s.release();
}

void tiny_host_free(void* h) { delete (Shim*)h; }

void tiny_host_step(void* h, const int8_t* actions_in) {
    Shim& s = *(Shim*)h;
    switch (s.A) {
        case 1: step_all<1>(s, actions_in); break;
        case 2: step_all<2>(s, actions_in); break;
        case 3: step_all<3>(s, actions_in); break;
        default: step_all<4>(s, actions_in); break;
    }
}

// k: 0 obs f32, 1 state f32, 2 avail u8, 3 reward f32, 4 done u8, 5 events u8, 6 actions i8, 7 err u8
void* tiny_host_buffer(void* h, int k) {
    Shim& s = *(Shim*)h;
    switch (k) {
        case 0: return s.obs.data();
        case 1: return s.state.data();
        case 2: return s.avail.data();
        case 3: return s.reward.data();
        case 4: return s.done.data();
        case 5: return s.events.data();
        case 6: return s.actions.data();
        default: return s.err.data();
    }
}
void tiny_host_dims(void* h, long* out) {
    Shim& s = *(Shim*)h;
    out[0] = s.A; out[1] = s.G; out[2] = s.H; out[3] = s.W; out[4] = (long)s.ostr; out[5] = s.R; out[6] = s.maps[0].C; out[7] = s.S;
}

}  // extern "C"
