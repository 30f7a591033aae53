// TEST HARNESS ONLY — never loaded by the product.
//
// Compiles the __host__ __device__ logic of lle_b200/csrc (step_core.cuh, vec_kernels.cuh) and the host
// map compiler with g++ and drives them exactly like lle_fused_kernel does (work units of 32 worlds,
// two tile buffers per "warp" with un-patch / patch, chunked tiles), lane by lane on the CPU, so that
// the device logic can be checked bit-for-bit against the oracle in this GPU-less container.
// The real parity tests (-m gpu) run the CUDA kernel itself.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../lle_b200/csrc/map_compiler.hpp"
#include "../../lle_b200/csrc/vec_kernels.cuh"

using namespace lle;

namespace {

struct Shim {
    std::vector<CompiledMap> maps;
    std::vector<const uint8_t*> blobs;
    std::vector<int32_t> map_of_env;
    int64_t N = 0, N_pad = 0;
    int A = 0, G = 0, NBmax = 0, C = 0, H = 0, W = 0, S = 0, R = 1, max_beam_len = 0, bucket = 0;
    LleStateLayout L;
    int64_t obs_stride = 0;
    int E = 1, n_chunks = 1, chunk_floats = 0, tile_floats = 0;
    std::vector<uint32_t> words;
    std::vector<float> obs, state, reward;
    std::vector<uint8_t> avail, done, events, err;
    std::vector<int8_t> actions;
    uint64_t seed = 0, env_id_base = 0, t = 0;
    int auto_reset = 1, lle_semantics = 1, walkable = 1;
    // persistent "warp" state (kept across launches here to stress the un-patch path)
    std::vector<float> tiles;
    std::vector<uint32_t> applied;
    std::vector<int32_t> tags;
    int buf = 0;
    std::string error;

    KParams params() {
        KParams p;
        std::memset(&p, 0, sizeof p);
        p.blobs = blobs.data();
        p.map_of_env = maps.size() > 1 ? map_of_env.data() : nullptr;
        p.words = words.data();
        p.L = L;
        p.N = N; p.N_pad = N_pad;
        p.A = A; p.G = G; p.NBmax = NBmax; p.C = C; p.H = H; p.W = W; p.S = S; p.R = R; p.HW = H * W;
        p.obs = obs.data(); p.obs_stride = obs_stride;
        p.state = state.data(); p.avail = avail.data(); p.reward = reward.data(); p.done = done.data();
        p.events = events.data(); p.actions = actions.data(); p.err = err.data();
        p.seed = seed; p.env_id_base = env_id_base; p.t = t;
        p.auto_reset = auto_reset; p.lle_semantics = lle_semantics; p.walkable = walkable; p.write_obs = 1;
        p.E = E; p.n_chunks = n_chunks; p.chunk_floats = chunk_floats; p.tile_floats = tile_floats;
        return p;
    }

    template <int AMAX, int NBMAX>
    void run(const KParams& p) {
        using D = Desc<AMAX, NBMAX>;
        if (applied.empty()) {
            applied.assign((size_t)2 * E * D::WORDS, 0);
            tags.assign((size_t)2 * E + 2, -1);
            tiles.assign((size_t)2 * tile_floats, 0.f);
        }
        std::vector<uint32_t> descs((size_t)D::WORDS * 32);
        for (int64_t unit = 0; unit < N_pad / 32; ++unit) {
            int map_ids[32];
            for (int lane = 0; lane < 32; ++lane) {
                int64_t env = unit * 32 + lane;
                map_ids[lane] = p.map_of_env ? p.map_of_env[env] : 0;
                MapView mv = MapView::make(p.blobs[map_ids[lane]]);
                Env<AMAX, NBMAX> e;
                unit_logic(p, env, mv, e);
                desc_write(e, descs.data() + lane, 32);
            }
            const int tiles_per_unit = n_chunks > 1 ? 32 : 32 / E;
            for (int chunk = 0; chunk < n_chunks; ++chunk) {
                const int lo = chunk * chunk_floats;
                const int hi = std::min(lo + chunk_floats, (int)obs_stride);
                for (int tix = 0; tix < tiles_per_unit; ++tix) {
                    float* tile = tiles.data() + (size_t)buf * tile_floats;
                    const int n_sub = n_chunks > 1 ? 1 : E;
                    for (int s = 0; s < n_sub; ++s) {
                        const int l = n_chunks > 1 ? tix : tix * E + s;
                        const int mid = map_ids[l];
                        const uint8_t* blob = p.blobs[mid];
                        float* sub = tile + (size_t)s * obs_stride;
                        uint32_t* old = applied.data() + ((size_t)buf * E + s) * D::WORDS;
                        const uint32_t* cur = descs.data() + l;
                        const bool same = tags[buf * E + s] == mid && tags[2 * E + buf] == chunk;
                        for (int lane = 0; lane < 32; ++lane) {
                            if (!same) tile_rebuild(sub, blob, lo, hi, lane);
                            else tile_unpatch(sub, blob, old, cur, 32, D::PW, A, H * W, W, lo, hi, lane);
                        }
                        for (int lane = 0; lane < 32; ++lane) tile_patch(sub, blob, cur, 32, D::PW, A, H * W, W, lo, hi, lane);
                        for (int w = 0; w < D::WORDS; ++w) old[w] = cur[w * 32];
                        tags[buf * E + s] = mid;
                    }
                    tags[2 * E + buf] = chunk;
                    const int64_t first_env = unit * 32 + (n_chunks > 1 ? tix : tix * E);
                    float* dst = p.obs + first_env * obs_stride + lo;
                    const size_t bytes = (size_t)((n_chunks > 1 ? (hi - lo) : E * (int)obs_stride) * 4);
                    std::memcpy(dst, tile, bytes);
                    buf ^= 1;
                }
            }
        }
    }

    void launch(const KParams& p) {
        switch (bucket) {
            case 0: run<4, 4>(p); break;
            case 1: run<8, 8>(p); break;
            case 2: run<8, 16>(p); break;
            default: run<16, 16>(p); break;
        }
    }
};

int pow2_floor(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }

}  // namespace

extern "C" {

const char* shim_error(void* h) { return ((Shim*)h)->error.c_str(); }

// tile_override_floats: 0 = the product's default tiling; otherwise force chunked tiles of that many floats
void* shim_create(const char** texts, int n_maps, const int* map_of_env, long n_envs, int reward_dim, int walkable,
                  int auto_reset, int lle_semantics, uint64_t seed, uint64_t env_id_base, int tile_override_floats,
                  char* errbuf, int errlen) {
    auto s = std::make_unique<Shim>();
    try {
        for (int k = 0; k < n_maps; ++k) s->maps.push_back(compile_map(texts[k]));
    } catch (const MapError& e) {
        std::snprintf(errbuf, errlen, "%d:%s", e.status, e.what());
        return nullptr;
    }
    for (auto& m : s->maps) s->blobs.push_back(m.blob.data());
    const CompiledMap& m0 = s->maps[0];
    s->A = m0.A; s->G = m0.G; s->C = m0.C; s->H = m0.H; s->W = m0.W; s->R = reward_dim; s->S = 3 * m0.A + m0.G;
    for (auto& m : s->maps) {
        if (m.A != s->A || m.G != s->G || m.H != s->H || m.W != s->W) {
            std::snprintf(errbuf, errlen, "202:shape mismatch");
            return nullptr;
        }
        s->NBmax = std::max(s->NBmax, m.NB);
        s->max_beam_len = std::max(s->max_beam_len, m.max_beam_len);
    }
    s->bucket = -1;
    for (int b = 0; b < kNumBuckets; ++b)
        if (s->A <= kBuckets[b].amax && s->NBmax <= kBuckets[b].nbmax) { s->bucket = b; break; }
    if (s->bucket < 0) {
        std::snprintf(errbuf, errlen, "21:bucket");
        return nullptr;
    }
    s->N = n_envs;
    s->N_pad = (n_envs + 31) / 32 * 32;
    s->map_of_env.assign((size_t)s->N_pad, 0);
    for (long e = 0; e < s->N_pad; ++e) s->map_of_env[(size_t)e] = map_of_env ? map_of_env[std::min(e, n_envs - 1)] : 0;
    s->L = lle_state_layout(s->A, s->G, s->NBmax, s->max_beam_len);
    s->obs_stride = ((int64_t)s->C * s->H * s->W + 3) / 4 * 4;
    const int64_t stride = s->obs_stride;
    if (tile_override_floats > 0) {
        s->E = 1;
        s->chunk_floats = tile_override_floats / 4 * 4;
        s->n_chunks = (int)((stride + s->chunk_floats - 1) / s->chunk_floats);
        s->tile_floats = s->chunk_floats;
        if (s->n_chunks == 1) { s->chunk_floats = (int)stride; s->tile_floats = (int)stride; }
    } else if (stride <= 6144) {
        s->n_chunks = 1;
        s->E = std::max(1, std::min(32, pow2_floor((int)std::max<int64_t>(1, 2048 / stride))));
        s->chunk_floats = (int)stride;
        s->tile_floats = (int)(s->E * stride);
    } else {
        s->E = 1;
        s->chunk_floats = 3072;
        s->n_chunks = (int)((stride + s->chunk_floats - 1) / s->chunk_floats);
        s->tile_floats = s->chunk_floats;
    }
    size_t Np = (size_t)s->N_pad;
    s->words.assign((size_t)s->L.n_words * Np, 0);
    s->obs.assign((size_t)stride * Np, -7.f);
    s->state.assign((size_t)s->S * Np, 0); s->avail.assign((size_t)s->A * 5 * Np, 0); s->reward.assign((size_t)s->R * Np, 0);
    s->done.assign(Np, 0); s->events.assign((size_t)s->A * Np, 0); s->actions.assign((size_t)s->A * Np, 0); s->err.assign(Np, 0);
    s->seed = seed; s->env_id_base = env_id_base;
    s->auto_reset = auto_reset; s->lle_semantics = lle_semantics; s->walkable = walkable;
    KParams p = s->params();
    p.mode = MODE_RESET;
    s->launch(p);
    return s.release();
}
void shim_free(void* h) { delete (Shim*)h; }

// out: N, A, G, C, H, W, R, S, NBmax, obs_stride, E, n_chunks
void shim_dims(void* h, long* out) {
    Shim& s = *(Shim*)h;
    long d[12] = {(long)s.N, s.A, s.G, s.C, s.H, s.W, s.R, s.S, s.NBmax, (long)s.obs_stride, s.E, s.n_chunks};
    std::memcpy(out, d, sizeof d);
}
// obs, state, avail, reward, done, events, actions, err
void shim_buffers(void* h, void** out) {
    Shim& s = *(Shim*)h;
    out[0] = s.obs.data(); out[1] = s.state.data(); out[2] = s.avail.data(); out[3] = s.reward.data();
    out[4] = s.done.data(); out[5] = s.events.data(); out[6] = s.actions.data(); out[7] = s.err.data();
}
void shim_reset(void* h, const uint8_t* mask) {
    Shim& s = *(Shim*)h;
    KParams p = s.params();
    p.mode = MODE_RESET;
    p.reset_mask = mask;
    s.launch(p);
}
void shim_step(void* h, const int8_t* actions) {
    Shim& s = *(Shim*)h;
    KParams p = s.params();
    p.mode = MODE_STEP;
    p.actions_in = actions;
    s.launch(p);
    s.t++;
}
void shim_set_state(void* h, const int32_t* pos, const uint8_t* gems, const uint8_t* alive) {
    Shim& s = *(Shim*)h;
    KParams p = s.params();
    p.mode = MODE_SET_STATE;
    p.ss_pos = pos; p.ss_gems = gems; p.ss_alive = alive;
    s.launch(p);
}
void shim_set_step_count(void* h, uint64_t t) { ((Shim*)h)->t = t; }

// raw record, same layout as lle_vec_export_raw
void shim_export_raw(void* h, int16_t* pos, uint8_t* alive, uint8_t* arrived, uint8_t* slot, uint64_t* beam_on,
                     uint64_t* collected, uint8_t* counters) {
    Shim& s = *(Shim*)h;
    const LleStateLayout& L = s.L;
    for (int64_t env = 0; env < s.N; ++env) {
        auto ld = [&](int w) { return s.words[(size_t)w * s.N_pad + env]; };
        uint32_t al, ar, sl, na, nd, dn;
        if (!L.wide_flags) {
            uint32_t f = ld(L.w_flags);
            al = f & 0xFF; ar = (f >> 8) & 0xFF; sl = (f >> 16) & 0xFF; na = (f >> 24) & 0xF; nd = (f >> 28) & 7; dn = f >> 31;
        } else {
            al = ld(L.w_flags); ar = ld(L.w_flags + 1); sl = ld(L.w_flags + 2);
            uint32_t m = ld(L.w_flags + 3);
            na = m & 0xFF; nd = (m >> 8) & 0xFF; dn = (m >> 16) & 1;
        }
        for (int a = 0; a < s.A; ++a) {
            uint32_t w = ld(a >> 1);
            uint32_t pp = (a & 1) ? (w >> 16) : (w & 0xFFFF);
            pos[(env * s.A + a) * 2] = (int16_t)(pp >> 8); pos[(env * s.A + a) * 2 + 1] = (int16_t)(pp & 0xFF);
            alive[env * s.A + a] = (al >> a) & 1; arrived[env * s.A + a] = (ar >> a) & 1; slot[env * s.A + a] = (sl >> a) & 1;
        }
        uint64_t c = 0;
        if (L.gem_words >= 1) c = ld(L.w_gems);
        if (L.gem_words == 2) c |= (uint64_t)ld(L.w_gems + 1) << 32;
        collected[env] = c;
        for (int b = 0; b < s.NBmax; ++b) {
            uint64_t v = ld(L.w_on + b * L.on_words);
            if (L.on_words == 2) v |= (uint64_t)ld(L.w_on + b * 2 + 1) << 32;
            beam_on[env * std::max(s.NBmax, 1) + b] = v;
        }
        counters[env * 3] = (uint8_t)na; counters[env * 3 + 1] = (uint8_t)nd; counters[env * 3 + 2] = (uint8_t)dn;
    }
}

// map compiler only (no worlds): status code, message in errbuf
int shim_compile(const char* text, char* errbuf, int errlen) {
    try {
        compile_map(text);
        return 0;
    } catch (const MapError& e) {
        std::snprintf(errbuf, errlen, "%s", e.what());
        return e.status;
    }
}
}
