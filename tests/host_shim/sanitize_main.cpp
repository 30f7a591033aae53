// TEST INFRASTRUCTURE: the product's HOST code that touches untrusted input (map texts: lle_b200/csrc/map_compiler.cpp,
// toml_config.cpp / toml_lite.hpp) and the per-world core of the tiny-map kernel (tiny_core.cuh, through tests/host_shim/tiny_host.cpp)
// under AddressSanitizer + UndefinedBehaviorSanitizer.  compute-sanitizer is not available on this GPU pool; this covers the
// host side.  Reads map texts from the files given on the command line (one map per file), compiles each for every observation
// type (errors are expected for the malformed ones and must be clean exceptions), and steps the tiny-eligible ones.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../lle_b200/csrc/map_compiler.hpp"

extern "C" {
void* tiny_host_create(const char** texts, int n_maps, const int* map_of_env, long n_envs, int reward_dim, int walkable, int auto_reset,
                       int lle_semantics, uint64_t seed, uint64_t env_id_base, int E, int n_warps, int obs_kind, int obs_param, char* err,
                       int errlen);
void tiny_host_free(void* h);
void tiny_host_step(void* h, const int8_t* actions_in);
}

int main(int argc, char** argv) {
    int compiled = 0, rejected = 0, stepped = 0;
    for (int k = 1; k < argc; ++k) {
        std::ifstream f(argv[k]);
        std::stringstream ss;
        ss << f.rdbuf();
        const std::string text = ss.str();
        const int kinds[][2] = {{LLE_OBS_LAYERED, 0}, {LLE_OBS_LAYERED, 2}, {LLE_OBS_PARTIAL, 3}, {LLE_OBS_PARTIAL, 7}, {LLE_OBS_PERSPECTIVE, 0}, {LLE_OBS_STATE, 1}};
        for (const auto& kp : kinds) {
            lle::ObsSpec spec;
            spec.kind = kp[0];
            spec.param = kp[1];
            try {
                const lle::CompiledMap m = lle::compile_map(text, spec);
                compiled += m.blob.size() > 0;
            } catch (const std::exception&) {
                ++rejected;
            }
        }
        for (const int obs : {0, 3}) {
            char err[256];
            const char* texts[1] = {text.c_str()};
            void* h = tiny_host_create(texts, 1, nullptr, 70, obs ? 1 : 4, obs ? 0 : 1, 1, 1, 11u + (unsigned)k, 5, obs ? 4 : 8, 2, obs ? LLE_OBS_PARTIAL : LLE_OBS_LAYERED, obs, err,
                                       (int)sizeof err);
            if (!h) continue;  // not a tiny record, or a malformed map
            for (int t = 0; t < 40; ++t) tiny_host_step(h, nullptr);
            std::vector<int8_t> acts(70 * 8, (int8_t)(k % 5));
            tiny_host_step(h, acts.data());  // supplied actions, many of them unavailable
            tiny_host_free(h);
            ++stepped;
        }
    }
    std::printf("{\"compiled\": %d, \"rejected\": %d, \"stepped\": %d}\n", compiled, rejected, stepped);
    return 0;
}
