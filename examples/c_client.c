/* A host program that uses the batched World through the C ABI alone (include/lle_b200.h): no Python, no torch, no CUDA
 * headers.  It is what a Rust / C++ host (the reference's src/core + src/bindings side) would do through FFI:
 *   parse a map -> create N worlds on the device -> reset -> step with host actions / device-sampled actions -> read
 *   reward and done on the host -> recolour a laser source -> step again.
 * Build:  gcc -O2 -Iinclude examples/c_client.c -Llle_b200/_native -llle_b200 -Wl,-rpath,$PWD/lle_b200/_native -o c_client
 * Prints one summary line per phase; exit code 0 on success.  tests/test_c_client.py runs it on the GPU box. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lle_b200.h"

#define CHECK(call)                                                             \
    do {                                                                        \
        int rc_ = (call);                                                       \
        if (rc_ != LLE_OK) {                                                    \
            fprintf(stderr, "%s failed: %d %s\n", #call, rc_, lle_last_error()); \
            return 1;                                                           \
        }                                                                       \
    } while (0)

int main(int argc, char** argv) {
    const int64_t n_envs = argc > 1 ? atoll(argv[1]) : 4096;
    const int steps = argc > 2 ? atoi(argv[2]) : 200;
    lle_map* map = NULL;
    CHECK(lle_map_level(6, &map));
    lle_map_info info;
    CHECK(lle_map_get_info(map, &info));
    printf("%s: level 6 is %dx%d, %d agents, %d gems, %d sources, %d channels\n", lle_version(), info.height, info.width,
           info.n_agents, info.n_gems, info.n_sources, info.n_channels);

    lle_vec_options opts;
    lle_vec_default_options(&opts);
    opts.seed = 7;
    lle_vec* vec = NULL;
    const lle_map* maps[1] = {map};
    CHECK(lle_vec_create(maps, 1, NULL, n_envs, &opts, &vec));
    lle_vec_buffers buf;
    CHECK(lle_vec_get_buffers(vec, &buf));
    const int A = buf.n_agents;

    int8_t* actions = (int8_t*)malloc((size_t)n_envs * A);
    float* reward = (float*)malloc((size_t)n_envs * buf.reward_dim * sizeof(float));
    uint8_t* done = (uint8_t*)malloc((size_t)n_envs);
    long episodes = 0, invalid = 0;
    double total_reward = 0.0;

    /* phase 1: device-sampled (Philox) actions, reward / done copied back by lle_vec_step_host */
    for (int t = 0; t < steps; ++t) {
        CHECK(lle_vec_step_host(vec, NULL, reward, done, NULL));
        for (int64_t e = 0; e < n_envs; ++e) {
            episodes += done[e];
            total_reward += reward[e];
        }
    }
    printf("sampled actions : %d steps x %lld envs, %ld episodes finished, reward sum %.1f\n", steps, (long long)n_envs, episodes,
           total_reward);
    if (episodes == 0) return 2;

    /* phase 2: host-supplied actions; everybody stays (always available), nothing can happen */
    CHECK(lle_vec_reset(vec, NULL, NULL));
    memset(actions, 4 /* Action::Stay */, (size_t)n_envs * A);
    for (int t = 0; t < 3; ++t) {
        CHECK(lle_vec_step_host(vec, actions, reward, done, NULL));
        for (int64_t e = 0; e < n_envs; ++e)
            if (reward[e] != 0.0f || done[e]) ++invalid;
    }
    printf("supplied actions: 3 x STAY, %ld unexpected transitions\n", invalid);
    if (invalid) return 3;

    /* phase 3: pipelined host stepping, four steps in flight (device-sampled actions) */
    CHECK(lle_vec_reset(vec, NULL, NULL));
    /* the pipelined calls need pinned host buffers: the library hands them out (no CUDA runtime in this program) */
    float* ring_reward = NULL;
    uint8_t* ring_done = NULL;
    CHECK(lle_host_alloc(4 * (size_t)n_envs * buf.reward_dim * sizeof(float), (void**)&ring_reward));
    CHECK(lle_host_alloc(4 * (size_t)n_envs, (void**)&ring_done));
    long pipelined_episodes = 0;
    for (int t = 0; t < steps + 4; ++t) {
        if (t >= 4) {
            int32_t left = 0;
            CHECK(lle_vec_pipeline_wait(vec, &left));
            const uint8_t* d = ring_done + (size_t)((t - 4) % 4) * n_envs;
            for (int64_t e = 0; e < n_envs; ++e) pipelined_episodes += d[e];
        }
        if (t < steps)
            CHECK(lle_vec_pipeline_submit(vec, NULL, ring_reward + (size_t)(t % 4) * n_envs * buf.reward_dim,
                                          ring_done + (size_t)(t % 4) * n_envs, NULL));
    }
    printf("pipelined       : %d steps, %ld episodes finished\n", steps, pipelined_episodes);
    if (pipelined_episodes == 0) return 4;

    /* phase 4: a source mutator (PyLaserSource.disable) and the exit setter, then more steps */
    CHECK(lle_vec_set_source(vec, 0, 1, -1, 0, NULL));
    const int32_t exits[8] = {11, 0, 11, 1, 11, 2, 11, 3};
    CHECK(lle_vec_set_exits(vec, 0, exits, 4, NULL));
    CHECK(lle_vec_reset(vec, NULL, NULL));
    for (int t = 0; t < 20; ++t) CHECK(lle_vec_step_host(vec, NULL, reward, done, NULL));
    uint64_t launches = 0;
    CHECK(lle_vec_launch_count(vec, &launches));
    printf("mutators        : source 1 disabled, exits moved; %llu kernel launches in total\n", (unsigned long long)launches);

    free(actions); free(reward); free(done);
    CHECK(lle_host_free(ring_reward));
    CHECK(lle_host_free(ring_done));
    CHECK(lle_vec_destroy(vec));
    lle_map_free(map);

    /* phase 5: layouts generated on the device (lle.generate(5, 5, 2).lasers(2)), read back, compiled and stepped */
    lle_gen_options gopts;
    lle_gen_default_options(&gopts);
    gopts.n_lasers = 2;
    lle_gen* gen = NULL;
    enum { ATTEMPTS = 4096, KEEP = 64 };
    CHECK(lle_gen_create(&gopts, 0, ATTEMPTS, &gen));
    CHECK(lle_gen_run(gen, NULL, 0, ATTEMPTS, 1, LLE_GEN_WALKABLE, NULL)); /* attempt i = _try_generate(seed = i) */
    static uint8_t cells[ATTEMPTS * 25], status[ATTEMPTS];
    CHECK(lle_gen_fetch(gen, 0, ATTEMPTS, cells, status, NULL, NULL, NULL));
    lle_map* gmaps[KEEP];
    int kept = 0, accepted = 0;
    for (int i = 0; i < ATTEMPTS; ++i) {
        accepted += status[i];
        if (status[i] && kept < KEEP) {
            char text[256];
            size_t len = 0;
            CHECK(lle_gen_cells_to_text(cells + i * 25, 5, 5, text, sizeof(text), &len));
            CHECK(lle_map_parse(text, len, &gmaps[kept]));
            ++kept;
        }
    }
    if (kept < KEEP) return 5;
    int32_t* map_of_env = (int32_t*)malloc(sizeof(int32_t) * KEEP * 16);
    for (int e = 0; e < KEEP * 16; ++e) map_of_env[e] = e / 16;
    lle_vec* gvec = NULL;
    CHECK(lle_vec_create((const lle_map* const*)gmaps, KEEP, map_of_env, KEEP * 16, &opts, &gvec));
    float* greward = (float*)malloc(sizeof(float) * KEEP * 16);
    uint8_t* gdone = (uint8_t*)malloc(KEEP * 16);
    long gepisodes = 0;
    CHECK(lle_vec_reset(gvec, NULL, NULL));
    for (int t = 0; t < 50; ++t) {
        CHECK(lle_vec_step_host(gvec, NULL, greward, gdone, NULL));
        for (int e = 0; e < KEEP * 16; ++e) gepisodes += gdone[e];
    }
    printf("generated maps  : %d of %d attempts accepted, %d maps x 16 envs stepped 50 times, %ld episodes finished\n", accepted, ATTEMPTS, kept,
           gepisodes);
    free(map_of_env); free(greward); free(gdone);
    CHECK(lle_vec_destroy(gvec));
    for (int k = 0; k < kept; ++k) lle_map_free(gmaps[k]);
    CHECK(lle_gen_destroy(gen));
    printf("ok\n");
    return 0;
}
