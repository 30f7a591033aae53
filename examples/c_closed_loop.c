/* Closed-loop end-to-end throughput of the batched World from a COMPILED host on the C ABI alone (include/lle_b200.h) — what
 * the reference's Rust side would see through FFI: every step takes its actions from (pinned) host memory and delivers reward
 * and done to host memory, and the host reads EVERY result byte of step t before it submits step t + 1 (an acting agent's
 * dependency).  The batch is cut into P sub-batches that are in flight together (the EnvPool pattern): while the host handles
 * one, the others step.
 *
 * The "policy" is a table lookup that keeps the rollout valid without a neural network: a device-sampled rollout is recorded
 * first (actions and done flags of K steps); in the timed loop the host compares the done flags it received with the recorded
 * ones (they match, since the env is deterministic) and only then hands over the recorded actions of the next step — had they
 * differed, everybody would STAY.
 *
 *   c_closed_loop <device> <n_envs> <steps> <parts> [<parts> ...]     one JSON line; exit code 0 on success
 * <parts> = P: P sub-batch vecs through lle_vec_pipeline_submit/_wait;  sP: ONE vec, one launch per step, P parts ordered on the
 * device through lle_vec_parts_* (the host still reads every result of a part before it releases the part's next actions).
 * Build: gcc -O2 -Iinclude examples/c_closed_loop.c -Llle_b200/_native -llle_b200 -Wl,-rpath,... -o c_closed_loop */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lle_b200.h"

#define CHECK(call)                                                             \
    do {                                                                        \
        int rc_ = (call);                                                       \
        if (rc_ != LLE_OK) {                                                    \
            fprintf(stderr, "%s failed: %d %s\n", #call, rc_, lle_last_error()); \
            return 1;                                                           \
        }                                                                       \
    } while (0)

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define MAX_PARTS 16
#define SEED 2026u
#define T0 10000000ull

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s <device> <n_envs> <steps> <parts>...\n", argv[0]);
        return 2;
    }
    const int device = atoi(argv[1]);
    const int64_t n_envs = atoll(argv[2]);
    const int K = atoi(argv[3]);
    lle_map* map = NULL;
    CHECK(lle_map_level(6, &map));
    const lle_map* maps[1] = {map};
    lle_vec_options opts;
    lle_vec_default_options(&opts);
    opts.device = device;
    opts.seed = SEED;

    /* 1. record a valid rollout on the device (sampled actions), K steps of the whole batch */
    lle_vec* full = NULL;
    CHECK(lle_vec_create(maps, 1, NULL, n_envs, &opts, &full));
    lle_vec_buffers buf;
    CHECK(lle_vec_get_buffers(full, &buf));
    const int A = buf.n_agents, R = buf.reward_dim;
    int8_t* rec_act = NULL;
    uint8_t* rec_done = NULL;
    CHECK(lle_host_alloc((size_t)K * n_envs * A, (void**)&rec_act));
    CHECK(lle_host_alloc((size_t)K * n_envs, (void**)&rec_done));
    CHECK(lle_vec_set_step_count(full, T0));
    for (int s = 0; s < K; ++s) {
        CHECK(lle_vec_step(full, NULL, NULL));
        CHECK(lle_vec_fetch(full, LLE_BUF_ACTIONS, 0, (size_t)n_envs * A, rec_act + (size_t)s * n_envs * A, NULL));
        CHECK(lle_vec_fetch(full, LLE_BUF_DONE, 0, (size_t)n_envs, rec_done + (size_t)s * n_envs, NULL));
    }
    CHECK(lle_vec_destroy(full));

    printf("{\"n_envs\": %lld, \"steps\": %d, \"agents\": %d, \"host\": \"compiled C on the C ABI\", \"parts\": {", (long long)n_envs, K, A);
    for (int arg = 4; arg < argc; ++arg) {
        if (argv[arg][0] == 's') {
            /* streamed mode: ONE vec, one launch per step of the whole batch, the dependency enforced per part on the device
             * (lle_vec_parts_*): the host waits for a part's step-s results, chooses the part's actions of step s+1, releases them */
            const int P = atoi(argv[arg] + 1);
            if (P < 1 || P > 1024) {
                fprintf(stderr, "bad part count %s\n", argv[arg]);
                return 2;
            }
            lle_vec* vec = NULL;
            int8_t* act = NULL;
            float* rw = NULL;
            uint8_t* dn = NULL;
            CHECK(lle_vec_create(maps, 1, NULL, n_envs, &opts, &vec));
            CHECK(lle_host_alloc((size_t)n_envs * A, (void**)&act));
            CHECK(lle_host_alloc((size_t)n_envs * R * sizeof(float), (void**)&rw));
            CHECK(lle_host_alloc((size_t)n_envs, (void**)&dn));
            double best = 1e30;
            long mismatches = 0;
            int32_t np = 0;
            for (int rep = 0; rep < 3; ++rep) {
                CHECK(lle_vec_reset(vec, NULL, NULL));
                CHECK(lle_vec_set_step_count(vec, T0));
                CHECK(lle_vec_fetch(vec, LLE_BUF_DONE, 0, 1, dn, NULL)); /* drains the stream */
                CHECK(lle_vec_parts_begin(vec, P, act, rw, dn, LLE_STREAM_NONE));
                CHECK(lle_vec_parts_count(vec, &np));
                mismatches = 0;
                double t_wait = 0, t_policy = 0, t_feed = 0, t_launch = 0;
                const double t0 = now_s();
                CHECK(lle_vec_parts_launch(vec)); /* step 0 */
                for (int h = 0; h < np; ++h) {
                    int64_t first, cnt;
                    CHECK(lle_vec_parts_range(vec, h, &first, &cnt));
                    memcpy(act + first * A, rec_act + first * A, (size_t)cnt * A);
                    CHECK(lle_vec_parts_feed(vec, h));
                }
                const int ahead = getenv("LLE_LOOP_AHEAD") ? atoi(getenv("LLE_LOOP_AHEAD")) : 2; /* step launches kept in flight (<= 4) */
                for (int k = 1; k < ahead && k < K; ++k) CHECK(lle_vec_parts_launch(vec)); /* they wait on the device for their actions */
                for (int s = 0; s < K; ++s) {
                    for (int h = 0; h < np; ++h) {
                        int64_t first, cnt;
                        CHECK(lle_vec_parts_range(vec, h, &first, &cnt));
                        const double c0 = now_s();
                        CHECK(lle_vec_parts_wait(vec, h)); /* results of step s of this part are in host memory */
                        const double c1 = now_s();
                        int ok = memcmp(dn + first, rec_done + (size_t)s * n_envs + first, (size_t)cnt) == 0; /* reads every byte */
                        if (!ok) ++mismatches;
                        double c2 = c1, c3 = c1;
                        if (s + 1 < K) {
                            if (ok) memcpy(act + first * A, rec_act + ((size_t)(s + 1) * n_envs + first) * A, (size_t)cnt * A);
                            else memset(act + first * A, 4, (size_t)cnt * A);
                            c2 = now_s();
                            CHECK(lle_vec_parts_feed(vec, h));
                            c3 = now_s();
                        }
                        t_wait += c1 - c0; t_policy += c2 - c1; t_feed += c3 - c2;
                    }
                    const double c4 = now_s();
                    if (s + ahead < K) CHECK(lle_vec_parts_launch(vec));
                    t_launch += now_s() - c4;
                }
                if (rep == 2) fprintf(stderr, "s%d host us per step: wait %.1f policy %.1f feed %.1f launch %.1f\n", P, 1e6 * t_wait / K, 1e6 * t_policy / K, 1e6 * t_feed / K, 1e6 * t_launch / K);
                CHECK(lle_vec_parts_end(vec));
                const double dt = now_s() - t0;
                if (rep > 0 && dt < best) best = dt;
            }
            long errs = 0;
            uint8_t* e = (uint8_t*)malloc((size_t)n_envs);
            CHECK(lle_vec_fetch(vec, LLE_BUF_ERR, 0, (size_t)n_envs, e, NULL));
            for (int64_t k = 0; k < n_envs; ++k) errs += e[k];
            CHECK(lle_vec_fetch(vec, LLE_BUF_DONE, 0, (size_t)n_envs, e, NULL)); /* the device copy agrees with the recording too */
            if (memcmp(e, rec_done + (size_t)(K - 1) * n_envs, (size_t)n_envs) != 0) ++mismatches;
            free(e);
            CHECK(lle_vec_destroy(vec));
            CHECK(lle_host_free(act));
            CHECK(lle_host_free(rw));
            CHECK(lle_host_free(dn));
            printf("%s\"s%d\": {\"us_per_step\": %.3f, \"env_steps_per_s\": %.6e, \"mismatches\": %ld, \"env_errors\": %ld, \"parts\": %d}",
                   arg > 4 ? ", " : "", P, 1e6 * best / K, (double)n_envs * K / best, mismatches, errs, (int)np);
            if (mismatches || errs) {
                printf("}}\n");
                return 3;
            }
            continue;
        }
        const int P = atoi(argv[arg]);
        if (P < 1 || P > MAX_PARTS || n_envs % P) {
            fprintf(stderr, "bad part count %d\n", P);
            return 2;
        }
        const int64_t n = n_envs / P;
        lle_vec* part[MAX_PARTS];
        int8_t* act[MAX_PARTS];
        int8_t* stay = NULL;
        float* rw[MAX_PARTS];
        uint8_t* dn[MAX_PARTS];
        CHECK(lle_host_alloc((size_t)n * A, (void**)&stay));
        memset(stay, 4, (size_t)n * A);
        for (int h = 0; h < P; ++h) {
            lle_vec_options o = opts;
            o.env_id_base = (uint64_t)(h * n);
            CHECK(lle_vec_create(maps, 1, NULL, n, &o, &part[h]));
            CHECK(lle_host_alloc((size_t)n * A, (void**)&act[h]));
            CHECK(lle_host_alloc((size_t)n * R * sizeof(float), (void**)&rw[h]));
            CHECK(lle_host_alloc((size_t)n, (void**)&dn[h]));
        }
        double best = 1e30;
        long mismatches = 0;
        for (int rep = 0; rep < 3; ++rep) { /* the first repetition is the warm-up (streams, staging buffers) */
            for (int h = 0; h < P; ++h) {
                CHECK(lle_vec_reset(part[h], NULL, NULL));
                CHECK(lle_vec_set_step_count(part[h], T0));
                CHECK(lle_vec_fetch(part[h], LLE_BUF_DONE, 0, 1, dn[h], NULL)); /* drains the stream */
            }
            mismatches = 0;
            const double t0 = now_s();
            for (int s = 0; s < K; ++s) {
                for (int h = 0; h < P; ++h) {
                    const int8_t* next = rec_act + ((size_t)s * n_envs + (size_t)h * n) * A;
                    if (s > 0) {
                        CHECK(lle_vec_pipeline_wait(part[h], NULL)); /* results of step s-1 of this sub-batch are in host memory */
                        if (memcmp(dn[h], rec_done + (size_t)(s - 1) * n_envs + (size_t)h * n, (size_t)n) != 0) { /* reads every byte */
                            next = stay;
                            ++mismatches;
                        }
                    }
                    memcpy(act[h], next, (size_t)n * A); /* the policy's output lands in the action buffer */
                    CHECK(lle_vec_pipeline_submit(part[h], act[h], rw[h], dn[h], LLE_STREAM_NONE));
                }
            }
            for (int h = 0; h < P; ++h) CHECK(lle_vec_pipeline_wait(part[h], NULL));
            const double dt = now_s() - t0;
            if (rep > 0 && dt < best) best = dt;
        }
        long errs = 0;
        for (int h = 0; h < P; ++h) {
            uint8_t* e = (uint8_t*)malloc((size_t)n);
            CHECK(lle_vec_fetch(part[h], LLE_BUF_ERR, 0, (size_t)n, e, NULL));
            for (int64_t k = 0; k < n; ++k) errs += e[k];
            free(e);
            CHECK(lle_vec_destroy(part[h]));
            CHECK(lle_host_free(act[h]));
            CHECK(lle_host_free(rw[h]));
            CHECK(lle_host_free(dn[h]));
        }
        CHECK(lle_host_free(stay));
        printf("%s\"%d\": {\"us_per_step\": %.3f, \"env_steps_per_s\": %.6e, \"mismatches\": %ld, \"env_errors\": %ld}", arg > 4 ? ", " : "", P,
               1e6 * best / K, (double)n_envs * K / best, mismatches, errs);
        if (mismatches || errs) {
            printf("}}\n");
            return 3;
        }
    }
    printf("}}\n");
    CHECK(lle_host_free(rec_act));
    CHECK(lle_host_free(rec_done));
    lle_map_free(map);
    return 0;
}
