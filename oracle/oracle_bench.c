/*
 * oracle_bench.c — TEST INFRASTRUCTURE ONLY.  Multi-threaded driver around the
 * CPU oracle's orc_step_many, used by bench.py's native CPU baseline leg
 * (the same restated algorithm, one contiguous board range per thread).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>

#include "../include/b2048.h"

int orc_step_many(const uint64_t*, uint64_t*, uint32_t*, uint32_t*, uint8_t*, const uint8_t*, uint8_t*,
                  const uint8_t*, const b2048_env_cfg*, int32_t*, float*, double*, uint8_t*, float*,
                  int32_t*, int64_t, uint64_t, uint64_t, uint32_t);

typedef struct {
    uint64_t* board; uint32_t* score; uint32_t* step; uint8_t* max_exp; float* reward; uint8_t* flags;
    const b2048_env_cfg* cfg; int64_t lo, hi; uint64_t seed, gid0; uint32_t t0; int n_steps;
} job_t;

static void* worker(void* p) {
    job_t* j = (job_t*)p;
    int64_t n = j->hi - j->lo;
    for (int s = 0; s < j->n_steps; ++s)
        orc_step_many(j->board + j->lo, j->board + j->lo, j->score ? j->score + j->lo : 0,
                      j->step ? j->step + j->lo : 0, j->max_exp ? j->max_exp + j->lo : 0, 0, 0, 0, j->cfg, 0,
                      j->reward + j->lo, 0, j->flags + j->lo, 0, 0, n, j->seed, j->gid0 + (uint64_t)j->lo,
                      j->t0 + (uint32_t)s);
    return 0;
}

/* Runs n_steps random-action steps over n boards on n_threads threads; returns seconds. */
double orc_bench_steps(uint64_t* board, uint32_t* score, uint32_t* step, uint8_t* max_exp, float* reward,
                       uint8_t* flags, const b2048_env_cfg* cfg, int64_t n, uint64_t seed, uint64_t gid0,
                       uint32_t t0, int n_steps, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)n_threads);
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int k = 0; k < n_threads; ++k) {
        job_t j = {board, score, step, max_exp, reward, flags, cfg, n * k / n_threads, n * (k + 1) / n_threads,
                   seed, gid0, t0, n_steps};
        jobs[k] = j;
        pthread_create(&th[k], 0, worker, &jobs[k]);
    }
    for (int k = 0; k < n_threads; ++k) pthread_join(th[k], 0);
    clock_gettime(CLOCK_MONOTONIC, &b);
    free(th); free(jobs);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
