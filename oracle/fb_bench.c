/*
 * fb_bench.c — multi-threaded CPU baseline over the oracle (TEST / BENCH INFRASTRUCTURE ONLY, see fb_oracle.h).
 *
 * What bench.py's `--impl reference` arm and every `cpu_baseline` record time.  It restates how the reference runs a
 * query on one node: executor.mapReduce -> mapperLocal hands the shards to a pool of long-lived workers
 * (executor.go:6449-6533, 6742-6812, task/pool.go), every worker evaluates the call tree on its shard with the roaring
 * kernels (fb_oracle.c), the reducer adds the per-shard results.  Go is absent here, so this is `kind: "port"`.
 *
 *   - the pool is created once and its threads are pinned; no thread is created inside a timed region
 *   - fragment.row hands out views over the fragment's frozen containers (no payload copy), as rowFromStorage does
 *   - shards are pulled from a shared counter in blocks of 1 (the reference's job channel)
 */
#define _GNU_SOURCE
#include "fb_oracle.h"
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef void (*job_fn)(void *arg, int64_t shard_index, int worker);

struct fbo_pool {
    int n;
    pthread_t *th;
    pthread_mutex_t mu;
    pthread_cond_t cv_start, cv_done;
    uint64_t gen;            /* bumped for every job */
    int running, quit;
    job_fn fn; void *arg; int64_t n_items;
    volatile int64_t next;   /* next shard index to hand out */
};

typedef struct { fbo_pool *p; int id; } worker_arg;

static void *pool_worker(void *a) {
    worker_arg *wa = a; fbo_pool *p = wa->p; const int id = wa->id; free(wa);
    uint64_t seen = 0;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (!p->quit && p->gen == seen) pthread_cond_wait(&p->cv_start, &p->mu);
        if (p->quit) { pthread_mutex_unlock(&p->mu); return NULL; }
        seen = p->gen;
        pthread_mutex_unlock(&p->mu);
        for (;;) {
            int64_t i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
            if (i >= p->n_items) break;
            p->fn(p->arg, i, id);
        }
        pthread_mutex_lock(&p->mu);
        if (--p->running == 0) pthread_cond_signal(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
}

fbo_pool *fbo_pool_create2(int n_threads, int pin) {
    if (n_threads < 1) n_threads = 1;
    fbo_pool *p = calloc(1, sizeof *p);
    if (!p) return NULL;
    p->n = n_threads; p->th = calloc((size_t)n_threads, sizeof(pthread_t));
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->cv_start, NULL); pthread_cond_init(&p->cv_done, NULL);
    cpu_set_t allowed; CPU_ZERO(&allowed);
    int n_allowed = 0, cpus[CPU_SETSIZE];
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0) for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus[n_allowed++] = c;
    for (int t = 0; t < n_threads; t++) {
        worker_arg *wa = malloc(sizeof *wa); wa->p = p; wa->id = t;
        pthread_create(&p->th[t], NULL, pool_worker, wa);
        if (pin && n_allowed > 0) { cpu_set_t one; CPU_ZERO(&one); CPU_SET(cpus[t % n_allowed], &one); pthread_setaffinity_np(p->th[t], sizeof one, &one); }
    }
    return p;
}
fbo_pool *fbo_pool_create(int n_threads) { return fbo_pool_create2(n_threads, 1); }
void fbo_pool_destroy(fbo_pool *p) {
    if (!p) return;
    pthread_mutex_lock(&p->mu); p->quit = 1; pthread_cond_broadcast(&p->cv_start); pthread_mutex_unlock(&p->mu);
    for (int t = 0; t < p->n; t++) pthread_join(p->th[t], NULL);
    pthread_mutex_destroy(&p->mu); pthread_cond_destroy(&p->cv_start); pthread_cond_destroy(&p->cv_done);
    free(p->th); free(p);
}
int fbo_pool_threads(const fbo_pool *p) { return p ? p->n : 0; }

/* runs fn(arg, i, worker) for i in [0, n_items) on the pool; returns the wall time of the whole job */
static double pool_run(fbo_pool *p, job_fn fn, void *arg, int64_t n_items) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_mutex_lock(&p->mu);
    p->fn = fn; p->arg = arg; p->n_items = n_items; p->next = 0; p->running = p->n; p->gen++;
    pthread_cond_broadcast(&p->cv_start);
    while (p->running) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

#define MAX_WORKERS 1024
typedef struct { uint64_t v; char pad[56]; } padded_u64;   /* one cache line per worker */

/* ---- Count(Intersect(Union(rows a), Union(rows b))) ---- */
typedef struct {
    const fbo_bitmap *const *frags; const uint64_t *shards;
    const uint64_t *ra; int na; const uint64_t *rb; int nb;
    padded_u64 *tot; uint64_t *per_shard;   /* per_shard: optional [n_shards] output, one writer per entry */
} uic_arg;

static fbo_bitmap *union_rows(const fbo_bitmap *frag, uint64_t shard, const uint64_t *rows, int n) { /* executeUnionShard executor.go:5382 */
    if (n == 0) return fbo_b_new();
    if (n == 1) return fbo_frag_row_view(frag, rows[0], shard);
    fbo_bitmap *r[64]; fbo_bitmap **rv = n <= 64 ? r : malloc(sizeof(void *) * (size_t)n);
    for (int i = 0; i < n; i++) rv[i] = fbo_frag_row_view(frag, rows[i], shard);
    fbo_bitmap *o = fbo_b_union_n(rv[0], (const fbo_bitmap *const *)(rv + 1), n - 1);   /* Row.Union row.go:288 -> unionInPlace */
    for (int i = 0; i < n; i++) fbo_b_free(rv[i]);
    if (rv != r) free(rv);
    return o;
}
static void uic_job(void *p, int64_t s, int w) {
    uic_arg *a = p;
    fbo_bitmap *ua = union_rows(a->frags[s], a->shards[s], a->ra, a->na);
    fbo_bitmap *ub = union_rows(a->frags[s], a->shards[s], a->rb, a->nb);
    fbo_bitmap *x = fbo_b_intersect(ua, ub);          /* executeIntersectShard executor.go:5357 */
    const uint64_t c = fbo_b_count(x);                /* executeCount executor.go:5871-5877 */
    a->tot[w].v += c;
    if (a->per_shard) a->per_shard[s] = c;
    fbo_b_free(ua); fbo_b_free(ub); fbo_b_free(x);
}
uint64_t fbo_bench_union_intersect_count(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                                         const uint64_t *rows_a, int na, const uint64_t *rows_b, int nb, double *seconds) {
    padded_u64 *tot = calloc((size_t)p->n, sizeof *tot);
    uic_arg a = { frags, shards, rows_a, na, rows_b, nb, tot, NULL };
    double sec = pool_run(p, uic_job, &a, n_shards);
    uint64_t total = 0; for (int t = 0; t < p->n; t++) total += tot[t].v;
    free(tot);
    if (seconds) *seconds = sec;
    return total;
}
/* the same query with the per-shard counts kept (full-size parity tests compare EVERY shard with the device's per-shard vector) */
uint64_t fbo_bench_union_intersect_per_shard(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                                             const uint64_t *rows_a, int na, const uint64_t *rows_b, int nb, uint64_t *per_shard) {
    padded_u64 *tot = calloc((size_t)p->n, sizeof *tot);
    uic_arg a = { frags, shards, rows_a, na, rows_b, nb, tot, per_shard };
    pool_run(p, uic_job, &a, n_shards);
    uint64_t total = 0; for (int t = 0; t < p->n; t++) total += tot[t].v;
    free(tot);
    return total;
}

/* ---- n_pairs x Count(Intersect(Row a_k, Row b_k)) ---- */
typedef struct {
    const fbo_bitmap *const *frags; const uint64_t *shards; int64_t n_shards;
    const uint64_t *ra, *rb; int n_pairs, materialise; padded_u64 *tot; uint64_t *per_pair;
} pc_arg;
static void pc_job(void *p, int64_t item, int w) {
    pc_arg *a = p;
    const int64_t s = item % a->n_shards; const int k = (int)(item / a->n_shards);
    fbo_bitmap *x = fbo_frag_row_view(a->frags[s], a->ra[k], a->shards[s]);
    fbo_bitmap *y = fbo_frag_row_view(a->frags[s], a->rb[k], a->shards[s]);
    uint64_t c;
    if (a->materialise) { fbo_bitmap *z = fbo_b_intersect(x, y); c = fbo_b_count(z); fbo_b_free(z); }
    else c = fbo_b_intersection_count(x, y);
    fbo_b_free(x); fbo_b_free(y);
    a->tot[w].v += c;
    if (a->per_pair && c) __atomic_fetch_add(&a->per_pair[k], c, __ATOMIC_RELAXED);
}
uint64_t fbo_bench_pair_counts(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                               const uint64_t *rows_a, const uint64_t *rows_b, int n_pairs, int materialise,
                               uint64_t *out_counts, double *seconds) {
    padded_u64 *tot = calloc((size_t)p->n, sizeof *tot);
    pc_arg a = { frags, shards, n_shards, rows_a, rows_b, n_pairs, materialise, tot, out_counts };
    double sec = pool_run(p, pc_job, &a, n_shards * (int64_t)n_pairs);
    uint64_t total = 0; for (int t = 0; t < p->n; t++) total += tot[t].v;
    free(tot);
    if (seconds) *seconds = sec;
    return total;
}

/* ---- Count(Row(v <op> predicate)) ---- */
typedef struct { const fbo_bitmap *const *frags; const uint64_t *shards; int op; uint64_t depth; int64_t pred, pmax; padded_u64 *tot; } rc_arg;
static void rc_job(void *p, int64_t s, int w) {
    rc_arg *a = p;
    fbo_bitmap *r = fbo_frag_range_op(a->frags[s], a->shards[s], a->op, a->depth, a->pred, a->pmax);
    a->tot[w].v += fbo_b_count(r);
    fbo_b_free(r);
}
uint64_t fbo_bench_range_count(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                               int op, uint64_t bit_depth, int64_t predicate, int64_t predicate_max, double *seconds) {
    padded_u64 *tot = calloc((size_t)p->n, sizeof *tot);
    rc_arg a = { frags, shards, op, bit_depth, predicate, predicate_max, tot };
    double sec = pool_run(p, rc_job, &a, n_shards);
    uint64_t total = 0; for (int t = 0; t < p->n; t++) total += tot[t].v;
    free(tot);
    if (seconds) *seconds = sec;
    return total;
}

/* ---- GroupBy ---- */
typedef struct {
    const fbo_bitmap *const *frags; int nf; const uint64_t *shards; int64_t n_shards; const uint64_t *ids; const int32_t *n_rows;
    uint64_t **acc; size_t n_groups;
} gb_arg;
static void gb_job(void *p, int64_t s, int w) {
    gb_arg *a = p;
    const fbo_bitmap *fr[8];
    for (int f = 0; f < a->nf; f++) fr[f] = a->frags[(int64_t)f * a->n_shards + s];
    if (!a->acc[w]) a->acc[w] = calloc(a->n_groups, 8);        /* per-worker partial result, merged at the end (mergeGroupCounts) */
    fbo_groupby_shard(fr, a->nf, a->shards[s], a->ids, a->n_rows, NULL, a->acc[w]);
}
int fbo_bench_groupby(fbo_pool *p, const fbo_bitmap *const *frags, int n_fields, const uint64_t *shards, int64_t n_shards,
                      const uint64_t *row_ids_flat, const int32_t *n_rows, uint64_t *out_counts, double *seconds) {
    if (n_fields < 1 || n_fields > 8) return -1;
    size_t ng = 1; for (int f = 0; f < n_fields; f++) ng *= (size_t)n_rows[f];
    uint64_t **acc = calloc((size_t)p->n, sizeof(void *));
    gb_arg a = { frags, n_fields, shards, n_shards, row_ids_flat, n_rows, acc, ng };
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    pool_run(p, gb_job, &a, n_shards);
    for (int t = 0; t < p->n; t++) if (acc[t]) { for (size_t g = 0; g < ng; g++) out_counts[g] += acc[t][g]; free(acc[t]); }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(acc);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return 0;
}
