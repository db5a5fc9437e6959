/*
 * fb_oracle.h — CPU ORACLE for the FeatureBase roaring hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (featurebase_b200/csrc, libfbgpu.so) never links, loads or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's Go algorithms (FeatureBaseDB/featurebase,
 * all cites are path:line under /root/reference):
 *   roaring/roaring.go           containers, 9-way type-pair kernels, optimize(), serialisation
 *   row.go / fragment.go         Row algebra, fragment.row(), BSI range ops, top()
 *   executor.go                  per-shard evaluation, TopK, GroupBy iterator, reducers
 *
 * Parity pinning: the Go toolchain is absent, so the reference itself cannot run here.  The
 * oracle is pinned against the reference's own literal golden vectors restated under
 * tests/golden/ (TestContainerCombinations table, per-kernel vectors, serialised golden bytes,
 * BSI diagonal tests, executor goldens) and cross-checked against an independent Python
 * set model (oracle/naive.py, the reference's own roaring/naive.go idea).
 */
#ifndef FB_ORACLE_H
#define FB_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* roaring/roaring.go:53-58 */
#define FBO_ARRAY 1
#define FBO_BITMAP 2
#define FBO_RUN 3
#define FBO_ARRAY_MAX_SIZE 4096 /* roaring.go:3036 */
#define FBO_RUN_MAX_SIZE 2048   /* roaring.go:3039 */
#define FBO_BITMAP_N 1024       /* roaring.go:44   */
#define FBO_SHARD_WIDTH_EXP 20  /* shardwidth/helper.go:13 */

typedef struct { uint16_t start, last; } fbo_interval; /* roaring.go:3041-3044, inclusive */

typedef struct fbo_container {
    uint8_t typ;  /* FBO_ARRAY / FBO_BITMAP / FBO_RUN */
    int32_t n;    /* cardinality 0..65536 */
    int32_t len;  /* array: elements; bitmap: 1024; run: intervals */
    void *data;   /* uint16_t[len] | uint64_t[1024] | fbo_interval[len] */
} fbo_container;

/* roaring.Bitmap with slice containers (roaring.go:232, containers_slice.go:5) */
typedef struct fbo_bitmap {
    int64_t n;   /* number of keys */
    int64_t cap;
    uint64_t *keys;
    fbo_container **cs;
    int32_t view; /* != 0: cs[] are borrowed from another bitmap (frozen containers handed out by fragment.row,
                   * fragment.go:283-333 / rbf tx.OffsetRange): fbo_b_free leaves them alone; a view is read-only */
} fbo_bitmap;

/* ---- containers ---- */
fbo_container *fbo_c_array(const uint16_t *v, int32_t n);
fbo_container *fbo_c_bitmap(const uint64_t *words /* 1024 or NULL */);
fbo_container *fbo_c_run(const fbo_interval *iv, int32_t n);
fbo_container *fbo_c_clone(const fbo_container *c);
void fbo_c_free(fbo_container *c);
int32_t fbo_c_n(const fbo_container *c);
int fbo_c_contains(const fbo_container *c, uint16_t v);
int32_t fbo_c_count_runs(const fbo_container *c);
/* in-place canonicalisation; may return NULL (empty) or a different pointer (old one freed) */
fbo_container *fbo_c_optimize(fbo_container *c);
/* convert to an explicit encoding (returns new container) */
fbo_container *fbo_c_convert(const fbo_container *c, int typ);
/* expands to 1024 words */
void fbo_c_to_words(const fbo_container *c, uint64_t *out);

fbo_container *fbo_intersect(const fbo_container *a, const fbo_container *b);
fbo_container *fbo_union(const fbo_container *a, const fbo_container *b);
fbo_container *fbo_difference(const fbo_container *a, const fbo_container *b);
fbo_container *fbo_xor(const fbo_container *a, const fbo_container *b);
fbo_container *fbo_flip(const fbo_container *a);
int32_t fbo_intersection_count(const fbo_container *a, const fbo_container *b);
int32_t fbo_c_count_range(const fbo_container *c, int32_t start, int32_t end);

/* ---- bitmaps ---- */
fbo_bitmap *fbo_b_new(void);
void fbo_b_free(fbo_bitmap *b);
fbo_bitmap *fbo_b_clone(const fbo_bitmap *b);
void fbo_b_put(fbo_bitmap *b, uint64_t key, fbo_container *c); /* takes ownership; keeps sorted */
const fbo_container *fbo_b_get(const fbo_bitmap *b, uint64_t key);
int fbo_b_add(fbo_bitmap *b, uint64_t v); /* DirectAdd */
void fbo_b_add_many(fbo_bitmap *b, const uint64_t *v, int64_t n);
int fbo_b_contains(const fbo_bitmap *b, uint64_t v);
uint64_t fbo_b_count(const fbo_bitmap *b);
int fbo_b_any(const fbo_bitmap *b);
/* fills out[] (cap entries) with up to cap values in ascending order; returns total count */
uint64_t fbo_b_slice(const fbo_bitmap *b, uint64_t *out, uint64_t cap);
fbo_bitmap *fbo_b_intersect(const fbo_bitmap *a, const fbo_bitmap *b);
fbo_bitmap *fbo_b_union(const fbo_bitmap *a, const fbo_bitmap *b);
fbo_bitmap *fbo_b_union_n(const fbo_bitmap *a, const fbo_bitmap *const *others, int n);
fbo_bitmap *fbo_b_difference(const fbo_bitmap *a, const fbo_bitmap *b);
fbo_bitmap *fbo_b_xor(const fbo_bitmap *a, const fbo_bitmap *b);
uint64_t fbo_b_intersection_count(const fbo_bitmap *a, const fbo_bitmap *b);
void fbo_b_optimize(fbo_bitmap *b);
/* OffsetRange(offset,start,end) roaring.go:678-701 */
fbo_bitmap *fbo_b_offset_range(const fbo_bitmap *b, uint64_t offset, uint64_t start, uint64_t end);

/* Pilosa roaring serialisation (roaring.go:1730-1817). Returns bytes needed; writes if cap suffices. */
uint64_t fbo_b_write(fbo_bitmap *b, uint8_t *out, uint64_t cap, int optimize);
/* reads Pilosa (cookie 12348) or official (12346/12347) format; NULL on error */
fbo_bitmap *fbo_b_read(const uint8_t *buf, uint64_t len);

/* ---- fragment-level (fragment.go) : a fragment is a bitmap with pos = row<<20 | col&(2^20-1) ---- */
fbo_bitmap *fbo_frag_row(const fbo_bitmap *frag, uint64_t row, uint64_t shard); /* fragment.go:283-333 */

/* pql tokens used by rangeOp (pql/token.go) — values are ours */
#define FBO_OP_EQ 1
#define FBO_OP_NEQ 2
#define FBO_OP_LT 3
#define FBO_OP_LTE 4
#define FBO_OP_GT 5
#define FBO_OP_GTE 6
#define FBO_OP_BETWEEN 7
/* fragment.rangeOp / rangeBetween (fragment.go:937-1303); result keys are shard-absolute */
fbo_bitmap *fbo_frag_range_op(const fbo_bitmap *frag, uint64_t shard, int op, uint64_t bit_depth,
                              int64_t predicate, int64_t predicate_max);

/* doTopK (executor.go:2705-2746): per-row count of fragment ∩ filter (filter may be NULL, shard-absolute keys).
 * Writes up to cap (row,count) with count>0 in ascending row order; returns number of rows. */
int64_t fbo_frag_row_counts(const fbo_bitmap *frag, uint64_t shard, const fbo_bitmap *filter,
                            uint64_t *rows, uint64_t *counts, int64_t cap);
/* fragment.rows() (fragment.go:2465): distinct row ids present */
int64_t fbo_frag_rows(const fbo_bitmap *frag, uint64_t *rows, int64_t cap);

/* groupByIterator for 2..4 set fields on one shard (executor.go:8617-8934, executeGroupByShard :3918):
 * frags[i] are the fragments (NULL => whole shard contributes nothing), row_ids per field (flat, n_rows[i]),
 * filter optional (shard-absolute). Adds counts into out_counts (dense, row-major, rightmost fastest). */
int fbo_groupby_shard(const fbo_bitmap *const *frags, int n_fields, uint64_t shard,
                      const uint64_t *row_ids_flat, const int32_t *n_rows,
                      const fbo_bitmap *filter, uint64_t *out_counts);

/* fragment.row as the reference hands it out: a Bitmap over the fragment's own (frozen) containers, no payload copy
 * (fragment.go:283-333, rbf/tx.go:1586-1637).  Read-only; free with fbo_b_free. */
fbo_bitmap *fbo_frag_row_view(const fbo_bitmap *frag, uint64_t row, uint64_t shard);

/* ---- multi-threaded CPU baseline (bench.py --impl reference / cpu_baseline; fb_bench.c) ----
 * The worker pool is created ONCE (threads pinned round-robin to the host cores) and reused by every call: the
 * reference's task.Pool workers are long-lived goroutines (executor.go:6742-6812), so thread creation is not part of a
 * query.  Shards are handed out dynamically in small blocks (what mapperLocal's job channel does).  Rows are fetched as
 * views (no container cloning).  Every bench call returns its wall time through *seconds. */
typedef struct fbo_pool fbo_pool;
fbo_pool *fbo_pool_create(int n_threads);            /* pinned round-robin to the allowed CPUs */
fbo_pool *fbo_pool_create2(int n_threads, int pin);  /* pin == 0: the OS places the threads */
void fbo_pool_destroy(fbo_pool *p);
int fbo_pool_threads(const fbo_pool *p);
/* Count(Intersect(Union(rows a..), Union(rows b..))): frags[s] is the fragment of shard shards[s] (executeUnionShard
 * executor.go:5382 -> Row.Union row.go:288 -> unionInPlace roaring.go:1410; executeIntersectShard :5357; Count :5871) */
uint64_t fbo_bench_union_intersect_count(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                                         const uint64_t *rows_a, int na, const uint64_t *rows_b, int nb, double *seconds);
/* n_pairs queries Count(Intersect(Row(a_k), Row(b_k))) over the same fragments; materialise != 0: the executor's path
 * (Row.Intersect materialises, then Count sums N; executor.go:5357,5871); 0: Bitmap.IntersectionCount (roaring.go:928).
 * out_counts[k] += the count of pair k (may be NULL). */
uint64_t fbo_bench_union_intersect_per_shard(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                                             const uint64_t *rows_a, int na, const uint64_t *rows_b, int nb, uint64_t *per_shard);
uint64_t fbo_bench_pair_counts(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                               const uint64_t *rows_a, const uint64_t *rows_b, int n_pairs, int materialise,
                               uint64_t *out_counts, double *seconds);
/* Count(Row(v <op> predicate)) over BSI fragments: fragment.rangeOp (fragment.go:937-1303) + Count */
uint64_t fbo_bench_range_count(fbo_pool *p, const fbo_bitmap *const *frags, const uint64_t *shards, int64_t n_shards,
                               int op, uint64_t bit_depth, int64_t predicate, int64_t predicate_max, double *seconds);
/* GroupBy over n_fields set fields: frags[f * n_shards + s]; out_counts dense, summed over the shards (mergeGroupCounts
 * executor.go:3728); groupByIterator executor.go:8617-8934 */
int fbo_bench_groupby(fbo_pool *p, const fbo_bitmap *const *frags, int n_fields, const uint64_t *shards, int64_t n_shards,
                      const uint64_t *row_ids_flat, const int32_t *n_rows, uint64_t *out_counts, double *seconds);

void fbo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
