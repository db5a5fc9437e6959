/*
 * fbdatagen.c — synthetic fragment generator for tests and bench.py (NOT product, NOT oracle).
 *
 * Emits Pilosa-roaring bytes exactly as roaring.Bitmap.WriteTo would for the generated fragment
 * (format: /root/reference/roaring/roaring.go:1738-1817; canonical encodings per optimize() :3412-3461),
 * with fragment-relative keys row*16+slot (fragment.go:2780-2782).  Both the CPU oracle and the GPU
 * library are fed the same bytes, so the generator needs no twin.
 *
 * Counter-based RNG: every (seed, field, row, shard) stream is splitmix64, so output is independent of
 * thread count.  Workloads (SURVEY.md §8d / BASELINE.md §3):
 *   mode 0 "uniform"   : Bernoulli(p) per column (geometric gap sampling)
 *   mode 1 "clustered" : alternating gaps/runs, geometric run length with the given mean, overall density p
 *   bsi                : exists/sign/bit-plane rows of uniform values (cfg 3)
 *   groupby            : each record column gets one row in field a and one in field b (cfg 4)
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#define MASTER_SEED 0xFEA7B45E5EED0001ull

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t mix(uint64_t a, uint64_t b) { uint64_t s = a ^ (b * 0xD6E8FEB86659FD93ull); return splitmix64(&s); }
static inline double u01(uint64_t *s) { return ((double)(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

typedef struct { uint8_t *p; uint64_t len, cap; } buf_t;
static void b_need(buf_t *b, uint64_t n) {
    if (b->len + n <= b->cap) return;
    uint64_t c = b->cap ? b->cap * 2 : 1 << 16; while (c < b->len + n) c *= 2;
    b->p = realloc(b->p, c); if (!b->p) abort(); b->cap = c;
}
static void b_put(buf_t *b, const void *src, uint64_t n) { b_need(b, n); memcpy(b->p + b->len, src, n); b->len += n; }

/* one fragment under construction: header entries + payload stream */
typedef struct { uint64_t key; uint16_t typ; uint16_t n1; uint32_t size; } chdr_t;
typedef struct { chdr_t *h; int64_t n, cap; buf_t payload; } frag_t;

static void frag_add_hdr(frag_t *f, uint64_t key, uint16_t typ, int32_t n, uint32_t size) {
    if (f->n == f->cap) { f->cap = f->cap ? f->cap * 2 : 64; f->h = realloc(f->h, sizeof(chdr_t) * f->cap); }
    f->h[f->n++] = (chdr_t){ key, typ, (uint16_t)(n - 1), size };
}

/* emit one container given its 1024-word bitmap image, choosing the canonical encoding */
static void emit_words(frag_t *f, uint64_t key, const uint64_t *w) {
    int32_t n = 0, runs = 0;
    for (int i = 0; i < 1024; i++) {
        uint64_t v = w[i]; n += __builtin_popcountll(v);
        uint64_t prev = i ? (w[i - 1] >> 63) : 0;
        runs += __builtin_popcountll(v & ~((v << 1) | prev));
    }
    if (n == 0) return;
    if (runs <= 2048 && runs <= n / 2) { /* run container: u16 count then {start,last} */
        uint16_t rc = (uint16_t)runs; b_put(&f->payload, &rc, 2);
        int inrun = 0; uint32_t start = 0;
        for (uint32_t v = 0; v < 65536; v++) {
            if ((v & 63) == 0 && !inrun && w[v >> 6] == 0) { v += 63; continue; }
            int bit = (w[v >> 6] >> (v & 63)) & 1;
            if (bit && !inrun) { inrun = 1; start = v; }
            else if (!bit && inrun) { inrun = 0; uint16_t iv[2] = { (uint16_t)start, (uint16_t)(v - 1) }; b_put(&f->payload, iv, 4); }
        }
        if (inrun) { uint16_t iv[2] = { (uint16_t)start, 65535 }; b_put(&f->payload, iv, 4); }
        frag_add_hdr(f, key, 3, n, 2 + 4 * (uint32_t)runs);
    } else if (n < 4096) {
        b_need(&f->payload, (uint64_t)n * 2);
        uint16_t *o = (uint16_t *)(f->payload.p + f->payload.len);
        for (int i = 0; i < 1024; i++) { uint64_t v = w[i]; while (v) { *o++ = (uint16_t)(i * 64 + __builtin_ctzll(v)); v &= v - 1; } }
        f->payload.len += (uint64_t)n * 2;
        frag_add_hdr(f, key, 1, n, (uint32_t)n * 2);
    } else {
        b_put(&f->payload, w, 8192);
        frag_add_hdr(f, key, 2, n, 8192);
    }
}

/* assemble Pilosa roaring file: cookie, count, headers, offsets, payloads */
static uint8_t *frag_finish(frag_t *f, uint64_t *out_len) {
    uint64_t total = 8 + (uint64_t)f->n * 16 + f->payload.len;
    uint8_t *o = malloc(total ? total : 1); if (!o) abort();
    uint32_t cookie = 12348, cnt = (uint32_t)f->n;
    memcpy(o, &cookie, 4); memcpy(o + 4, &cnt, 4);
    uint8_t *h = o + 8, *offp = o + 8 + f->n * 12; uint64_t off = 8 + (uint64_t)f->n * 16;
    for (int64_t i = 0; i < f->n; i++) {
        memcpy(h, &f->h[i].key, 8); memcpy(h + 8, &f->h[i].typ, 2); memcpy(h + 10, &f->h[i].n1, 2); h += 12;
        uint32_t o32 = (uint32_t)off; memcpy(offp, &o32, 4); offp += 4; off += f->h[i].size;
    }
    memcpy(o + 8 + (uint64_t)f->n * 16, f->payload.p, f->payload.len);
    free(f->h); free(f->payload.p);
    *out_len = total;
    return o;
}

/* fill a 2^20-bit row image (16384 words) */
static void gen_row(uint64_t *row, uint64_t seed, uint32_t field, uint64_t rowid, uint64_t shard, double p, int mode, double mean_run) {
    memset(row, 0, 16384 * 8);
    uint64_t s = mix(mix(mix(seed, field), rowid), shard);
    const uint64_t W = 1ull << 20;
    if (p <= 0) return;
    if (mode == 0) {
        if (p >= 1.0) { memset(row, 0xff, 16384 * 8); return; }
        double il = 1.0 / log1p(-p);
        uint64_t c = 0;
        for (;;) {
            double g = floor(log(u01(&s)) * il); /* failures before next success */
            if (g >= (double)W) break;
            c += (uint64_t)g; if (c >= W) break;
            row[c >> 6] |= 1ull << (c & 63); c++;
        }
    } else {
        /* runs of geometric length (mean mean_run) separated by geometric gaps (mean g) with p = run/(run+gap) */
        double mr = mean_run < 1 ? 1 : mean_run, mg = mr * (1.0 - p) / p; if (mg < 1) mg = 1;
        double lr = mr > 1 ? 1.0 / log1p(-1.0 / mr) : 0, lg = mg > 1 ? 1.0 / log1p(-1.0 / mg) : 0;
        uint64_t c = 0;
        for (;;) {
            uint64_t gap = 1 + (lg != 0 ? (uint64_t)fmin(floor(log(u01(&s)) * lg), 1e9) : 0);
            c += gap; if (c >= W) break;
            uint64_t len = 1 + (lr != 0 ? (uint64_t)fmin(floor(log(u01(&s)) * lr), 1e9) : 0);
            for (uint64_t k = 0; k < len && c < W; k++, c++) row[c >> 6] |= 1ull << (c & 63);
        }
    }
}

/* ---- public: set-field fragment with the given rows ---- */
uint8_t *fbdg_fragment(uint64_t seed, uint32_t field, uint64_t shard, const uint64_t *rows, int n_rows,
                       double p, int mode, double mean_run, uint64_t *out_len) {
    frag_t f = {0}; uint64_t *row = malloc(16384 * 8);
    if (!seed) seed = MASTER_SEED;
    for (int r = 0; r < n_rows; r++) { /* rows must be ascending so keys are ascending */
        gen_row(row, seed, field, rows[r], shard, p, mode, mean_run);
        for (int slot = 0; slot < 16; slot++) emit_words(&f, rows[r] * 16 + (uint64_t)slot, row + slot * 1024);
    }
    free(row);
    return frag_finish(&f, out_len);
}

/* ---- BSI fragment (fragment.go:63-65: row0 exists, row1 sign, row 2+i bit i); values uniform in [lo,hi], stored as value-base ---- */
uint8_t *fbdg_bsi_fragment(uint64_t seed, uint32_t field, uint64_t shard, uint64_t n_cols, int bit_depth,
                           int64_t lo, int64_t hi, int64_t base, double null_frac, uint64_t *out_len) {
    if (!seed) seed = MASTER_SEED;
    int nrows = 2 + bit_depth;
    uint64_t *img = calloc((size_t)nrows * 16384, 8);
    uint64_t span = (uint64_t)(hi - lo) + 1; /* 0 means 2^64 */
    for (uint64_t c = 0; c < n_cols && c < (1ull << 20); c++) {
        uint64_t s = mix(mix(mix(seed, field), shard), c);
        uint64_t r1 = splitmix64(&s), r2 = splitmix64(&s);
        if (null_frac > 0 && (double)(r2 >> 11) * (1.0 / 9007199254740992.0) < null_frac) continue;
        int64_t v = lo + (int64_t)(span ? r1 % span : r1);
        int64_t d = v - base; uint64_t mag = d < 0 ? (uint64_t)(-d) : (uint64_t)d;
        img[0 * 16384 + (c >> 6)] |= 1ull << (c & 63);
        if (d < 0) img[1 * 16384 + (c >> 6)] |= 1ull << (c & 63);
        for (int i = 0; i < bit_depth; i++) if ((mag >> i) & 1) img[(size_t)(2 + i) * 16384 + (c >> 6)] |= 1ull << (c & 63);
    }
    frag_t f = {0};
    for (int r = 0; r < nrows; r++) for (int slot = 0; slot < 16; slot++) emit_words(&f, (uint64_t)r * 16 + slot, img + (size_t)r * 16384 + slot * 1024);
    free(img);
    return frag_finish(&f, out_len);
}
/* the value the BSI generator assigned to a column (for test oracles); returns 0 if null */
int fbdg_bsi_value(uint64_t seed, uint32_t field, uint64_t shard, uint64_t col, int64_t lo, int64_t hi, double null_frac, int64_t *out) {
    if (!seed) seed = MASTER_SEED;
    uint64_t s = mix(mix(mix(seed, field), shard), col);
    uint64_t r1 = splitmix64(&s), r2 = splitmix64(&s);
    if (null_frac > 0 && (double)(r2 >> 11) * (1.0 / 9007199254740992.0) < null_frac) return 0;
    uint64_t span = (uint64_t)(hi - lo) + 1;
    *out = lo + (int64_t)(span ? r1 % span : r1);
    return 1;
}

/* ---- GroupBy pair of fragments: record columns Bernoulli(p_rec); each record gets row a in [0,na), row b in [0,nb) ---- */
int fbdg_groupby_fragments(uint64_t seed, uint32_t field_a, uint32_t field_b, uint64_t shard, double p_rec, int na, int nb,
                           uint8_t **out_a, uint64_t *len_a, uint8_t **out_b, uint64_t *len_b) {
    if (!seed) seed = MASTER_SEED;
    uint64_t *rec = malloc(16384 * 8);
    gen_row(rec, seed, 0xC01u, 0, shard, p_rec, 0, 0);
    /* bucket columns by row: two passes per field */
    for (int which = 0; which < 2; which++) {
        int nr = which ? nb : na; uint32_t fld = which ? field_b : field_a;
        uint32_t *cnt = calloc((size_t)nr + 1, 4);
        for (uint64_t c = 0; c < (1ull << 20); c++) if ((rec[c >> 6] >> (c & 63)) & 1) { uint64_t r = mix(mix(mix(seed, fld), shard), c) % (uint64_t)nr; cnt[r + 1]++; }
        for (int r = 0; r < nr; r++) cnt[r + 1] += cnt[r];
        uint32_t total = cnt[nr]; uint32_t *cols = malloc((size_t)(total ? total : 1) * 4); uint32_t *fill = malloc((size_t)nr * 4); memcpy(fill, cnt, (size_t)nr * 4);
        for (uint64_t c = 0; c < (1ull << 20); c++) if ((rec[c >> 6] >> (c & 63)) & 1) { uint64_t r = mix(mix(mix(seed, fld), shard), c) % (uint64_t)nr; cols[fill[r]++] = (uint32_t)c; }
        frag_t f = {0}; uint64_t w[1024];
        for (int r = 0; r < nr; r++) {
            uint32_t i = cnt[r], e = cnt[r + 1];
            while (i < e) {
                uint32_t slot = cols[i] >> 16; memset(w, 0, sizeof w);
                while (i < e && (cols[i] >> 16) == slot) { uint32_t v = cols[i] & 0xffff; w[v >> 6] |= 1ull << (v & 63); i++; }
                emit_words(&f, (uint64_t)r * 16 + slot, w);
            }
        }
        free(cnt); free(cols); free(fill);
        if (which) *out_b = frag_finish(&f, len_b); else *out_a = frag_finish(&f, len_a);
    }
    free(rec);
    return 0;
}

/* ---- bulk, multi-threaded: n_shards set-field fragments into one buffer ---- */
typedef struct {
    uint64_t seed; uint32_t field; const uint64_t *shards; int64_t lo, hi; const uint64_t *rows; int n_rows;
    double p; int mode; double mean_run; uint8_t **bufs; uint64_t *lens;
} bulk_arg;
static void *bulk_worker(void *vp) {
    bulk_arg *a = vp;
    for (int64_t s = a->lo; s < a->hi; s++) a->bufs[s] = fbdg_fragment(a->seed, a->field, a->shards[s], a->rows, a->n_rows, a->p, a->mode, a->mean_run, &a->lens[s]);
    return NULL;
}
/* returns one malloc'd buffer; offsets[n_shards+1] gives each fragment's byte range */
uint8_t *fbdg_fragments(uint64_t seed, uint32_t field, const uint64_t *shards, int64_t n_shards, const uint64_t *rows, int n_rows,
                        double p, int mode, double mean_run, int n_threads, uint64_t *offsets) {
    if (n_threads < 1) n_threads = 1; if (n_threads > n_shards) n_threads = (int)(n_shards ? n_shards : 1);
    uint8_t **bufs = calloc((size_t)n_shards + 1, sizeof(void *)); uint64_t *lens = calloc((size_t)n_shards + 1, 8);
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads); bulk_arg *args = malloc(sizeof(bulk_arg) * n_threads);
    for (int t = 0; t < n_threads; t++) {
        args[t] = (bulk_arg){ seed, field, shards, n_shards * t / n_threads, n_shards * (t + 1) / n_threads, rows, n_rows, p, mode, mean_run, bufs, lens };
        pthread_create(&th[t], NULL, bulk_worker, &args[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    uint64_t total = 0; for (int64_t s = 0; s < n_shards; s++) { offsets[s] = total; total += lens[s]; } offsets[n_shards] = total;
    uint8_t *out = malloc(total ? total : 1);
    for (int64_t s = 0; s < n_shards; s++) { memcpy(out + offsets[s], bufs[s], lens[s]); free(bufs[s]); }
    free(bufs); free(lens); free(th); free(args);
    return out;
}

void fbdg_free(void *p) { free(p); }
