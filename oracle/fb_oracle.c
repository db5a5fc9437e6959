/*
 * fb_oracle.c — CPU ORACLE (test infrastructure only; see fb_oracle.h header comment).
 *
 * Plain-C restatement of the reference's Go roaring / Row / fragment / executor algorithms for the
 * hot path.  Every function cites the reference file:line it follows (paths under /root/reference).
 * No SIMD, -O2: it doubles as the "CPU restatement of reference algorithms" baseline.
 */
#define _GNU_SOURCE
#include "fb_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

#define MAXV 65535u

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) abort(); return p; }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s ? s : 1); if (!p) abort(); return p; }
void fbo_free(void *p) { free(p); }

static inline int popcnt64(uint64_t x) { return __builtin_popcountll(x); }
static inline int bitlen64(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; } /* bits.Len64 */

/* ------------------------------------------------------------------ containers */

static fbo_container *c_alloc(uint8_t typ, int32_t len) {
    fbo_container *c = xmalloc(sizeof *c);
    c->typ = typ; c->n = 0; c->len = len;
    size_t bytes = typ == FBO_ARRAY ? (size_t)len * 2 : typ == FBO_BITMAP ? 8192 : (size_t)len * 4;
    c->data = xcalloc(1, bytes ? bytes : 8);
    return c;
}
static inline uint16_t *ARR(const fbo_container *c) { return (uint16_t *)c->data; }
static inline uint64_t *BMP(const fbo_container *c) { return (uint64_t *)c->data; }
static inline fbo_interval *RUN(const fbo_container *c) { return (fbo_interval *)c->data; }

fbo_container *fbo_c_array(const uint16_t *v, int32_t n) {
    fbo_container *c = c_alloc(FBO_ARRAY, n);
    if (n) memcpy(c->data, v, (size_t)n * 2);
    c->n = n;
    return c;
}
fbo_container *fbo_c_bitmap(const uint64_t *w) {
    fbo_container *c = c_alloc(FBO_BITMAP, FBO_BITMAP_N);
    if (w) { memcpy(c->data, w, 8192); int32_t n = 0; for (int i = 0; i < 1024; i++) n += popcnt64(w[i]); c->n = n; }
    return c;
}
fbo_container *fbo_c_run(const fbo_interval *iv, int32_t n) {
    fbo_container *c = c_alloc(FBO_RUN, n);
    int32_t card = 0;
    for (int i = 0; i < n; i++) { RUN(c)[i] = iv[i]; card += (int32_t)iv[i].last - iv[i].start + 1; }
    c->n = card;
    return c;
}
fbo_container *fbo_c_clone(const fbo_container *c) {
    if (!c) return NULL;
    fbo_container *o = c_alloc(c->typ, c->len);
    size_t bytes = c->typ == FBO_ARRAY ? (size_t)c->len * 2 : c->typ == FBO_BITMAP ? 8192 : (size_t)c->len * 4;
    memcpy(o->data, c->data, bytes); o->n = c->n;
    return o;
}
void fbo_c_free(fbo_container *c) { if (c) { free(c->data); free(c); } }
int32_t fbo_c_n(const fbo_container *c) { return c ? c->n : 0; } /* nil container == empty, container_stash.go:38 */

int fbo_c_contains(const fbo_container *c, uint16_t v) {
    if (!c) return 0;
    if (c->typ == FBO_BITMAP) return (BMP(c)[v >> 6] >> (v & 63)) & 1;
    if (c->typ == FBO_ARRAY) {
        int lo = 0, hi = c->len;
        while (lo < hi) { int m = (lo + hi) >> 1; if (ARR(c)[m] < v) lo = m + 1; else hi = m; }
        return lo < c->len && ARR(c)[lo] == v;
    }
    for (int i = 0; i < c->len; i++) if (v >= RUN(c)[i].start && v <= RUN(c)[i].last) return 1;
    return 0;
}

/* arrayCountRuns roaring.go:3382, bitmapCountRuns :3372, countRuns :3397 */
int32_t fbo_c_count_runs(const fbo_container *c) {
    if (!c) return 0;
    if (c->typ == FBO_RUN) return c->len;
    if (c->typ == FBO_ARRAY) {
        int32_t r = 0; int32_t prev = -2;
        for (int i = 0; i < c->len; i++) { if ((int32_t)ARR(c)[i] != prev + 1) r++; prev = ARR(c)[i]; }
        return r;
    }
    /* bitmap: count 0->1 transitions */
    int32_t r = 0; const uint64_t *w = BMP(c);
    for (int i = 0; i < 1024; i++) {
        uint64_t v = w[i];
        uint64_t prevbit = i ? (w[i - 1] >> 63) : 0;
        /* starts = bits set whose predecessor bit is clear */
        uint64_t starts = v & ~((v << 1) | prevbit);
        r += popcnt64(starts);
    }
    return r;
}

void fbo_c_to_words(const fbo_container *c, uint64_t *out) {
    memset(out, 0, 8192);
    if (!c) return;
    if (c->typ == FBO_BITMAP) { memcpy(out, c->data, 8192); return; }
    if (c->typ == FBO_ARRAY) { /* arrayToBitmap roaring.go:3756 */
        for (int i = 0; i < c->len; i++) { uint16_t v = ARR(c)[i]; out[v >> 6] |= 1ull << (v & 63); }
        return;
    }
    for (int i = 0; i < c->len; i++) { /* runToBitmap roaring.go:3792-3856 */
        uint32_t s = RUN(c)[i].start, l = RUN(c)[i].last;
        uint32_t ws = s >> 6, wl = l >> 6;
        uint64_t ms = ~0ull << (s & 63), ml = ~0ull >> (63 - (l & 63));
        if (ws == wl) out[ws] |= ms & ml;
        else { out[ws] |= ms; for (uint32_t k = ws + 1; k < wl; k++) out[k] = ~0ull; out[wl] |= ml; }
    }
}

static fbo_container *words_to_array(const uint64_t *w, int32_t n) { /* bitmapToArray roaring.go:3687 */
    fbo_container *c = c_alloc(FBO_ARRAY, n);
    int k = 0;
    for (int i = 0; i < 1024; i++) { uint64_t v = w[i]; while (v) { int b = __builtin_ctzll(v); ARR(c)[k++] = (uint16_t)(i * 64 + b); v &= v - 1; } }
    c->n = n;
    return c;
}
static fbo_container *words_to_run(const uint64_t *w, int32_t runs, int32_t n) { /* bitmapToRun roaring.go:3859 */
    fbo_container *c = c_alloc(FBO_RUN, runs);
    int k = 0; int inrun = 0; uint32_t start = 0;
    for (uint32_t v = 0; v < 65536; v++) {
        int bit = (w[v >> 6] >> (v & 63)) & 1;
        if (bit && !inrun) { inrun = 1; start = v; }
        else if (!bit && inrun) { inrun = 0; RUN(c)[k].start = (uint16_t)start; RUN(c)[k].last = (uint16_t)(v - 1); k++; }
        if ((v & 63) == 0 && !inrun && w[v >> 6] == 0) v += 63; /* skip empty word */
    }
    if (inrun) { RUN(c)[k].start = (uint16_t)start; RUN(c)[k].last = 65535; k++; }
    c->len = k; c->n = n;
    return c;
}
static int32_t words_count(const uint64_t *w) { int32_t n = 0; for (int i = 0; i < 1024; i++) n += popcnt64(w[i]); return n; }

fbo_container *fbo_c_convert(const fbo_container *c, int typ) {
    uint64_t w[1024];
    if (!c) return NULL;
    if (c->typ == typ) return fbo_c_clone(c);
    fbo_c_to_words(c, w);
    int32_t n = words_count(w);
    if (typ == FBO_BITMAP) { fbo_container *o = fbo_c_bitmap(w); return o; }
    if (typ == FBO_ARRAY) return words_to_array(w, n);
    fbo_container tmp = { FBO_BITMAP, n, 1024, w };
    return words_to_run(w, fbo_c_count_runs(&tmp), n);
}

/* optimize roaring.go:3412-3461 */
fbo_container *fbo_c_optimize(fbo_container *c) {
    if (!c) return NULL;
    if (c->n == 0) { fbo_c_free(c); return NULL; }
    int32_t runs = fbo_c_count_runs(c);
    int newtyp;
    if (runs <= FBO_RUN_MAX_SIZE && runs <= c->n / 2) newtyp = FBO_RUN;
    else if (c->n < FBO_ARRAY_MAX_SIZE) newtyp = FBO_ARRAY;
    else newtyp = FBO_BITMAP;
    if (newtyp == c->typ) return c;
    fbo_container *o = fbo_c_convert(c, newtyp);
    fbo_c_free(c);
    return o;
}

/* BitmapCountRange roaring.go:3092, ArrayCountRange :3074, RunCountRange :3200, countRange :3057 ; [start,end) */
int32_t fbo_c_count_range(const fbo_container *c, int32_t start, int32_t end) {
    if (!c || start >= end) return 0;
    if (end > 65536) end = 65536;
    if (start < 0) start = 0;
    int32_t n = 0;
    if (c->typ == FBO_ARRAY) {
        for (int i = 0; i < c->len; i++) { int32_t v = ARR(c)[i]; if (v >= start && v < end) n++; }
    } else if (c->typ == FBO_RUN) {
        for (int i = 0; i < c->len; i++) {
            int32_t s = RUN(c)[i].start, l = RUN(c)[i].last;
            int32_t lo = s > start ? s : start, hi = l < end - 1 ? l : end - 1;
            if (hi >= lo) n += hi - lo + 1;
        }
    } else {
        const uint64_t *w = BMP(c);
        int32_t i = start >> 6, j = (end - 1) >> 6;
        uint64_t mi = ~0ull << (start & 63), mj = ~0ull >> (63 - ((end - 1) & 63));
        if (i == j) return popcnt64(w[i] & mi & mj);
        n += popcnt64(w[i] & mi);
        for (int k = i + 1; k < j; k++) n += popcnt64(w[k]);
        n += popcnt64(w[j] & mj);
    }
    return n;
}

/* ---- AND-count: intersectionCount roaring.go:4477-4614 ---- */
static int32_t icount_array_array(const fbo_container *a, const fbo_container *b) { /* :4514 */
    const uint16_t *ca = ARR(a), *cb = ARR(b); int na = a->len, nb = b->len;
    if (na > nb) { const uint16_t *t = ca; ca = cb; cb = t; int x = na; na = nb; nb = x; }
    int32_t n = 0; int j = 0;
    if (nb == 0) return 0;
    for (int i = 0; i < na; i++) {
        uint16_t va = ca[i];
        while (cb[j] < va) { j++; if (j >= nb) return n; }
        if (cb[j] == va) n++;
    }
    return n;
}
static int32_t icount_array_run(const fbo_container *a, const fbo_container *b) { /* :4537 */
    int32_t n = 0;
    for (int i = 0, j = 0; i < a->len && j < b->len;) {
        uint16_t va = ARR(a)[i]; fbo_interval vb = RUN(b)[j];
        if (va < vb.start) i++;
        else if (va <= vb.last) { i++; n++; }
        else j++;
    }
    return n;
}
static int32_t icount_run_run(const fbo_container *a, const fbo_container *b) { /* :4555 */
    int32_t n = 0;
    for (int i = 0, j = 0; i < a->len && j < b->len;) {
        fbo_interval va = RUN(a)[i], vb = RUN(b)[j];
        if (va.last < vb.start) i++;
        else if (va.start > vb.last) j++;
        else if (va.last > vb.last && va.start >= vb.start) { n += 1 + (int32_t)(vb.last - va.start); j++; }
        else if (va.last > vb.last && va.start < vb.start) { n += 1 + (int32_t)(vb.last - vb.start); j++; }
        else if (va.last <= vb.last && va.start >= vb.start) { n += 1 + (int32_t)(va.last - va.start); i++; }
        else { n += 1 + (int32_t)(va.last - vb.start); i++; }
    }
    return n;
}
static int32_t icount_bitmap_run(const fbo_container *a, const fbo_container *b) { /* :4588 */
    int32_t n = 0;
    for (int i = 0; i < b->len; i++) n += fbo_c_count_range(a, RUN(b)[i].start, (int32_t)RUN(b)[i].last + 1);
    return n;
}
static int32_t icount_array_bitmap(const fbo_container *a, const fbo_container *b) { /* :4596 */
    int32_t n = 0; const uint64_t *w = BMP(b);
    for (int i = 0; i < a->len; i++) { uint16_t v = ARR(a)[i]; n += (int32_t)((w[v >> 6] >> (v & 63)) & 1); }
    return n;
}
static int32_t icount_bitmap_bitmap(const fbo_container *a, const fbo_container *b) { /* :4611, popcountAndSlice :6928 */
    int32_t n = 0; const uint64_t *x = BMP(a), *y = BMP(b);
    for (int i = 0; i < 1024; i++) n += popcnt64(x[i] & y[i]);
    return n;
}
int32_t fbo_intersection_count(const fbo_container *a, const fbo_container *b) { /* :4477 */
    if (fbo_c_n(a) == 65536) return fbo_c_n(b);
    if (fbo_c_n(b) == 65536) return fbo_c_n(a);
    if (fbo_c_n(a) == 0 || fbo_c_n(b) == 0) return 0;
    if (a->typ == FBO_ARRAY) {
        if (b->typ == FBO_ARRAY) return icount_array_array(a, b);
        if (b->typ == FBO_RUN) return icount_array_run(a, b);
        return icount_array_bitmap(a, b);
    } else if (a->typ == FBO_RUN) {
        if (b->typ == FBO_ARRAY) return icount_array_run(b, a);
        if (b->typ == FBO_RUN) return icount_run_run(a, b);
        return icount_bitmap_run(b, a);
    } else {
        if (b->typ == FBO_ARRAY) return icount_array_bitmap(b, a);
        if (b->typ == FBO_RUN) return icount_bitmap_run(a, b);
        return icount_bitmap_bitmap(a, b);
    }
}

/* run builder: runAppendInterval roaring.go:5159-5181 */
typedef struct { fbo_interval *iv; int len, cap; } runbuf;
static int32_t rb_append(runbuf *r, fbo_interval v) {
    if (r->len == 0) {
        if (r->cap == 0) { r->cap = 16; r->iv = xmalloc(sizeof(fbo_interval) * r->cap); }
        r->iv[r->len++] = v; return (int32_t)(v.last - v.start) + 1;
    }
    fbo_interval last = r->iv[r->len - 1];
    if (last.last == MAXV) return 0;
    if ((uint32_t)last.last + 1 >= v.start && v.last > last.last) { r->iv[r->len - 1].last = v.last; return (int32_t)(v.last - last.last); }
    else if ((uint32_t)last.last + 1 < v.start) {
        if (r->len == r->cap) { r->cap *= 2; r->iv = realloc(r->iv, sizeof(fbo_interval) * r->cap); }
        r->iv[r->len++] = v; return (int32_t)(v.last - v.start) + 1;
    }
    return 0;
}
static fbo_container *rb_finish(runbuf *r, int32_t n) {
    fbo_container *c = xmalloc(sizeof *c);
    c->typ = FBO_RUN; c->n = n; c->len = r->len; c->data = r->iv ? r->iv : xcalloc(1, 8);
    return c;
}
static fbo_container *run_to_array_own(fbo_container *c) { fbo_container *o = fbo_c_convert(c, FBO_ARRAY); fbo_c_free(c); return o; }
static fbo_container *run_to_bitmap_own(fbo_container *c) { fbo_container *o = fbo_c_convert(c, FBO_BITMAP); fbo_c_free(c); return o; }

/* ---- AND: intersect roaring.go:4753-4978 ---- */
static fbo_container *isect_array_array(const fbo_container *a, const fbo_container *b) { /* :4793 */
    fbo_container *o = c_alloc(FBO_ARRAY, a->len); int k = 0;
    for (int i = 0, j = 0; i < a->len && j < b->len;) {
        uint16_t va = ARR(a)[i], vb = ARR(b)[j];
        if (va < vb) i++; else if (va > vb) j++; else { ARR(o)[k++] = va; i++; j++; }
    }
    o->len = k; o->n = k; return o;
}
static fbo_container *isect_array_run(const fbo_container *a, const fbo_container *b) { /* :4815 */
    fbo_container *o = c_alloc(FBO_ARRAY, a->len); int k = 0;
    for (int i = 0, j = 0; i < a->len && j < b->len;) {
        uint16_t va = ARR(a)[i]; fbo_interval vb = RUN(b)[j];
        if (va < vb.start) i++; else if (va > vb.last) j++; else { ARR(o)[k++] = va; i++; }
    }
    o->len = k; o->n = k; return o;
}
static fbo_container *isect_run_run(const fbo_container *a, const fbo_container *b) { /* :4835 */
    runbuf r = {0}; int32_t n = 0;
    for (int i = 0, j = 0; i < a->len && j < b->len;) {
        fbo_interval va = RUN(a)[i], vb = RUN(b)[j];
        if (va.last < vb.start) i++;
        else if (vb.last < va.start) j++;
        else if (va.last > vb.last && va.start >= vb.start) { n += rb_append(&r, (fbo_interval){va.start, vb.last}); j++; }
        else if (va.last > vb.last && va.start < vb.start) { n += rb_append(&r, vb); j++; }
        else if (va.last <= vb.last && va.start >= vb.start) { n += rb_append(&r, va); i++; }
        else { n += rb_append(&r, (fbo_interval){vb.start, va.last}); i++; }
    }
    fbo_container *o = rb_finish(&r, n);
    if (n < FBO_ARRAY_MAX_SIZE && o->len > n / 2) o = run_to_array_own(o);
    else if (o->len > FBO_RUN_MAX_SIZE) o = run_to_bitmap_own(o);
    return o;
}
static fbo_container *isect_bitmap_run(const fbo_container *a, const fbo_container *b) { /* :4879 */
    uint64_t rw[1024], ow[1024];
    fbo_c_to_words(b, rw);
    for (int i = 0; i < 1024; i++) ow[i] = BMP(a)[i] & rw[i];
    int32_t n = words_count(ow);
    if (b->n <= FBO_ARRAY_MAX_SIZE) return words_to_array(ow, n);
    return fbo_c_bitmap(ow);
}
static fbo_container *isect_array_bitmap(const fbo_container *a, const fbo_container *b) { /* :4944 */
    fbo_container *o = c_alloc(FBO_ARRAY, a->len); int k = 0; const uint64_t *w = BMP(b);
    for (int i = 0; i < a->len; i++) { uint16_t v = ARR(a)[i]; if ((w[v >> 6] >> (v & 63)) & 1) ARR(o)[k++] = v; }
    o->len = k; o->n = k; return o;
}
static fbo_container *isect_bitmap_bitmap(const fbo_container *a, const fbo_container *b) { /* :4960: always bitmap */
    fbo_container *o = c_alloc(FBO_BITMAP, 1024); int32_t n = 0;
    for (int i = 0; i < 1024; i++) { BMP(o)[i] = BMP(a)[i] & BMP(b)[i]; n += popcnt64(BMP(o)[i]); }
    o->n = n; return o;
}
fbo_container *fbo_intersect(const fbo_container *a, const fbo_container *b) { /* :4753 */
    if (fbo_c_n(a) == 65536) return fbo_c_clone(b);
    if (fbo_c_n(b) == 65536) return fbo_c_clone(a);
    if (fbo_c_n(a) == 0 || fbo_c_n(b) == 0) return NULL;
    if (a->typ == FBO_ARRAY) {
        if (b->typ == FBO_ARRAY) return isect_array_array(a, b);
        if (b->typ == FBO_RUN) return isect_array_run(a, b);
        return isect_array_bitmap(a, b);
    } else if (a->typ == FBO_RUN) {
        if (b->typ == FBO_ARRAY) return isect_array_run(b, a);
        if (b->typ == FBO_RUN) return isect_run_run(a, b);
        return isect_bitmap_run(b, a);
    } else {
        if (b->typ == FBO_ARRAY) return isect_array_bitmap(b, a);
        if (b->typ == FBO_RUN) return isect_bitmap_run(a, b);
        return isect_bitmap_bitmap(a, b);
    }
}

/* ---- OR: union roaring.go:4980-5231, 5424-5493 ---- */
static fbo_container *full_container(void) { fbo_interval f = {0, 65535}; return fbo_c_run(&f, 1); } /* fullContainer :67 */

static fbo_container *union_array_array(const fbo_container *a, const fbo_container *b) { /* :5016; may exceed 4096 */
    if (a->n == 0) return fbo_c_clone(b);
    if (b->n == 0) return fbo_c_clone(a);
    fbo_container *o = c_alloc(FBO_ARRAY, a->len + b->len); int k = 0, i = 0, j = 0;
    while (i < a->len && j < b->len) {
        uint16_t va = ARR(a)[i], vb = ARR(b)[j];
        if (va < vb) { ARR(o)[k++] = va; i++; } else if (va > vb) { ARR(o)[k++] = vb; j++; } else { ARR(o)[k++] = va; i++; j++; }
    }
    while (i < a->len) ARR(o)[k++] = ARR(a)[i++];
    while (j < b->len) ARR(o)[k++] = ARR(b)[j++];
    o->len = k; o->n = k; return o;
}
static fbo_container *union_array_run(const fbo_container *a, const fbo_container *b) { /* :5120 */
    runbuf r = {0}; int32_t n = 0;
    for (int i = 0, j = 0; i < a->len || j < b->len;) {
        uint16_t va = i < a->len ? ARR(a)[i] : 0; fbo_interval vb = j < b->len ? RUN(b)[j] : (fbo_interval){0, 0};
        if (i < a->len && (j >= b->len || va < vb.start)) { n += rb_append(&r, (fbo_interval){va, va}); i++; }
        else { n += rb_append(&r, vb); j++; }
    }
    fbo_container *o = rb_finish(&r, n);
    if (n < FBO_ARRAY_MAX_SIZE) o = run_to_array_own(o);
    else if (o->len > FBO_RUN_MAX_SIZE) o = run_to_bitmap_own(o);
    return o;
}
static fbo_container *union_run_run(const fbo_container *a, const fbo_container *b) { /* :5182 */
    runbuf r = {0}; int32_t n = 0;
    for (int i = 0, j = 0; i < a->len || j < b->len;) {
        fbo_interval va = i < a->len ? RUN(a)[i] : (fbo_interval){0, 0}, vb = j < b->len ? RUN(b)[j] : (fbo_interval){0, 0};
        if (i < a->len && (j >= b->len || va.start < vb.start)) { n += rb_append(&r, va); i++; }
        else { n += rb_append(&r, vb); j++; }
    }
    fbo_container *o = rb_finish(&r, n);
    if (o->len > FBO_RUN_MAX_SIZE) o = run_to_bitmap_own(o);
    return o;
}
static fbo_container *union_via_bitmap(const fbo_container *a, const fbo_container *b) { /* unionBitmapRun :5211, ArrayBitmap :5424, BitmapBitmap :5450 -> bitmap */
    uint64_t x[1024], y[1024];
    fbo_c_to_words(a, x); fbo_c_to_words(b, y);
    for (int i = 0; i < 1024; i++) x[i] |= y[i];
    return fbo_c_bitmap(x);
}
fbo_container *fbo_union(const fbo_container *a, const fbo_container *b) { /* :4980 */
    if (fbo_c_n(a) == 65536 || fbo_c_n(b) == 65536) return full_container();
    if (!a) return fbo_c_clone(b);
    if (!b) return fbo_c_clone(a);
    if (a->typ == FBO_ARRAY) {
        if (b->typ == FBO_ARRAY) return union_array_array(a, b);
        if (b->typ == FBO_RUN) return union_array_run(a, b);
        return union_via_bitmap(a, b);
    } else if (a->typ == FBO_RUN) {
        if (b->typ == FBO_ARRAY) return union_array_run(b, a);
        if (b->typ == FBO_RUN) return union_run_run(a, b);
        return union_via_bitmap(b, a);
    }
    return union_via_bitmap(a, b);
}

/* ---- ANDNOT: difference roaring.go:5692-6050. Result encodings follow SURVEY Appendix B. ---- */
static fbo_container *diff_array_array(const fbo_container *a, const fbo_container *b) { /* :5730 */
    fbo_container *o = c_alloc(FBO_ARRAY, a->len); int k = 0, i = 0, j = 0;
    while (i < a->len) {
        if (j >= b->len) { ARR(o)[k++] = ARR(a)[i++]; continue; }
        uint16_t va = ARR(a)[i], vb = ARR(b)[j];
        if (va < vb) { ARR(o)[k++] = va; i++; } else if (va > vb) j++; else { i++; j++; }
    }
    o->len = k; o->n = k; return o;
}
static fbo_container *diff_array_other(const fbo_container *a, const fbo_container *b) { /* ArrayRun :5757, ArrayBitmap :5991 -> array */
    fbo_container *o = c_alloc(FBO_ARRAY, a->len); int k = 0;
    if (b->typ == FBO_RUN) {
        int j = 0;
        for (int i = 0; i < a->len; i++) {
            uint16_t va = ARR(a)[i];
            while (j < b->len && RUN(b)[j].last < va) j++;
            if (j < b->len && va >= RUN(b)[j].start) continue;
            ARR(o)[k++] = va;
        }
    } else {
        const uint64_t *w = BMP(b);
        for (int i = 0; i < a->len; i++) { uint16_t v = ARR(a)[i]; if (!((w[v >> 6] >> (v & 63)) & 1)) ARR(o)[k++] = v; }
    }
    o->len = k; o->n = k; return o;
}
fbo_container *fbo_difference(const fbo_container *a, const fbo_container *b) { /* :5692 */
    if (fbo_c_n(a) == 0 || fbo_c_n(b) == 65536) return NULL;
    if (fbo_c_n(b) == 0) return fbo_c_clone(a);
    if (a->typ == FBO_ARRAY) {
        if (b->typ == FBO_ARRAY) return diff_array_array(a, b);
        return diff_array_other(a, b);
    }
    uint64_t x[1024], y[1024];
    fbo_c_to_words(a, x); fbo_c_to_words(b, y);
    for (int i = 0; i < 1024; i++) x[i] &= ~y[i];
    int32_t n = words_count(x);
    fbo_container tmp = { FBO_BITMAP, n, 1024, x };
    if (a->typ == FBO_BITMAP) {
        if (b->typ == FBO_RUN) return fbo_c_bitmap(x);           /* BitmapRun :5802: bitmap, no down-convert */
        if (n < FBO_ARRAY_MAX_SIZE) return words_to_array(x, n);  /* BitmapArray :6008-6021, BitmapBitmap :6027-6046 */
        return fbo_c_bitmap(x);
    }
    /* a is run */
    if (b->typ == FBO_ARRAY) { /* RunArray :5813 -> run then optimize() :5860 */
        fbo_container *o = words_to_run(x, fbo_c_count_runs(&tmp), n);
        return fbo_c_optimize(o);
    }
    if (b->typ == FBO_RUN) { /* RunRun :5931 -> run */
        return words_to_run(x, fbo_c_count_runs(&tmp), n);
    }
    /* RunBitmap :5866-5928 */
    if (a->n == 65536) return fbo_c_bitmap(x); /* flipBitmap(b) :4239 */
    int32_t runs = fbo_c_count_runs(&tmp);
    if (runs > FBO_RUN_MAX_SIZE) return fbo_c_bitmap(x);
    if (n < FBO_ARRAY_MAX_SIZE && runs > n / 2) return words_to_array(x, n);
    return words_to_run(x, runs, n);
}

/* ---- XOR: xor roaring.go:6052-6179, 6607-6827 ---- */
static fbo_container *xor_array_array(const fbo_container *a, const fbo_container *b) { /* :6089 */
    fbo_container *o = c_alloc(FBO_ARRAY, a->len + b->len); int k = 0, i = 0, j = 0;
    while (i < a->len && j < b->len) {
        uint16_t va = ARR(a)[i], vb = ARR(b)[j];
        if (va < vb) { ARR(o)[k++] = va; i++; } else if (va > vb) { ARR(o)[k++] = vb; j++; } else { i++; j++; }
    }
    while (i < a->len) ARR(o)[k++] = ARR(a)[i++];
    while (j < b->len) ARR(o)[k++] = ARR(b)[j++];
    o->len = k; o->n = k; return o;
}
fbo_container *fbo_xor(const fbo_container *a, const fbo_container *b) { /* :6052 */
    if (fbo_c_n(a) == 0) return fbo_c_clone(fbo_c_n(b) ? b : NULL);
    if (fbo_c_n(b) == 0) return fbo_c_clone(a);
    if (a->typ == FBO_ARRAY && b->typ == FBO_ARRAY) return xor_array_array(a, b);
    uint64_t x[1024], y[1024];
    fbo_c_to_words(a, x); fbo_c_to_words(b, y);
    for (int i = 0; i < 1024; i++) x[i] ^= y[i];
    int32_t n = words_count(x);
    fbo_container tmp = { FBO_BITMAP, n, 1024, x };
    int ta = a->typ, tb = b->typ;
    if ((ta == FBO_BITMAP && tb == FBO_RUN) || (ta == FBO_RUN && tb == FBO_BITMAP)) return fbo_c_bitmap(x); /* xorBitmapRun :6816 */
    if (ta == FBO_BITMAP || tb == FBO_BITMAP) { /* xorArrayBitmap :6135-6148, xorBitmapBitmap :6155-6175 */
        if (n < FBO_ARRAY_MAX_SIZE) return words_to_array(x, n);
        return fbo_c_bitmap(x);
    }
    int32_t runs = fbo_c_count_runs(&tmp);
    if (ta == FBO_RUN && tb == FBO_RUN) { /* xorRunRun :6769-6811 */
        if (n < FBO_ARRAY_MAX_SIZE && runs > n / 2) return words_to_array(x, n);
        if (runs > FBO_RUN_MAX_SIZE) return fbo_c_bitmap(x);
        return words_to_run(x, runs, n);
    }
    /* xorArrayRun :6607-6670 */
    if (n < FBO_ARRAY_MAX_SIZE) return words_to_array(x, n);
    if (runs > FBO_RUN_MAX_SIZE) return fbo_c_bitmap(x);
    return words_to_run(x, runs, n);
}

/* flip roaring.go:4221-4264 (whole-container complement) */
fbo_container *fbo_flip(const fbo_container *a) {
    uint64_t x[1024];
    fbo_c_to_words(a, x);
    for (int i = 0; i < 1024; i++) x[i] = ~x[i];
    int32_t n = words_count(x);
    if (!a || a->typ == FBO_BITMAP) return fbo_c_bitmap(x);
    fbo_container tmp = { FBO_BITMAP, n, 1024, x };
    if (a->typ == FBO_ARRAY) return words_to_array(x, n);
    return words_to_run(x, fbo_c_count_runs(&tmp), n);
}

/* ------------------------------------------------------------------ bitmaps */

fbo_bitmap *fbo_b_new(void) { return xcalloc(1, sizeof(fbo_bitmap)); }
void fbo_b_free(fbo_bitmap *b) {
    if (!b) return;
    if (!b->view) for (int64_t i = 0; i < b->n; i++) fbo_c_free(b->cs[i]);
    free(b->keys); free(b->cs); free(b);
}
static void b_reserve(fbo_bitmap *b, int64_t need) {
    if (need <= b->cap) return;
    int64_t nc = b->cap ? b->cap * 2 : 16; if (nc < need) nc = need;
    b->keys = realloc(b->keys, sizeof(uint64_t) * nc); b->cs = realloc(b->cs, sizeof(void *) * nc); b->cap = nc;
}
static int64_t b_search(const fbo_bitmap *b, uint64_t key) { /* search64 roaring.go:6528; returns index or -(ins+1) */
    int64_t lo = 0, hi = b->n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (b->keys[m] < key) lo = m + 1; else hi = m; }
    if (lo < b->n && b->keys[lo] == key) return lo;
    return -(lo + 1);
}
static void b_append(fbo_bitmap *b, uint64_t key, fbo_container *c) { /* key must exceed all existing */
    if (!c) return;
    b_reserve(b, b->n + 1); b->keys[b->n] = key; b->cs[b->n] = c; b->n++;
}
void fbo_b_put(fbo_bitmap *b, uint64_t key, fbo_container *c) {
    int64_t i = b_search(b, key);
    if (i >= 0) { fbo_c_free(b->cs[i]); b->cs[i] = c; return; }
    i = -(i + 1);
    b_reserve(b, b->n + 1);
    memmove(b->keys + i + 1, b->keys + i, sizeof(uint64_t) * (b->n - i));
    memmove(b->cs + i + 1, b->cs + i, sizeof(void *) * (b->n - i));
    b->keys[i] = key; b->cs[i] = c; b->n++;
}
const fbo_container *fbo_b_get(const fbo_bitmap *b, uint64_t key) { int64_t i = b_search(b, key); return i >= 0 ? b->cs[i] : NULL; }
fbo_bitmap *fbo_b_clone(const fbo_bitmap *a) {
    fbo_bitmap *o = fbo_b_new();
    for (int64_t i = 0; i < a->n; i++) b_append(o, a->keys[i], fbo_c_clone(a->cs[i]));
    return o;
}

/* Container.add semantics (roaring.go:3527-3626): arrays grow to ArrayMaxSize then become bitmaps; runs extend */
static fbo_container *c_add(fbo_container *c, uint16_t v, int *changed) {
    *changed = 0;
    if (!c) { *changed = 1; return fbo_c_array(&v, 1); }
    if (fbo_c_contains(c, v)) return c;
    *changed = 1;
    if (c->typ == FBO_BITMAP) { BMP(c)[v >> 6] |= 1ull << (v & 63); c->n++; return c; }
    if (c->typ == FBO_RUN) { fbo_container *o = fbo_c_convert(c, c->n < FBO_ARRAY_MAX_SIZE ? FBO_ARRAY : FBO_BITMAP); fbo_c_free(c); int ch; return c_add(o, v, &ch); }
    if (c->n >= FBO_ARRAY_MAX_SIZE) { fbo_container *o = fbo_c_convert(c, FBO_BITMAP); fbo_c_free(c); int ch; return c_add(o, v, &ch); }
    int lo = 0, hi = c->len;
    while (lo < hi) { int m = (lo + hi) >> 1; if (ARR(c)[m] < v) lo = m + 1; else hi = m; }
    c->data = realloc(c->data, (size_t)(c->len + 1) * 2);
    memmove(ARR(c) + lo + 1, ARR(c) + lo, (size_t)(c->len - lo) * 2);
    ARR(c)[lo] = v; c->len++; c->n++;
    return c;
}
int fbo_b_add(fbo_bitmap *b, uint64_t v) { /* DirectAdd roaring.go:367 */
    uint64_t key = v >> 16; int ch;
    int64_t i = b_search(b, key);
    if (i >= 0) { b->cs[i] = c_add(b->cs[i], (uint16_t)v, &ch); return ch; }
    uint16_t lo = (uint16_t)v;
    fbo_b_put(b, key, fbo_c_array(&lo, 1));
    return 1;
}
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }
void fbo_b_add_many(fbo_bitmap *b, const uint64_t *vals, int64_t n) { /* bulk: sort, group by key, merge with union */
    if (n <= 0) return;
    uint64_t *s = xmalloc(sizeof(uint64_t) * n); memcpy(s, vals, sizeof(uint64_t) * n);
    qsort(s, n, sizeof(uint64_t), cmp_u64);
    uint16_t *tmp = xmalloc(65536 * 2);
    for (int64_t i = 0; i < n;) {
        uint64_t key = s[i] >> 16; int k = 0; int64_t j = i;
        while (j < n && (s[j] >> 16) == key) { uint16_t lo = (uint16_t)s[j]; if (k == 0 || tmp[k - 1] != lo) tmp[k++] = lo; j++; }
        fbo_container *nc = fbo_c_array(tmp, k);
        int64_t at = b_search(b, key);
        if (at >= 0) { fbo_container *u = fbo_union(b->cs[at], nc); fbo_c_free(nc); fbo_c_free(b->cs[at]); b->cs[at] = u; }
        else fbo_b_put(b, key, nc);
        i = j;
    }
    free(tmp); free(s);
}
int fbo_b_contains(const fbo_bitmap *b, uint64_t v) { return fbo_c_contains(fbo_b_get(b, v >> 16), (uint16_t)v); }
uint64_t fbo_b_count(const fbo_bitmap *b) { uint64_t n = 0; for (int64_t i = 0; i < b->n; i++) n += (uint64_t)fbo_c_n(b->cs[i]); return n; } /* containers_slice.go:120 */
int fbo_b_any(const fbo_bitmap *b) { for (int64_t i = 0; i < b->n; i++) if (fbo_c_n(b->cs[i]) > 0) return 1; return 0; } /* roaring.go:547 */
uint64_t fbo_b_slice(const fbo_bitmap *b, uint64_t *out, uint64_t cap) { /* Slice via Iterator roaring.go:2815 */
    uint64_t k = 0; uint64_t w[1024];
    for (int64_t i = 0; i < b->n; i++) {
        if (!b->cs[i] || b->cs[i]->n == 0) continue;
        fbo_c_to_words(b->cs[i], w);
        for (int j = 0; j < 1024; j++) { uint64_t v = w[j]; while (v) { int bit = __builtin_ctzll(v); if (k < cap) out[k] = (b->keys[i] << 16) | (uint64_t)(j * 64 + bit); k++; v &= v - 1; } }
    }
    return k;
}

/* Bitmap.Intersect roaring.go:736-759 */
fbo_bitmap *fbo_b_intersect(const fbo_bitmap *a, const fbo_bitmap *b) {
    fbo_bitmap *o = fbo_b_new(); int64_t i = 0, j = 0;
    while (i < a->n && j < b->n) {
        if (a->keys[i] < b->keys[j]) i++; else if (a->keys[i] > b->keys[j]) j++;
        else { b_append(o, a->keys[i], fbo_intersect(a->cs[i], b->cs[j])); i++; j++; }
    }
    return o;
}
/* Bitmap.IntersectionCount roaring.go:711-733 */
uint64_t fbo_b_intersection_count(const fbo_bitmap *a, const fbo_bitmap *b) {
    uint64_t n = 0; int64_t i = 0, j = 0;
    while (i < a->n && j < b->n) {
        if (a->keys[i] < b->keys[j]) i++; else if (a->keys[i] > b->keys[j]) j++;
        else { n += (uint64_t)fbo_intersection_count(a->cs[i], b->cs[j]); i++; j++; }
    }
    return n;
}
/* unionIntoTargetSingle roaring.go:1292-1315 */
fbo_bitmap *fbo_b_union(const fbo_bitmap *a, const fbo_bitmap *b) {
    fbo_bitmap *o = fbo_b_new(); int64_t i = 0, j = 0;
    while (i < a->n || j < b->n) {
        if (i < a->n && (j >= b->n || a->keys[i] < b->keys[j])) { b_append(o, a->keys[i], fbo_c_clone(a->cs[i])); i++; }
        else if (j < b->n && (i >= a->n || a->keys[i] > b->keys[j])) { b_append(o, b->keys[j], fbo_c_clone(b->cs[j])); j++; }
        else { b_append(o, a->keys[i], fbo_union(a->cs[i], b->cs[j])); i++; j++; }
    }
    return o;
}
/* Container.unionInPlace into a bitmap target: unionBitmapArrayInPlace roaring.go:5440 (sets the bits), unionBitmapBitmapInPlace
 * :5473 (ORs the words), unionBitmapRunInPlace :5222 (bitmapSetRangeIgnoreN per run) */
static void c_or_into_words(const fbo_container *c, uint64_t *acc) {
    if (!c || c->n == 0) return;
    if (c->typ == FBO_ARRAY) { const uint16_t *a = ARR(c); for (int i = 0; i < c->len; i++) acc[a[i] >> 6] |= 1ull << (a[i] & 63); return; }
    if (c->typ == FBO_BITMAP) { const uint64_t *w = BMP(c); for (int q = 0; q < 1024; q++) acc[q] |= w[q]; return; }
    for (int i = 0; i < c->len; i++) {
        uint32_t s = RUN(c)[i].start, l = RUN(c)[i].last, ws = s >> 6, wl = l >> 6;
        uint64_t ms = ~0ull << (s & 63), ml = ~0ull >> (63 - (l & 63));
        if (ws == wl) acc[ws] |= ms & ml;
        else { acc[ws] |= ms; for (uint32_t k = ws + 1; k < wl; k++) acc[k] = ~0ull; acc[wl] |= ml; }
    }
}
/* n-ary Bitmap.Union -> unionInPlace roaring.go:1272-1284,1410-1561.  Per key: any full => fullContainer;
 * single source => reuse; expectedN >= 512 => accumulate in a bitmap container, popcount once (Repair :1560);
 * else left-fold union(). */
fbo_bitmap *fbo_b_union_n(const fbo_bitmap *a, const fbo_bitmap *const *others, int n) {
    if (n == 1) return fbo_b_union(a, others[0]);
    int total = n + 1;
    const fbo_bitmap **src = xmalloc(sizeof(void *) * total); int64_t *pos = xcalloc(total, sizeof(int64_t));
    src[0] = a; for (int i = 0; i < n; i++) src[i + 1] = others[i];
    fbo_bitmap *o = fbo_b_new();
    uint64_t acc[1024];
    for (;;) {
        uint64_t key = ~0ull; int have = 0;
        for (int s = 0; s < total; s++) if (pos[s] < src[s]->n) { uint64_t k = src[s]->keys[pos[s]]; if (!have || k < key) { key = k; have = 1; } }
        if (!have) break;
        int cnt = 0, full = 0; int64_t expected = 0; const fbo_container *only = NULL;
        for (int s = 0; s < total; s++) if (pos[s] < src[s]->n && src[s]->keys[pos[s]] == key) {
            const fbo_container *c = src[s]->cs[pos[s]];
            if (fbo_c_n(c) == 65536) full = 1;
            if (fbo_c_n(c) > 0) { cnt++; only = c; expected += c->n; }
        }
        fbo_container *res = NULL;
        if (full) res = full_container();
        else if (cnt == 1) res = fbo_c_clone(only);
        else if (cnt > 1 && expected >= 512) {
            memset(acc, 0, sizeof acc);          /* NewContainerBitmapN(nil, 0), then Container.unionInPlace per source */
            for (int s = 0; s < total; s++) if (pos[s] < src[s]->n && src[s]->keys[pos[s]] == key) c_or_into_words(src[s]->cs[pos[s]], acc);
            res = fbo_c_bitmap(acc);             /* Repair :1560: one popcount per target container */
        } else if (cnt > 1) {
            for (int s = 0; s < total; s++) if (pos[s] < src[s]->n && src[s]->keys[pos[s]] == key) {
                fbo_container *u = fbo_union(res, src[s]->cs[pos[s]]); fbo_c_free(res); res = u;
            }
        }
        b_append(o, key, res);
        for (int s = 0; s < total; s++) if (pos[s] < src[s]->n && src[s]->keys[pos[s]] == key) pos[s]++;
    }
    free(src); free(pos);
    return o;
}
/* Bitmap.Difference roaring.go:1564-1595 */
fbo_bitmap *fbo_b_difference(const fbo_bitmap *a, const fbo_bitmap *b) {
    fbo_bitmap *o = fbo_b_new(); int64_t i = 0, j = 0;
    while (i < a->n) {
        if (j >= b->n || a->keys[i] < b->keys[j]) { b_append(o, a->keys[i], fbo_c_clone(a->cs[i])); i++; }
        else if (a->keys[i] > b->keys[j]) j++;
        else { b_append(o, a->keys[i], fbo_difference(a->cs[i], b->cs[j])); i++; j++; }
    }
    return o;
}
/* Bitmap.Xor roaring.go:1598-1623 */
fbo_bitmap *fbo_b_xor(const fbo_bitmap *a, const fbo_bitmap *b) {
    fbo_bitmap *o = fbo_b_new(); int64_t i = 0, j = 0;
    while (i < a->n || j < b->n) {
        if (i < a->n && (j >= b->n || a->keys[i] < b->keys[j])) { b_append(o, a->keys[i], fbo_c_clone(a->cs[i])); i++; }
        else if (j < b->n && (i >= a->n || a->keys[i] > b->keys[j])) { b_append(o, b->keys[j], fbo_c_clone(b->cs[j])); j++; }
        else { b_append(o, a->keys[i], fbo_xor(a->cs[i], b->cs[j])); i++; j++; }
    }
    return o;
}
void fbo_b_optimize(fbo_bitmap *b) { /* Bitmap.Optimize roaring.go:1706-1720 (drops empties) */
    int64_t k = 0;
    for (int64_t i = 0; i < b->n; i++) { fbo_container *c = fbo_c_optimize(b->cs[i]); if (c) { b->keys[k] = b->keys[i]; b->cs[k] = c; k++; } }
    b->n = k;
}
/* OffsetRange roaring.go:678-701: containers with key in [start>>16, end>>16) re-keyed to offset>>16 + (key - start>>16) */
fbo_bitmap *fbo_b_offset_range(const fbo_bitmap *b, uint64_t offset, uint64_t start, uint64_t end) {
    fbo_bitmap *o = fbo_b_new();
    uint64_t off = offset >> 16, hi0 = start >> 16, hi1 = end >> 16;
    int64_t i = b_search(b, hi0); if (i < 0) i = -(i + 1);
    for (; i < b->n && b->keys[i] < hi1; i++) if (fbo_c_n(b->cs[i]) > 0) b_append(o, off + (b->keys[i] - hi0), fbo_c_clone(b->cs[i]));
    return o;
}

/* the same range as a view: keys re-based, containers borrowed (what the storage layer returns: frozen containers) */
static fbo_bitmap *b_offset_range_view(const fbo_bitmap *b, uint64_t offset, uint64_t start, uint64_t end) {
    fbo_bitmap *o = fbo_b_new(); o->view = 1;
    uint64_t off = offset >> 16, hi0 = start >> 16, hi1 = end >> 16;
    int64_t i = b_search(b, hi0); if (i < 0) i = -(i + 1);
    for (; i < b->n && b->keys[i] < hi1; i++) if (fbo_c_n(b->cs[i]) > 0) b_append(o, off + (b->keys[i] - hi0), b->cs[i]);
    return o;
}

/* ---- serialisation: writeToUnoptimized roaring.go:1738-1817; container WriteTo :3628-3685 ---- */
static void put16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static void put64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static uint16_t get16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }
static uint32_t get32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint64_t get64(const uint8_t *p) { return (uint64_t)get32(p) | (uint64_t)get32(p + 4) << 32; }
static uint64_t c_size(const fbo_container *c) { return c->typ == FBO_ARRAY ? (uint64_t)c->len * 2 : c->typ == FBO_BITMAP ? 8192 : 2 + (uint64_t)c->len * 4; }

uint64_t fbo_b_write(fbo_bitmap *b, uint8_t *out, uint64_t cap, int optimize) {
    if (optimize) fbo_b_optimize(b);
    uint64_t cnt = 0, payload = 0;
    for (int64_t i = 0; i < b->n; i++) if (fbo_c_n(b->cs[i]) > 0) { cnt++; payload += c_size(b->cs[i]); }
    uint64_t need = 8 + cnt * 16 + payload;
    if (need > cap || !out) return need;
    put32(out, 12348u); put32(out + 4, (uint32_t)cnt);
    uint8_t *h = out + 8, *offp = out + 8 + cnt * 12; uint64_t off = 8 + cnt * 16;
    for (int64_t i = 0; i < b->n; i++) {
        const fbo_container *c = b->cs[i]; if (fbo_c_n(c) == 0) continue;
        put64(h, b->keys[i]); put16(h + 8, c->typ); put16(h + 10, (uint16_t)(c->n - 1)); h += 12;
        put32(offp, (uint32_t)off); offp += 4;
        uint8_t *p = out + off;
        if (c->typ == FBO_ARRAY) for (int k = 0; k < c->len; k++) put16(p + 2 * k, ARR(c)[k]);
        else if (c->typ == FBO_BITMAP) for (int k = 0; k < 1024; k++) put64(p + 8 * k, BMP(c)[k]);
        else { put16(p, (uint16_t)c->len); for (int k = 0; k < c->len; k++) { put16(p + 2 + 4 * k, RUN(c)[k].start); put16(p + 4 + 4 * k, RUN(c)[k].last); } }
        off += c_size(c);
    }
    return need;
}

/* pilosaRoaringIterator roaring.go:2124-2178 ; officialRoaringIterator :2194-2260, header parse :6943-7006 */
static fbo_bitmap *b_read_containers(const uint8_t *buf, uint64_t len, uint64_t *end_out) {
    if (len < 8) return NULL;
    uint32_t magic = get16(buf);
    fbo_bitmap *o = fbo_b_new();
    *end_out = 0; /* lastDataOffset == 0: Remaining() has nothing (roaring.go:2103-2108) */
    if (magic == 12348) {
        if (buf[2] != 0) { fbo_b_free(o); return NULL; } /* storageVersion 0 */
        uint64_t keys = get32(buf + 4);
        if (8 + keys * 16 > len) { fbo_b_free(o); return NULL; }
        const uint8_t *hdr = buf + 8, *offs = buf + 8 + keys * 12;
        uint64_t chunk = 0; uint32_t prev = 0;
        *end_out = keys ? 8 + keys * 16 : (len > 8 ? 8 : 0); /* roaring.go:1994-2001, 2016-2017 */
        for (uint64_t i = 0; i < keys; i++) {
            uint64_t key = get64(hdr + i * 12); uint16_t typ = get16(hdr + i * 12 + 8); int32_t n = (int32_t)get16(hdr + i * 12 + 10) + 1;
            uint32_t o32 = get32(offs + i * 4); if (o32 < prev) chunk += 1ull << 32; prev = o32;
            uint64_t off = chunk + o32; fbo_container *c = NULL;
            if (typ == FBO_ARRAY) {
                if (off + (uint64_t)n * 2 > len) goto bad;
                c = c_alloc(FBO_ARRAY, n); for (int k = 0; k < n; k++) ARR(c)[k] = get16(buf + off + 2 * k); c->n = n;
                *end_out = off + (uint64_t)n * 2;
            } else if (typ == FBO_BITMAP) {
                if (off + 8192 > len) goto bad;
                c = c_alloc(FBO_BITMAP, 1024); for (int k = 0; k < 1024; k++) BMP(c)[k] = get64(buf + off + 8 * k); c->n = n;
                *end_out = off + 8192;
            } else if (typ == FBO_RUN) {
                if (off + 2 > len) goto bad;
                int rc = get16(buf + off); if (off + 2 + (uint64_t)rc * 4 > len) goto bad;
                c = c_alloc(FBO_RUN, rc); for (int k = 0; k < rc; k++) { RUN(c)[k].start = get16(buf + off + 2 + 4 * k); RUN(c)[k].last = get16(buf + off + 4 + 4 * k); } c->n = n;
                *end_out = off + 2 + (uint64_t)rc * 4;
            } else goto bad;
            fbo_b_put(o, key, c); /* Containers.Put unmarshal_binary.go:58: a repeated key replaces, order does not matter */
        }
        return o;
    }
    if (magic == 12346 || magic == 12347) {
        if (magic == 12346 && get32(buf) != 12346) goto bad; /* readOfficialHeader :6966: the no-run cookie is compared on 32 bits */
        uint64_t keys, pos; const uint8_t *runbits = NULL; int have_runs = magic == 12347;
        if (have_runs) { keys = (uint64_t)get16(buf + 2) + 1; pos = 4; runbits = buf + pos; pos += (keys + 7) / 8; if (pos > len) goto bad; }
        else { keys = get32(buf + 4); pos = 8; }
        if (keys > (1u << 16)) goto bad; /* :6992 */
        if (pos + keys * 4 >= len) goto bad; /* readOfficialHeader roaring.go:7000: '>=' => zero containers is an error */
        const uint8_t *hdr = buf + pos; pos += keys * 4;
        const uint8_t *offs = NULL;
        /* roaring.go:1966-1976: offsets are only used with the no-run cookie; with runs the reference reads
         * payloads sequentially right after the key/cardinality header */
        if (!have_runs) { if (pos + keys * 4 > len) goto bad; offs = buf + pos; }
        uint64_t cur = pos;
        for (uint64_t i = 0; i < keys; i++) {
            uint64_t key = get16(hdr + i * 4); int32_t n = (int32_t)get16(hdr + i * 4 + 2) + 1;
            int isrun = have_runs && ((runbits[i / 8] >> (i % 8)) & 1);
            uint64_t off = (!have_runs && offs) ? get32(offs + i * 4) : cur;
            fbo_container *c;
            if (isrun) {
                if (off + 2 > len) goto bad;
                int rc = get16(buf + off); if (off + 2 + (uint64_t)rc * 4 > len) goto bad;
                c = c_alloc(FBO_RUN, rc);
                for (int k = 0; k < rc; k++) { uint16_t s = get16(buf + off + 2 + 4 * k), l = get16(buf + off + 4 + 4 * k); RUN(c)[k].start = s; RUN(c)[k].last = (uint16_t)(s + l); }
                c->n = n; cur = off + 2 + (uint64_t)rc * 4;
            } else if (n < FBO_ARRAY_MAX_SIZE) { /* containerTyper roaring.go:6954-6960 */
                if (off + (uint64_t)n * 2 > len) goto bad;
                c = c_alloc(FBO_ARRAY, n); for (int k = 0; k < n; k++) ARR(c)[k] = get16(buf + off + 2 * k); c->n = n; cur = off + (uint64_t)n * 2;
            } else {
                if (off + 8192 > len) goto bad;
                c = c_alloc(FBO_BITMAP, 1024); for (int k = 0; k < 1024; k++) BMP(c)[k] = get64(buf + off + 8 * k); c->n = n; cur = off + 8192;
            }
            fbo_b_put(o, key, c);
        }
        *end_out = cur;
        return o;
    }
bad:
    fbo_b_free(o);
    return NULL;
}

static uint32_t fnv32a(uint32_t h, const uint8_t *p, uint64_t n) { for (uint64_t i = 0; i < n; i++) { h ^= p[i]; h *= 16777619u; } return h; }
static fbo_bitmap *b_from_values(const uint8_t *p, uint64_t n) {
    fbo_bitmap *x = fbo_b_new();
    for (uint64_t i = 0; i < n; i++) fbo_b_add(x, get64(p + 8 * i));
    return x;
}

/* Bitmap.UnmarshalBinary unmarshal_binary.go:17-95: the containers, then the ops log up to the end of the data — each op
 * decoded and checksummed as op.UnmarshalBinary does (roaring.go:6375-6427) and applied as op.apply does (:6303-6322).
 * Removal is written as a difference with the removed set, which is what remove() / DirectRemoveN / ImportRoaringBits(clear)
 * leave in the bitmap. */
fbo_bitmap *fbo_b_read(const uint8_t *buf, uint64_t len) {
    uint64_t pos = 0;
    fbo_bitmap *o = b_read_containers(buf, len, &pos);
    if (!o || pos == 0) return o;
    while (pos < len) {
        const uint8_t *d = buf + pos; uint64_t left = len - pos, size;
        if (left < 13) goto bad;
        uint8_t typ = d[0]; uint64_t value = get64(d + 1);
        uint32_t h = fnv32a(2166136261u, d, 9);
        fbo_bitmap *arg = NULL;
        if (typ <= 1) { size = 13; arg = b_from_values(d + 1, 1); }
        else if (typ <= 3) {
            if (value > (1ull << 59) || left < 13 + value * 8) goto bad;
            size = 13 + value * 8; h = fnv32a(h, d + 13, value * 8); arg = b_from_values(d + 13, value);
        } else if (typ <= 5) {
            if (left < 17 || left - 17 < value) goto bad;
            size = 17 + value; h = fnv32a(h, d + 13, 4 + value);
            uint64_t ignored; arg = b_read_containers(d + 17, value, &ignored); /* ImportRoaringBits iterates the containers only */
            if (!arg) goto bad;
        } else goto bad;
        if (get32(d + 9) != h) { fbo_b_free(arg); goto bad; }
        fbo_bitmap *r = (typ & 1) ? fbo_b_difference(o, arg) : fbo_b_union(o, arg);
        fbo_b_free(arg); fbo_b_free(o); o = r;
        pos += size;
    }
    return o;
bad:
    fbo_b_free(o);
    return NULL;
}

/* ------------------------------------------------------------------ fragment level */

/* fragment.row / rowFromStorage fragment.go:283-333: OffsetRange(shard*2^20, row*2^20, (row+1)*2^20) */
fbo_bitmap *fbo_frag_row(const fbo_bitmap *frag, uint64_t row, uint64_t shard) {
    if (!frag) return fbo_b_new();
    return fbo_b_offset_range(frag, shard << FBO_SHARD_WIDTH_EXP, row << FBO_SHARD_WIDTH_EXP, (row + 1) << FBO_SHARD_WIDTH_EXP);
}

static uint64_t abs_i64(int64_t v) { /* absInt64 fragment.go:952-961 */
    if (v > 0) return (uint64_t)v;
    if (v == INT64_MIN) return 9223372036854775808ull;
    return (uint64_t)(-v);
}
static uint64_t all_ones(uint64_t depth) { return depth >= 64 ? ~0ull : (1ull << depth) - 1; } /* Go: 1<<64 == 0 */
static uint64_t shl_ones(uint64_t depth) { return depth >= 64 ? 0 : ~0ull << depth; }

fbo_bitmap *fbo_frag_row_view(const fbo_bitmap *frag, uint64_t row, uint64_t shard) {
    if (!frag) return fbo_b_new();
    return b_offset_range_view(frag, shard << FBO_SHARD_WIDTH_EXP, row << FBO_SHARD_WIDTH_EXP, (row + 1) << FBO_SHARD_WIDTH_EXP);
}

/* rows inside rangeOp / groupBy are only ever read (every op builds a fresh result), so they are fetched as views; a view
 * that would leave one of these functions as its RESULT is cloned first (own_result) */
typedef struct { const fbo_bitmap *frag; uint64_t shard; } fctx;
static fbo_bitmap *frow(const fctx *f, uint64_t r) { return fbo_frag_row_view(f->frag, r, f->shard); }
static fbo_bitmap *own_result(fbo_bitmap *b) { if (b && b->view) { fbo_bitmap *o = fbo_b_clone(b); fbo_b_free(b); return o; } return b; }
#define OWN2(expr, a, b) ({ fbo_bitmap *_r = (expr); fbo_b_free(a); fbo_b_free(b); _r; })

static fbo_bitmap *range_eq(const fctx *f, uint64_t depth, int64_t pred) { /* rangeEQ fragment.go:963-1003 */
    fbo_bitmap *b = frow(f, 0);
    uint64_t up = abs_i64(pred);
    if ((uint64_t)bitlen64(up) > depth) { fbo_b_free(b); return fbo_b_new(); }
    fbo_bitmap *r = frow(f, 1);
    b = OWN2(pred < 0 ? fbo_b_intersect(b, r) : fbo_b_difference(b, r), b, r);
    for (int i = (int)depth - 1; i >= 0; i--) {
        fbo_bitmap *row = frow(f, 2 + (uint64_t)i);
        b = OWN2(((up >> i) & 1) ? fbo_b_intersect(b, row) : fbo_b_difference(b, row), b, row);
    }
    return b;
}
static fbo_bitmap *range_neq(const fctx *f, uint64_t depth, int64_t pred) { /* rangeNEQ :1005-1022 */
    fbo_bitmap *b = frow(f, 0), *eq = range_eq(f, depth, pred);
    return OWN2(fbo_b_difference(b, eq), b, eq);
}
static fbo_bitmap *range_lt_unsigned(const fctx *f, fbo_bitmap *filter /*owned*/, uint64_t depth, uint64_t pred, int eq) { /* :1070-1113 */
    if ((uint64_t)bitlen64(pred) > depth || (pred == all_ones(depth) && eq)) return filter;
    if (pred == all_ones(depth) && !eq) {
        fbo_bitmap *m = fbo_b_new();
        for (uint64_t i = 0; i < depth; i++) {
            fbo_bitmap *row = frow(f, 2 + i), *d = fbo_b_difference(filter, row);
            m = OWN2(fbo_b_union(m, d), m, d); fbo_b_free(row);
        }
        fbo_b_free(filter); return m;
    }
    if (eq) pred++;
    fbo_bitmap *matched = fbo_b_new(), *remaining = filter;
    for (int i = (int)depth - 1; i >= 0 && pred > 0 && fbo_b_any(remaining); i--) {
        fbo_bitmap *row = frow(f, 2 + (uint64_t)i), *zeroes = fbo_b_difference(remaining, row);
        fbo_b_free(row);
        if ((pred >> i) & 1) { fbo_bitmap *m = fbo_b_union(matched, zeroes); fbo_b_free(matched); fbo_b_free(zeroes); matched = m; pred &= ~(1ull << i); }
        else { fbo_b_free(remaining); remaining = zeroes; }
    }
    fbo_b_free(remaining);
    return matched;
}
static fbo_bitmap *range_gt_unsigned(const fctx *f, fbo_bitmap *filter /*owned*/, uint64_t depth, uint64_t pred, int eq) { /* :1157-1205 */
    for (;;) {
        if (pred == 0 && eq) return filter;
        if (pred == 0 && !eq) {
            fbo_bitmap *m = fbo_b_new();
            for (uint64_t i = 0; i < depth; i++) {
                fbo_bitmap *row = frow(f, 2 + i), *d = fbo_b_intersect(filter, row);
                m = OWN2(fbo_b_union(m, d), m, d); fbo_b_free(row);
            }
            fbo_b_free(filter); return m;
        }
        if (!eq && (uint64_t)bitlen64(pred) > depth) { fbo_b_free(filter); return fbo_b_new(); }
        if (eq) { pred--; eq = 0; continue; }
        break;
    }
    fbo_bitmap *matched = fbo_b_new(), *remaining = filter;
    pred |= shl_ones(depth);
    for (int i = (int)depth - 1; i >= 0 && pred < ~0ull && fbo_b_any(remaining); i--) {
        fbo_bitmap *row = frow(f, 2 + (uint64_t)i), *ones = fbo_b_intersect(remaining, row);
        fbo_b_free(row);
        if ((pred >> i) & 1) { fbo_b_free(remaining); remaining = ones; }
        else { fbo_bitmap *m = fbo_b_union(matched, ones); fbo_b_free(matched); fbo_b_free(ones); matched = m; pred |= 1ull << i; }
    }
    fbo_b_free(remaining);
    return matched;
}
static fbo_bitmap *range_lt(const fctx *f, uint64_t depth, int64_t pred, int eq) { /* rangeLT :1024-1067 */
    if (pred == 1 && !eq) { pred = 0; eq = 1; }
    fbo_bitmap *b = frow(f, 0), *sign = frow(f, 1);
    uint64_t up = abs_i64(pred);
    if (pred == 0 && !eq) return OWN2(fbo_b_intersect(b, sign), b, sign);
    if (pred == 0 && eq) {
        fbo_bitmap *zeroes = range_eq(f, depth, 0), *neg = OWN2(fbo_b_intersect(b, sign), b, sign);
        return OWN2(fbo_b_union(neg, zeroes), neg, zeroes);
    }
    if (pred < 0) { fbo_bitmap *flt = OWN2(fbo_b_intersect(b, sign), NULL, NULL); fbo_b_free(b); fbo_b_free(sign); return range_gt_unsigned(f, flt, depth, up, eq); }
    fbo_bitmap *posf = fbo_b_difference(b, sign);
    fbo_bitmap *pos = range_lt_unsigned(f, posf, depth, up, eq);
    fbo_bitmap *neg = OWN2(fbo_b_intersect(b, sign), b, sign);
    return OWN2(fbo_b_union(pos, neg), pos, neg);
}
static fbo_bitmap *range_gt(const fctx *f, uint64_t depth, int64_t pred, int eq) { /* rangeGT :1115-1155 */
    if (pred == -1 && !eq) { pred = 0; eq = 1; }
    fbo_bitmap *b = frow(f, 0);
    uint64_t up = abs_i64(pred);
    fbo_bitmap *sign = frow(f, 1);
    if (pred == 0 && !eq) { fbo_b_free(b); b = range_neq(f, depth, 0); return OWN2(fbo_b_difference(b, sign), b, sign); }
    if (pred == 0 && eq) return OWN2(fbo_b_difference(b, sign), b, sign);
    if (pred >= 0) { fbo_bitmap *flt = fbo_b_difference(b, sign); fbo_b_free(b); fbo_b_free(sign); return range_gt_unsigned(f, flt, depth, up, eq); }
    fbo_bitmap *negf = fbo_b_intersect(b, sign);
    fbo_bitmap *neg = range_lt_unsigned(f, negf, depth, up, eq);
    fbo_bitmap *pos = OWN2(fbo_b_difference(b, sign), b, sign);
    return OWN2(fbo_b_union(pos, neg), pos, neg);
}
static fbo_bitmap *range_between_unsigned(const fctx *f, fbo_bitmap *filter /*owned*/, uint64_t depth, uint64_t pmin, uint64_t pmax) { /* :1262-1303 */
    if (pmax > all_ones(depth)) return range_gt_unsigned(f, filter, depth, pmin, 1);
    if (pmin == 0) return range_lt_unsigned(f, filter, depth, pmax, 1);
    int diff_len = bitlen64(pmax ^ pmin);
    fbo_bitmap *remaining = filter;
    for (int i = (int)depth - 1; i >= diff_len; i--) {
        fbo_bitmap *row = frow(f, 2 + (uint64_t)i);
        remaining = OWN2(((pmin >> i) & 1) ? fbo_b_intersect(remaining, row) : fbo_b_difference(remaining, row), remaining, row);
    }
    uint64_t mask = shl_ones((uint64_t)diff_len);
    pmin &= ~mask; pmax &= ~mask;
    remaining = range_gt_unsigned(f, remaining, (uint64_t)diff_len, pmin, 1);
    remaining = range_lt_unsigned(f, remaining, (uint64_t)diff_len, pmax, 1);
    return remaining;
}
static fbo_bitmap *range_between(const fctx *f, uint64_t depth, int64_t pmin, int64_t pmax) { /* rangeBetween :1213-1259 */
    uint64_t umin = abs_i64(pmin), umax = abs_i64(pmax);
    if (pmin == pmax) return range_eq(f, depth, pmin);
    fbo_bitmap *b = frow(f, 0), *r = frow(f, 1);
    if (pmin >= 0) { fbo_bitmap *flt = OWN2(fbo_b_difference(b, r), b, r); return range_between_unsigned(f, flt, depth, umin, umax); }
    if (pmax < 0) { fbo_bitmap *flt = OWN2(fbo_b_intersect(b, r), b, r); return range_between_unsigned(f, flt, depth, umax, umin); }
    fbo_bitmap *pos = range_lt_unsigned(f, fbo_b_difference(b, r), depth, umax, 1);
    fbo_bitmap *neg = range_lt_unsigned(f, fbo_b_intersect(b, r), depth, umin, 1);
    fbo_b_free(b); fbo_b_free(r);
    return OWN2(fbo_b_union(pos, neg), pos, neg);
}
fbo_bitmap *fbo_frag_range_op(const fbo_bitmap *frag, uint64_t shard, int op, uint64_t depth, int64_t pred, int64_t pmax) { /* rangeOp :937 */
    fctx f = { frag, shard };
    if (!frag) return fbo_b_new();
    switch (op) {
    case FBO_OP_EQ: return own_result(range_eq(&f, depth, pred));
    case FBO_OP_NEQ: return own_result(range_neq(&f, depth, pred));
    case FBO_OP_LT: return own_result(range_lt(&f, depth, pred, 0));
    case FBO_OP_LTE: return own_result(range_lt(&f, depth, pred, 1));
    case FBO_OP_GT: return own_result(range_gt(&f, depth, pred, 0));
    case FBO_OP_GTE: return own_result(range_gt(&f, depth, pred, 1));
    case FBO_OP_BETWEEN: return own_result(range_between(&f, depth, pred, pmax));
    }
    return NULL;
}

/* doTopK executor.go:2705-2746: stream containers in key order, row = key/16, sub = key%16 */
int64_t fbo_frag_row_counts(const fbo_bitmap *frag, uint64_t shard, const fbo_bitmap *filter, uint64_t *rows, uint64_t *counts, int64_t cap) {
    int64_t k = 0; uint64_t row = ~0ull, count = 0;
    const fbo_container *fc[16]; memset(fc, 0, sizeof fc);
    if (filter) for (int64_t i = 0; i < filter->n; i++) if (filter->keys[i] / 16 == shard) fc[filter->keys[i] % 16] = filter->cs[i]; /* topKFilter.fill :2757 */
    for (int64_t i = 0; i < frag->n; i++) {
        uint64_t keyrow = frag->keys[i] / 16, sub = frag->keys[i] % 16;
        if (keyrow != row) { if (row != ~0ull && count > 0) { if (k < cap) { rows[k] = row; counts[k] = count; } k++; } row = keyrow; count = 0; }
        if (filter) { if (!fc[sub]) continue; count += (uint64_t)fbo_intersection_count(frag->cs[i], fc[sub]); }
        else count += (uint64_t)fbo_c_n(frag->cs[i]);
    }
    if (row != ~0ull && count > 0) { if (k < cap) { rows[k] = row; counts[k] = count; } k++; }
    return k;
}
/* fragment.rows fragment.go:2465-2486 (BitmapRowsFilter roaring/filter.go:252-292): distinct key/16 with N>0 */
int64_t fbo_frag_rows(const fbo_bitmap *frag, uint64_t *rows, int64_t cap) {
    int64_t k = 0; uint64_t last = ~0ull;
    for (int64_t i = 0; i < frag->n; i++) { if (fbo_c_n(frag->cs[i]) == 0) continue; uint64_t r = frag->keys[i] / 16; if (r != last) { if (k < cap) rows[k] = r; k++; last = r; } }
    return k;
}

/* groupByIterator executor.go:8617-8934 + executeGroupByShard :3918-3985, restated as the nested loop it
 * walks: level i row = row_i ∩ (level i-1 row) for i < last (filter folded into level 0 :8829-8835);
 * leaf count = intersectionCount(row_last, level last-1) :8893; only Count>0 emitted (:3960). */
static void gb_rec(const fbo_bitmap *const *frags, int nf, uint64_t shard, const uint64_t *const *ids, const int32_t *n_rows,
                   int level, const fbo_bitmap *prefix, uint64_t *out, uint64_t base) {
    for (int r = 0; r < n_rows[level]; r++) {
        fbo_bitmap *row = fbo_frag_row_view(frags[level], ids[level][r], shard);
        uint64_t idx = base * (uint64_t)n_rows[level] + (uint64_t)r;
        if (level == nf - 1) {
            out[idx] += prefix ? fbo_b_intersection_count(row, prefix) : fbo_b_count(row);
        } else {
            fbo_bitmap *cur = prefix ? fbo_b_intersect(row, prefix) : row;   /* level 0 without a filter walks the row itself :8829 */
            if (fbo_b_any(cur)) gb_rec(frags, nf, shard, ids, n_rows, level + 1, cur, out, idx); /* nextAtIdx skips empty rows :8869 */
            if (cur != row) fbo_b_free(cur);
        }
        fbo_b_free(row);
    }
}
int fbo_groupby_shard(const fbo_bitmap *const *frags, int nf, uint64_t shard, const uint64_t *row_ids_flat, const int32_t *n_rows,
                      const fbo_bitmap *filter, uint64_t *out) {
    if (nf < 1 || nf > 8) return -1;
    const uint64_t *ids[8]; const uint64_t *p = row_ids_flat;
    for (int i = 0; i < nf; i++) { if (!frags[i]) return 0; ids[i] = p; p += n_rows[i]; } /* missing fragment: shard contributes nothing :8769-8772 */
    gb_rec(frags, nf, shard, ids, n_rows, 0, filter, out, 0);
    return 0;
}

