/*
 * bsgs_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see bsgs_oracle.h).
 *
 * Restates, in plain C with unsigned __int128, the algorithm of the reference:
 *   lib/Curve64.pb          field / curve / batch-add arithmetic      (cited per function)
 *   1_9_7File.pb            table builders, on-disk formats, kernel launch, checker
 * No code is shared with the CUDA product.  Parity status: pinned (see header).
 */
#include "bsgs_oracle.h"
#include <string.h>
#include <time.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------ */
/* constants                                                                                    */
/* ------------------------------------------------------------------------------------------ */
static const orc_fe P_ = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const orc_fe N_ = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
static const orc_pt G_ = {
    {{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
    {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}}};
#define FOLD_C 0x1000003D1ULL   /* 2^256 mod p, the constant folded twice in Curve64.pb:1330-1434 */

const orc_fe *orc_curve_p(void) { return &P_; }
const orc_fe *orc_curve_n(void) { return &N_; }
const orc_pt *orc_curve_g(void) { return &G_; }

/* ------------------------------------------------------------------------------------------ */
/* raw 256-bit helpers (m_addX64 / m_subX64 / m_check_less_more_equilX64 / m_shrX64)            */
/* ------------------------------------------------------------------------------------------ */
int orc_raw_add(orc_fe *r, const orc_fe *a, const orc_fe *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (uint64_t)c; c >>= 64; }
    return (int)c;
}
int orc_raw_sub(orc_fe *r, const orc_fe *a, const orc_fe *b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
        uint64_t ai = a->v[i], bi = b->v[i];
        uint64_t d = ai - bi - br;
        br = (ai < bi) || (ai == bi && br);
        r->v[i] = d;
    }
    return (int)br;
}
static int fe_cmp(const orc_fe *a, const orc_fe *b) { /* -1,0,1 */
    for (int i = 3; i >= 0; i--) {
        if (a->v[i] < b->v[i]) return -1;
        if (a->v[i] > b->v[i]) return 1;
    }
    return 0;
}
static int fe_is_zero(const orc_fe *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int fe_eq(const orc_fe *a, const orc_fe *b) { return fe_cmp(a, b) == 0; }
static void fe_shr1(orc_fe *a) {
    a->v[0] = (a->v[0] >> 1) | (a->v[1] << 63);
    a->v[1] = (a->v[1] >> 1) | (a->v[2] << 63);
    a->v[2] = (a->v[2] >> 1) | (a->v[3] << 63);
    a->v[3] >>= 1;
}
static void fe_set_u64(orc_fe *a, uint64_t x) { a->v[0] = x; a->v[1] = a->v[2] = a->v[3] = 0; }

/* ------------------------------------------------------------------------------------------ */
/* generic modular add/sub/inverse (m_addModX64 C64:945-1036, m_subModX64 C64:893-943,          */
/* m_Ecc_modInvX64 C64:2470-2522)                                                               */
/* ------------------------------------------------------------------------------------------ */
void orc_mod_add(orc_fe *r, const orc_fe *a, const orc_fe *b, const orc_fe *m) {
    orc_fe t; int c = orc_raw_add(&t, a, b);
    if (c || fe_cmp(&t, m) >= 0) orc_raw_sub(&t, &t, m);
    *r = t;
}
void orc_mod_sub(orc_fe *r, const orc_fe *a, const orc_fe *b, const orc_fe *m) {
    orc_fe t; int br = orc_raw_sub(&t, a, b);
    if (br) orc_raw_add(&t, &t, m);
    *r = t;
}
/* m_Ecc_modInv_updateX64, C64:2456-2466 */
static void modinv_update(orc_fe *uv, const orc_fe *m) {
    int carry = 0;
    if (uv->v[0] & 1) carry = orc_raw_add(uv, uv, m);
    fe_shr1(uv);
    if (carry) uv->v[3] |= 0x8000000000000000ULL;
}
void orc_mod_inv(orc_fe *r, const orc_fe *inp, const orc_fe *m) {
    if (fe_is_zero(inp)) { memset(r, 0, sizeof *r); return; }       /* C64:2475-2477 */
    orc_fe a = *inp, b = *m, u, v;
    fe_set_u64(&u, 1); fe_set_u64(&v, 0);
    int cmp;
    while ((cmp = fe_cmp(&a, &b)) != 0) {
        if (!(a.v[0] & 1)) { fe_shr1(&a); modinv_update(&u, m); }
        else if (!(b.v[0] & 1)) { fe_shr1(&b); modinv_update(&v, m); }
        else if (cmp > 0) {
            orc_raw_sub(&a, &a, &b); fe_shr1(&a);
            if (fe_cmp(&u, &v) < 0) orc_raw_add(&u, &u, m);
            orc_raw_sub(&u, &u, &v);
            modinv_update(&u, m);
        } else {
            orc_raw_sub(&b, &b, &a); fe_shr1(&b);
            if (fe_cmp(&v, &u) < 0) orc_raw_add(&v, &v, m);
            orc_raw_sub(&v, &v, &u);
            modinv_update(&v, m);
        }
    }
    *r = u;
}

/* ------------------------------------------------------------------------------------------ */
/* field arithmetic mod p                                                                       */
/* ------------------------------------------------------------------------------------------ */
void orc_fe_add(orc_fe *r, const orc_fe *a, const orc_fe *b) { orc_mod_add(r, a, b, &P_); }
void orc_fe_sub(orc_fe *r, const orc_fe *a, const orc_fe *b) { orc_mod_sub(r, a, b, &P_); }
void orc_fe_neg(orc_fe *r, const orc_fe *a) {                      /* m_NegModX64 */
    if (fe_is_zero(a)) { *r = *a; return; }
    orc_raw_sub(r, &P_, a);
}
void orc_fe_inv(orc_fe *r, const orc_fe *a) { orc_mod_inv(r, a, &P_); }

/* 512 -> 256 reduction: fold the high half by 0x1000003D1 twice, then one conditional
 * subtraction (C64:1330-1434); output canonical. */
static void fe_reduce512(orc_fe *r, const uint64_t t[8]) {
    uint64_t s[4]; u128 acc = 0;
    for (int i = 0; i < 4; i++) {
        acc += (u128)t[i] + (u128)t[i + 4] * FOLD_C;
        s[i] = (uint64_t)acc; acc >>= 64;
    }
    uint64_t top = (uint64_t)acc;                 /* < 2^34 */
    acc = (u128)s[0] + (u128)top * FOLD_C; s[0] = (uint64_t)acc; acc >>= 64;
    for (int i = 1; i < 4; i++) { acc += s[i]; s[i] = (uint64_t)acc; acc >>= 64; }
    if (acc) {                                    /* wrapped 2^256 once more: add 2^256 mod p */
        acc = (u128)s[0] + FOLD_C; s[0] = (uint64_t)acc; acc >>= 64;
        for (int i = 1; i < 4 && acc; i++) { acc += s[i]; s[i] = (uint64_t)acc; acc >>= 64; }
    }
    orc_fe o = {{s[0], s[1], s[2], s[3]}};
    if (fe_cmp(&o, &P_) >= 0) orc_raw_sub(&o, &o, &P_);
    *r = o;
}
void orc_fe_mul(orc_fe *r, const orc_fe *a, const orc_fe *b) {      /* m_mulModX64 C64:1038-1437 */
    uint64_t t[8] = {0};
    for (int i = 0; i < 4; i++) {
        u128 carry = 0;
        for (int j = 0; j < 4; j++) {
            carry += (u128)a->v[i] * b->v[j] + t[i + j];
            t[i + j] = (uint64_t)carry; carry >>= 64;
        }
        t[i + 4] = (uint64_t)carry;
    }
    fe_reduce512(r, t);
}
void orc_fe_sqr(orc_fe *r, const orc_fe *a) { orc_fe_mul(r, a, a); } /* m_squareModX64 C64:2161-2455 */

/* ------------------------------------------------------------------------------------------ */
/* curve arithmetic                                                                             */
/* ------------------------------------------------------------------------------------------ */
void orc_pt_dbl(orc_pt *r, const orc_pt *a) {                       /* m_DBLTX64 C64:2524-2551 */
    orc_fe s, dx, tx, ds;
    orc_fe_add(&s, &a->y, &a->y);
    orc_fe_inv(&s, &s);
    orc_fe_sqr(&dx, &a->x);
    orc_fe_add(&tx, &dx, &dx);
    orc_fe_add(&tx, &dx, &tx);
    orc_fe_mul(&s, &tx, &s);
    orc_fe_sqr(&ds, &s);
    orc_fe_sub(&ds, &ds, &a->x);
    orc_fe_sub(&ds, &ds, &a->x);
    orc_fe_sub(&dx, &a->x, &ds);
    orc_fe_mul(&tx, &s, &dx);
    orc_fe y; orc_fe_sub(&y, &tx, &a->y);
    r->x = ds; r->y = y;
}
void orc_pt_add(orc_pt *r, const orc_pt *a, const orc_pt *b) {      /* m_ADDPTX64 C64:2553-2582 */
    if (fe_eq(&a->x, &b->x)) { orc_pt_dbl(r, a); return; }          /* C64:2560 */
    orc_fe s, cx, cy;
    orc_fe_sub(&s, &a->x, &b->x);
    orc_fe_inv(&s, &s);
    orc_fe_sub(&cy, &a->y, &b->y);
    orc_fe_mul(&s, &cy, &s);
    orc_fe_sqr(&cy, &s);
    orc_fe_sub(&cy, &cy, &a->x);
    orc_fe_sub(&cx, &cy, &b->x);
    orc_fe_sub(&cy, &a->x, &cx);
    orc_fe_mul(&cy, &s, &cy);
    orc_fe_sub(&cy, &cy, &a->y);
    r->x = cx; r->y = cy;
}
void orc_pt_mul(orc_pt *r, const orc_pt *a, const orc_fe *k) {      /* m_PTMULX64 C64:2586-2619 */
    orc_fe loc = *k; orc_pt scale = *a, c;
    memset(&c, 0, sizeof c);
    while (!fe_is_zero(&loc)) {
        if (loc.v[0] & 1) {
            if (fe_is_zero(&c.x) || fe_is_zero(&c.y)) c = scale;
            else orc_pt_add(&c, &c, &scale);
        }
        orc_pt_dbl(&scale, &scale);
        fe_shr1(&loc);
    }
    *r = c;
}
void orc_y_from_x(orc_fe *y, const orc_fe *x) {                     /* m_YfromX64 C64:2656-2682 */
    orc_fe s, seven, e, b, res;
    fe_set_u64(&seven, 7);
    orc_fe_mul(&s, x, x); orc_fe_mul(&s, &s, x); orc_fe_add(&s, &s, &seven);
    /* e = (p+1)/4 ; square-and-multiply LSB first (DoPowMod C64:2631-2654) */
    fe_set_u64(&e, 1); orc_raw_add(&e, &P_, &e); fe_shr1(&e); fe_shr1(&e);
    /* note p+1 overflows nothing: p < 2^256-1 */
    fe_set_u64(&res, 1); b = s;
    while (!fe_is_zero(&e)) {
        if (e.v[0] & 1) orc_fe_mul(&res, &res, &b);
        orc_fe_mul(&b, &b, &b);
        fe_shr1(&e);
    }
    *y = res;
}

/* ------------------------------------------------------------------------------------------ */
/* batch add (Curve64.pb:2914-3064)                                                             */
/* ------------------------------------------------------------------------------------------ */
void orc_begin_batch_add(orc_fe *inv_out, size_t n, const orc_pt *a, orc_rec96 *arr) {
    orc_fe s, t; fe_set_u64(&s, 1);
    for (size_t i = 0; i < n; i++) {
        if (fe_eq(&a->x, &arr[i].x)) orc_fe_add(&t, &a->y, &a->y);
        else orc_fe_sub(&t, &a->x, &arr[i].x);
        orc_fe_mul(&s, &s, &t);
        arr[i].chain = s;
    }
    orc_fe_inv(inv_out, &s);
}
void orc_complete_batch_add_with_double(orc_pt *out, size_t n, const orc_pt *a,
                                        const orc_rec96 *arr, const orc_fe *inv_total) {
    if (n == 0) return;
    orc_fe cur = *inv_total, s, t;
    for (size_t i = n; i-- > 0;) {
        if (i > 0) orc_fe_mul(&s, &cur, &arr[i - 1].chain); else s = cur;
        int eq = fe_eq(&a->x, &arr[i].x);
        if (i > 0) {
            if (eq) orc_fe_add(&t, &a->y, &a->y); else orc_fe_sub(&t, &a->x, &arr[i].x);
            orc_fe_mul(&cur, &cur, &t);
        }
        if (eq) { orc_pt_dbl(&out[i], a); continue; }
        orc_fe ny, nx;
        orc_fe_sub(&ny, &a->y, &arr[i].y);
        orc_fe_mul(&s, &ny, &s);
        orc_fe_sqr(&ny, &s);
        orc_fe_sub(&ny, &ny, &a->x);
        orc_fe_sub(&nx, &ny, &arr[i].x);
        orc_fe_sub(&ny, &a->x, &nx);
        orc_fe_mul(&ny, &ny, &s);
        orc_fe_sub(&ny, &ny, &a->y);
        out[i].x = nx; out[i].y = ny;
    }
}
/* arr[i] = (i+1)*A by doubling the filled prefix (fillarrayN C64:3033-3064) */
void orc_fillarray_n(orc_rec96 *arr, size_t n, const orc_pt *a) {
    if (!n) return;
    arr[0].x = a->x; arr[0].y = a->y;
    size_t i = 1;
    orc_pt *tmp = (orc_pt *)malloc(sizeof(orc_pt) * (n ? n : 1));
    while (i < n) {
        size_t k = i; if (k + i >= n) k = n - i;
        orc_pt last = {arr[i - 1].x, arr[i - 1].y};
        orc_fe inv;
        orc_begin_batch_add(&inv, k, &last, arr);
        orc_complete_batch_add_with_double(tmp, k, &last, arr, &inv);
        for (size_t j = 0; j < k; j++) { arr[i + j].x = tmp[j].x; arr[i + j].y = tmp[j].y; }
        i += k;
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------ */
/* small thread helper                                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { void (*fn)(void *, int, int); void *arg; int id, n; } thr_job;
static void *thr_tramp(void *p) { thr_job *j = (thr_job *)p; j->fn(j->arg, j->id, j->n); return NULL; }
static void run_threads(int n, void (*fn)(void *, int, int), void *arg) {
    if (n < 1) n = 1;
    if (n == 1) { fn(arg, 0, 1); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n);
    thr_job *jb = (thr_job *)malloc(sizeof(thr_job) * n);
    for (int i = 0; i < n; i++) { jb[i].fn = fn; jb[i].arg = arg; jb[i].id = i; jb[i].n = n; pthread_create(&th[i], NULL, thr_tramp, &jb[i]); }
    for (int i = 0; i < n; i++) pthread_join(th[i], NULL);
    free(th); free(jb);
}

/* ------------------------------------------------------------------------------------------ */
/* arithmetic progression of points: out(i) = S + i*A for i in [0,count), in batches of 4096    */
/* with a helper array of (i+1)*A -- the structure of baby()/giant() 1_9_7File.pb:1162,1418     */
/* ------------------------------------------------------------------------------------------ */
#define HELPSIZE 4096u    /* #helpsize 1_9_7File.pb:112 */
typedef void (*emit_fn)(void *ctx, uint64_t i, const orc_pt *pt);
static void walk_progression(const orc_pt *S, const orc_pt *A, uint64_t count, emit_fn emit, void *ctx) {
    if (!count) return;
    uint32_t hs = HELPSIZE;
    orc_rec96 *help = (orc_rec96 *)malloc(sizeof(orc_rec96) * hs);
    orc_pt *outb = (orc_pt *)malloc(sizeof(orc_pt) * hs);
    orc_fillarray_n(help, hs, A);
    orc_fe kk; fe_set_u64(&kk, hs);
    orc_pt step; orc_pt_mul(&step, A, &kk);                 /* hs*A */
    orc_pt cur = *S; uint64_t done = 0;
    while (done < count) {
        uint64_t nb = hs; if (nb >= count - done) nb = count - done;
        emit(ctx, done, &cur);
        orc_fe inv;
        orc_begin_batch_add(&inv, (size_t)(nb - 1), &cur, help);
        orc_complete_batch_add_with_double(outb, (size_t)(nb - 1), &cur, help, &inv);
        for (uint64_t i = 0; i + 1 < nb; i++) emit(ctx, done + 1 + i, &outb[i]);
        done += nb;
        if (done < count) orc_pt_add(&cur, &cur, &step);
    }
    free(help); free(outb);
}

/* ---- baby keys ---- */
typedef struct { uint64_t *keys; uint64_t w; } baby_ctx;
typedef struct { uint64_t *keys; uint64_t base; } baby_emit_ctx;
static void baby_emit(void *c, uint64_t i, const orc_pt *pt) { baby_emit_ctx *e = (baby_emit_ctx *)c; e->keys[e->base + i] = pt->x.v[0]; } /* low 8 bytes, F197:1127,1202 */
static void baby_worker(void *arg, int id, int n) {
    baby_ctx *b = (baby_ctx *)arg;
    uint64_t per = (b->w + n - 1) / n, lo = per * id, hi = lo + per; if (hi > b->w) hi = b->w;
    if (lo >= hi) return;
    orc_fe k; fe_set_u64(&k, lo + 1);
    orc_pt S; orc_pt_mul(&S, &G_, &k);
    baby_emit_ctx e = {b->keys, lo};
    walk_progression(&S, &G_, hi - lo, baby_emit, &e);
}
void orc_baby_keys(uint64_t *keys, uint64_t w, int nthreads) {
    baby_ctx b = {keys, w};
    run_threads(nthreads, baby_worker, &b);
}

/* ---- packed hash tables (packHTFile F197:3232-3335, packHTGPUFile F197:3337-3444) ---- */
size_t orc_htcpu_bytes(uint64_t w, uint32_t htsz) { return ((size_t)1 << htsz) * 4 + 4 + (size_t)w * 8; }
size_t orc_htgpu_bytes(uint64_t w, uint32_t htsz) { return ((size_t)1 << htsz) * 4 + 4 + (size_t)w * 4; }
typedef struct { uint32_t tag, j; } ht_item;
static int item_cmp(const void *a, const void *b) {
    const ht_item *x = (const ht_item *)a, *y = (const ht_item *)b;
    if (x->tag != y->tag) return x->tag < y->tag ? -1 : 1;      /* unsigned, check_LME32bit F197:2666 */
    if (x->j != y->j) return x->j > y->j ? -1 : 1;              /* equal tags: later insert first, F197:2797-2805 */
    return 0;
}
void orc_build_ht(const uint64_t *keys, uint64_t w, uint32_t htsz, uint8_t *htcpu, uint8_t *htgpu) {
    size_t H = (size_t)1 << htsz; uint32_t mask = (uint32_t)(H - 1);
    uint32_t *off = (uint32_t *)calloc(H + 1, 4);
    for (uint64_t j = 0; j < w; j++) off[((uint32_t)keys[j] & mask) + 1]++;      /* bucket = low32 & mask, F197:2561 */
    for (size_t h = 0; h < H; h++) off[h + 1] += off[h];
    ht_item *items = (ht_item *)malloc(sizeof(ht_item) * (w ? w : 1));
    uint32_t *cur = (uint32_t *)malloc(4 * H);
    memcpy(cur, off, 4 * H);
    for (uint64_t j = 0; j < w; j++) {
        uint32_t h = (uint32_t)keys[j] & mask;
        items[cur[h]].tag = (uint32_t)(keys[j] >> 32);                              /* tag = bytes 4..7, F197:2583 */
        items[cur[h]].j = (uint32_t)j;                                              /* payload, F197:2584 */
        cur[h]++;
    }
    for (size_t h = 0; h < H; h++) if (off[h + 1] - off[h] > 1) qsort(items + off[h], off[h + 1] - off[h], sizeof(ht_item), item_cmp);
    if (htcpu) { memcpy(htcpu, off, 4 * (H + 1)); memcpy(htcpu + 4 * (H + 1), items, (size_t)w * 8); }
    if (htgpu) {
        memcpy(htgpu, off, 4 * (H + 1));
        uint32_t *t = (uint32_t *)(htgpu + 4 * (H + 1));
        for (uint64_t i = 0; i < w; i++) t[i] = items[i].tag;
    }
    free(off); free(items); free(cur);
}
int orc_probe_htgpu(const uint8_t *htgpu, uint32_t htsz, uint64_t key64) {
    size_t H = (size_t)1 << htsz; const uint32_t *off = (const uint32_t *)htgpu;
    const uint32_t *tags = off + H + 1;
    uint32_t h = (uint32_t)key64 & (uint32_t)(H - 1), tag = (uint32_t)(key64 >> 32);
    uint32_t lo = 0, hi = off[h + 1] - off[h]; const uint32_t *t = tags + off[h];
    while (lo < hi) {                                   /* PTX197:L75-101 */
        uint32_t c = (lo + hi) >> 1;
        if (tag > t[c]) { lo = c + 1; if (lo > hi) lo = hi; }
        else if (tag < t[c]) hi = c;
        else return 1;
    }
    return 0;
}
uint32_t orc_lookup_htcpu(const uint8_t *htcpu, uint32_t htsz, uint64_t key64, uint32_t *js, uint32_t cap) {
    size_t H = (size_t)1 << htsz; const uint32_t *off = (const uint32_t *)htcpu;
    const ht_item *it = (const ht_item *)(htcpu + 4 * (H + 1));
    uint32_t h = (uint32_t)key64 & (uint32_t)(H - 1), tag = (uint32_t)(key64 >> 32), n = 0;
    for (uint32_t i = off[h]; i < off[h + 1]; i++) if (it[i].tag == tag && n < cap) js[n++] = it[i].j;
    return n;
}

/* ---- giants / G2 (giant F197:1418-1488, Writeint F197:1831-1903, toLittleInd32 F197:254) ---- */
static inline size_t g2_word_index(uint32_t T, uint32_t B, uint32_t p, uint64_t k, int word) {
    uint64_t TB = (uint64_t)T * B;
    uint64_t blk = k / ((uint64_t)T * p), thr = (k % ((uint64_t)T * p)) / p, r = k % p;
    uint64_t g = blk * T + thr;
    return (size_t)(r * 8 * TB + g + (uint64_t)word * TB);
}
static void g2_put(uint8_t *g2, uint32_t T, uint32_t B, uint32_t p, uint64_t k, const orc_pt *q) {
    uint64_t N = (uint64_t)T * B * p; uint32_t *X = (uint32_t *)g2, *Y = (uint32_t *)(g2 + N * 32);
    for (int wd = 0; wd < 8; wd++) {                      /* word 0 = most significant 32 bits */
        int limb = 7 - wd;
        X[g2_word_index(T, B, p, k, wd)] = (uint32_t)(q->x.v[limb >> 1] >> ((limb & 1) * 32));
        Y[g2_word_index(T, B, p, k, wd)] = (uint32_t)(q->y.v[limb >> 1] >> ((limb & 1) * 32));
    }
}
void orc_g2_get(orc_pt *q, const uint8_t *g2, uint32_t T, uint32_t B, uint32_t p, uint64_t k) {
    uint64_t N = (uint64_t)T * B * p; const uint32_t *X = (const uint32_t *)g2, *Y = (const uint32_t *)(g2 + N * 32);
    memset(q, 0, sizeof *q);
    for (int wd = 0; wd < 8; wd++) {
        int limb = 7 - wd;
        q->x.v[limb >> 1] |= (uint64_t)X[g2_word_index(T, B, p, k, wd)] << ((limb & 1) * 32);
        q->y.v[limb >> 1] |= (uint64_t)Y[g2_word_index(T, B, p, k, wd)] << ((limb & 1) * 32);
    }
}
typedef struct { uint8_t *out; uint32_t T, B, p; uint64_t w, N; orc_pt A; } g2_ctx;
typedef struct { g2_ctx *g; uint64_t base; } g2_emit_ctx;
static void g2_emit(void *c, uint64_t i, const orc_pt *pt) { g2_emit_ctx *e = (g2_emit_ctx *)c; g2_put(e->g->out, e->g->T, e->g->B, e->g->p, e->base + i, pt); }
static void g2_worker(void *arg, int id, int n) {
    g2_ctx *g = (g2_ctx *)arg;
    uint64_t per = (g->N + n - 1) / n, lo = per * id, hi = lo + per; if (hi > g->N) hi = g->N;
    if (lo >= hi) return;
    orc_fe k; fe_set_u64(&k, lo + 1);
    orc_pt S; orc_pt_mul(&S, &g->A, &k);
    g2_emit_ctx e = {g, lo};
    walk_progression(&S, &g->A, hi - lo, g2_emit, &e);
}
static void neg_2w_G(orc_pt *A, uint64_t w) {            /* ADDPUBG = -(2w)G, F197:4689-4698 */
    orc_fe k; memset(&k, 0, sizeof k);
    u128 t = (u128)w * 2; k.v[0] = (uint64_t)t; k.v[1] = (uint64_t)(t >> 64);
    orc_pt_mul(A, &G_, &k);
    orc_fe_neg(&A->y, &A->y);
}
void orc_build_g2(uint8_t *out, uint32_t T, uint32_t B, uint32_t p, uint64_t w, int nthreads) {
    g2_ctx g; g.out = out; g.T = T; g.B = B; g.p = p; g.w = w; g.N = (uint64_t)T * B * p;
    neg_2w_G(&g.A, w);
    run_threads(nthreads, g2_worker, &g);
}

/* ------------------------------------------------------------------------------------------ */
/* one kernel launch (PTX197:L49-109 code 5, L149-1297 PASS 1, L1319-26062 invert,  */
/* L26955-36020 PASS 2).  Records are a multiset; order is not part of the contract.            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *g2, *ht; uint32_t T, B, p, htsz, batch; const orc_pt *P;
    orc_hit *out; uint32_t cap; uint32_t count; uint64_t k0, k1; pthread_mutex_t mu;
} launch_ctx;
static void push_hit(launch_ctx *L, uint32_t code, uint32_t idx) {
    pthread_mutex_lock(&L->mu);
    if (L->count < L->cap) { L->out[L->count].code = code; L->out[L->count].idx = idx; }
    L->count++;
    pthread_mutex_unlock(&L->mu);
}
static void launch_worker(void *arg, int id, int n) {
    launch_ctx *L = (launch_ctx *)arg; const orc_pt *P = L->P;
    uint32_t bs = L->batch;
    orc_pt *Q = (orc_pt *)malloc(sizeof(orc_pt) * bs);
    orc_fe *chain = (orc_fe *)malloc(sizeof(orc_fe) * bs);
    orc_fe twoPy; orc_fe_add(&twoPy, &P->y, &P->y);
    uint64_t nb = (L->k1 - L->k0 + bs - 1) / bs;
    for (uint64_t b = id; b < nb; b += n) {
        uint64_t kb = L->k0 + b * bs, ke = kb + bs; if (ke > L->k1) ke = L->k1;
        size_t m = (size_t)(ke - kb);
        orc_fe acc, d; fe_set_u64(&acc, 1);
        for (size_t i = 0; i < m; i++) {                                   /* PASS 1 */
            orc_g2_get(&Q[i], L->g2, L->T, L->B, L->p, kb + i);
            if (fe_eq(&P->x, &Q[i].x)) d = twoPy; else orc_fe_sub(&d, &P->x, &Q[i].x);
            orc_fe_mul(&acc, &acc, &d);
            chain[i] = acc;
        }
        orc_fe_inv(&acc, &acc);                                            /* kernel: a^(p-2); same value */
        for (size_t i = m; i-- > 0;) {                                     /* PASS 2 */
            orc_fe inv_i, lam, X;
            if (i > 0) orc_fe_mul(&inv_i, &acc, &chain[i - 1]); else inv_i = acc;
            int eq = fe_eq(&P->x, &Q[i].x);
            if (eq) d = twoPy; else orc_fe_sub(&d, &P->x, &Q[i].x);
            orc_fe_mul(&acc, &acc, &d);
            if (eq) {                                                      /* true doubling -> code 4 */
                orc_fe t3, xx;
                orc_fe_sqr(&xx, &P->x); orc_fe_add(&t3, &xx, &xx); orc_fe_add(&t3, &t3, &xx);
                orc_fe_mul(&lam, &t3, &inv_i);
                orc_fe_sqr(&X, &lam); orc_fe_sub(&X, &X, &P->x); orc_fe_sub(&X, &X, &P->x);
                if (orc_probe_htgpu(L->ht, L->htsz, X.v[0])) push_hit(L, 4, (uint32_t)(kb + i));
                continue;
            }
            orc_fe s;
            orc_fe_add(&s, &P->y, &Q[i].y);                                 /* minus side: P - Q, code 2 */
            orc_fe_mul(&lam, &s, &inv_i);
            orc_fe_sqr(&X, &lam); orc_fe_sub(&X, &X, &P->x); orc_fe_sub(&X, &X, &Q[i].x);
            if (orc_probe_htgpu(L->ht, L->htsz, X.v[0])) push_hit(L, 2, (uint32_t)(kb + i));
            orc_fe_sub(&s, &P->y, &Q[i].y);                                 /* plus side: P + Q, code 1 */
            orc_fe_mul(&lam, &s, &inv_i);
            orc_fe_sqr(&X, &lam); orc_fe_sub(&X, &X, &P->x); orc_fe_sub(&X, &X, &Q[i].x);
            if (orc_probe_htgpu(L->ht, L->htsz, X.v[0])) push_hit(L, 1, (uint32_t)(kb + i));
        }
    }
    free(Q); free(chain);
}
uint32_t orc_launch(const uint8_t *g2, uint32_t T, uint32_t B, uint32_t p,
                    const uint8_t *htgpu, uint32_t htsz,
                    const orc_pt *P, orc_hit *out, uint32_t cap,
                    uint32_t batch, int nthreads, uint64_t k_begin, uint64_t k_end) {
    launch_ctx L; memset(&L, 0, sizeof L);
    uint64_t N = (uint64_t)T * B * p;
    if (k_end > N) k_end = N;
    L.g2 = g2; L.ht = htgpu; L.T = T; L.B = B; L.p = p; L.htsz = htsz; L.batch = batch ? batch : p;
    L.P = P; L.out = out; L.cap = cap; L.k0 = k_begin; L.k1 = k_end;
    pthread_mutex_init(&L.mu, NULL);
    if (k_begin == 0 && orc_probe_htgpu(htgpu, htsz, P->x.v[0])) push_hit(&L, 5, 0);   /* thread 0, PTX197:L49-109 */
    if (k_end > k_begin) run_threads(nthreads, launch_worker, &L);
    pthread_mutex_destroy(&L.mu);
    return L.count;
}

/* ------------------------------------------------------------------------------------------ */
/* window enumeration and hit decoding                                                          */
/* ------------------------------------------------------------------------------------------ */
static void fe_from_u128(orc_fe *r, u128 x) { r->v[0] = (uint64_t)x; r->v[1] = (uint64_t)(x >> 64); r->v[2] = r->v[3] = 0; }
static void neg_mul_G(orc_pt *r, const orc_fe *k) { orc_pt_mul(r, &G_, k); if (!fe_is_zero(&r->y)) orc_fe_neg(&r->y, &r->y); }

void orc_search_init(orc_search *s, uint32_t T, uint32_t B, uint32_t p, uint64_t w, uint32_t htsz,
                     const orc_fe *pk, const orc_pt *target, const orc_fe *start_key) {
    memset(s, 0, sizeof *s);
    s->T = T; s->B = B; s->p = p; s->w = w; s->htsz = htsz; s->pk = *pk; s->target = *target;
    u128 N = (u128)T * B * p;
    neg_2w_G(&s->addpubg, w);                                            /* F197:4689-4698 */
    fe_from_u128(&s->centerbig, (u128)p * w);                            /* F197:4708-4712 */
    neg_mul_G(&s->center, &s->centerbig);
    fe_from_u128(&s->prkaddbig, N * 2 * (u128)w * 2);                    /* F197:4759-4765 (N < 2^28, w < 2^32) */
    neg_mul_G(&s->pubaddbig, &s->prkaddbig);
    orc_pt negpk; neg_mul_G(&negpk, pk);                                 /* PubkeyBIG = -(pk)G */
    if (fe_is_zero(pk)) s->findpub = *target; else orc_pt_add(&s->findpub, target, &negpk);   /* F197:5042 */
    if (start_key) s->globkey = *start_key; else fe_set_u64(&s->globkey, 1);
    orc_pt gk; neg_mul_G(&gk, &s->globkey);                              /* F197:5054-5063 */
    orc_pt_add(&s->globpub, &s->findpub, &gk);
    orc_pt_add(&s->globpub, &s->globpub, &s->center);
}
void orc_get_job(orc_search *s, orc_fe *key, orc_pt *P) {                /* GetJob F197:2077-2092 */
    *key = s->globkey; *P = s->globpub;
    orc_pt_add(&s->globpub, &s->globpub, &s->pubaddbig);
    orc_mod_add(&s->globkey, &s->globkey, &s->prkaddbig, &N_);
}
/* K*G == findpub (x compare, then fix sign), then pk+K verified against the real pubkey */
static int try_candidate(const orc_search *s, const orc_fe *K, orc_fe *key_out) {
    orc_pt t; orc_fe k = *K;
    if (fe_is_zero(&k)) return 0;
    orc_pt_mul(&t, &G_, &k);
    if (!fe_eq(&t.x, &s->findpub.x)) return 0;
    if (!fe_eq(&t.y, &s->findpub.y)) orc_mod_sub(&k, &N_, &k, &N_);
    orc_fe fin; orc_mod_add(&fin, &s->pk, &k, &N_);
    orc_pt_mul(&t, &G_, &fin);
    if (!fe_eq(&t.x, &s->target.x)) return 0;
    if (!fe_eq(&t.y, &s->target.y)) orc_mod_sub(&fin, &N_, &fin, &N_);   /* F197:4135-4144 */
    *key_out = fin;
    return 1;
}
int orc_check_hit(const orc_search *s, const uint8_t *htcpu, const orc_hit *h,
                  const orc_fe *wkey, const orc_pt *P, orc_fe *key_out) {
    orc_fe base; orc_mod_add(&base, wkey, &s->centerbig, &N_);          /* key_j + p*w */
    orc_fe two_w_i; fe_from_u128(&two_w_i, (u128)(h->idx + 1) * (u128)s->w * 2);
    uint32_t js[64]; uint32_t nj;
    if (h->code == 4) {                                                  /* F197:3974-4019 */
        orc_fe K; orc_mod_add(&K, &base, &two_w_i, &N_);
        if (try_candidate(s, &K, key_out)) return 1;
        orc_mod_sub(&K, &base, &two_w_i, &N_);
        return try_candidate(s, &K, key_out);
    }
    if (h->code == 5) {                                                  /* F197:4025-4077 */
        nj = orc_lookup_htcpu(htcpu, s->htsz, P->x.v[0], js, 64);
        for (uint32_t i = 0; i < nj; i++) {
            orc_fe b, K; fe_set_u64(&b, (uint64_t)js[i] + 1);
            orc_mod_add(&K, &base, &b, &N_); if (try_candidate(s, &K, key_out)) return 1;
            orc_mod_sub(&K, &base, &b, &N_); if (try_candidate(s, &K, key_out)) return 1;
        }
        return 0;
    }
    if (h->code != 1 && h->code != 2) return 0;
    /* recompute the candidate point X to find b in htCPU (F197:4083-4124) */
    orc_fe ik; fe_set_u64(&ik, (uint64_t)h->idx + 1);
    orc_pt Q, R; orc_pt_mul(&Q, &s->addpubg, &ik);
    if (h->code == 2) orc_fe_neg(&Q.y, &Q.y);
    orc_pt_add(&R, P, &Q);
    nj = orc_lookup_htcpu(htcpu, s->htsz, R.x.v[0], js, 64);
    for (uint32_t i = 0; i < nj; i++) {
        orc_fe b, off, K; fe_set_u64(&b, (uint64_t)js[i] + 1);
        for (int sgn = 0; sgn < 2; sgn++) {
            if (sgn == 0) orc_mod_add(&off, &two_w_i, &b, &N_); else orc_mod_sub(&off, &two_w_i, &b, &N_);
            if (h->code == 1) orc_mod_add(&K, &base, &off, &N_); else orc_mod_sub(&K, &base, &off, &N_);
            if (try_candidate(s, &K, key_out)) return 1;
        }
    }
    return 0;
}

/* The reference's own single-core sanity timing (Curve64.pb:3303-3320): b = G; n times b = G + b with
 * m_ADDPTX64 (one binary-GCD inversion per addition).  Returns the elapsed seconds; *out = final b
 * (n = 10000: x = db7432110ba8...65a8, the known answer of that block). */
double orc_time_seq_adds(uint32_t n, orc_pt *out) {
    struct timespec t0, t1;
    orc_pt a = G_, b = G_;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t i = 0; i < n; i++) orc_pt_add(&b, &a, &b);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (out) *out = b;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

uint64_t orc_parse_w(const char *arg) {                                  /* F197:1009-1022 */
    double d = atof(arg);
    if (d <= 32.0) return (uint64_t)pow(2.0, d);
    return strtoull(arg, NULL, 10);
}
