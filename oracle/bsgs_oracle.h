/*
 * bsgs_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the reference's algorithm for the giant-step hot path of
 * Etayson/BSGS-cuda.  It exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs can CHECK the CUDA product (libbsgs_b200.so) and time a
 * CPU baseline beside it.  Nothing in the product path (bsgs-cuda_b200/, include/) may include,
 * link or call this file.
 *
 * Parity status: PINNED.  Field/curve arithmetic is pinned by every known-answer vector in the
 * reference's own self-test block lib/Curve64.pb:3067-3397 (tests/test_oracle_kat.py); the launch
 * restatement is pinned on the GPU box against the reference's own embedded PTX kernel (decoded
 * from 1_9_7File.pb:5181-23979 at build time into oracle/_ref/, see oracle/Makefile and
 * oracle/ref_harness.c) on identical tables and base points (tests/test_gpu_parity.py).
 *
 * Conventions: a field element / scalar is 4 x u64, little-endian limbs, always canonical
 * (lib/Curve64.pb:450-473 "m_sethex32": 32-byte little-endian big-ints).
 */
#ifndef BSGS_ORACLE_H
#define BSGS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t v[4]; } orc_fe;          /* element of F_p or Z_n, canonical */
typedef struct { orc_fe x, y; } orc_pt;            /* affine point; (0,0) = "empty" as in m_PTMULX64 */
typedef struct { uint32_t code, idx; } orc_hit;    /* one kernel record, 1_9_7File.pb:2475-2490 */

/* ---- field arithmetic mod p = 2^256-2^32-977 (lib/Curve64.pb:893-1437, 2161-2455) ---- */
void orc_fe_add(orc_fe *r, const orc_fe *a, const orc_fe *b);
void orc_fe_sub(orc_fe *r, const orc_fe *a, const orc_fe *b);
void orc_fe_neg(orc_fe *r, const orc_fe *a);
void orc_fe_mul(orc_fe *r, const orc_fe *a, const orc_fe *b);
void orc_fe_sqr(orc_fe *r, const orc_fe *a);
void orc_fe_inv(orc_fe *r, const orc_fe *a);       /* binary xGCD, inv(0)=0: Curve64.pb:2470-2522 */
/* generic add/sub/inv modulo an arbitrary odd modulus (used with n for scalars) */
void orc_mod_add(orc_fe *r, const orc_fe *a, const orc_fe *b, const orc_fe *m);
void orc_mod_sub(orc_fe *r, const orc_fe *a, const orc_fe *b, const orc_fe *m);
void orc_mod_inv(orc_fe *r, const orc_fe *a, const orc_fe *m);
/* raw 256-bit add/sub returning carry/borrow (m_addX64/m_subX64) */
int  orc_raw_add(orc_fe *r, const orc_fe *a, const orc_fe *b);
int  orc_raw_sub(orc_fe *r, const orc_fe *a, const orc_fe *b);
const orc_fe *orc_curve_p(void);
const orc_fe *orc_curve_n(void);
const orc_pt *orc_curve_g(void);

/* ---- curve arithmetic (lib/Curve64.pb:2524-2682) ---- */
void orc_pt_dbl(orc_pt *r, const orc_pt *a);                         /* m_DBLTX64 */
void orc_pt_add(orc_pt *r, const orc_pt *a, const orc_pt *b);        /* m_ADDPTX64: x equal -> double */
void orc_pt_mul(orc_pt *r, const orc_pt *a, const orc_fe *k);        /* m_PTMULX64, LSB first */
void orc_y_from_x(orc_fe *y, const orc_fe *x);                       /* m_YfromX64 */

/* ---- Montgomery-trick batch add over 96-byte (x,y,chain) records (Curve64.pb:2914-3064) ---- */
typedef struct { orc_fe x, y, chain; } orc_rec96;
void orc_begin_batch_add(orc_fe *inv_out, size_t n, const orc_pt *a, orc_rec96 *arr);
void orc_complete_batch_add_with_double(orc_pt *out, size_t n, const orc_pt *a,
                                        const orc_rec96 *arr, const orc_fe *inv_total);
void orc_fillarray_n(orc_rec96 *arr, size_t n, const orc_pt *a);

/* ---- tables (1_9_7File.pb:1162-1235, 2555-2622, 2771-2820, 3232-3444, 1418-1488, 1831-1903) ---- */
/* low 8 bytes of X((j+1)*G) for j = 0..w-1 */
void orc_baby_keys(uint64_t *keys, uint64_t w, int nthreads);
/* sizes in bytes of the packed images */
size_t orc_htcpu_bytes(uint64_t w, uint32_t htsz);
size_t orc_htgpu_bytes(uint64_t w, uint32_t htsz);
/* build both packed images from the keys; either output may be NULL */
void orc_build_ht(const uint64_t *keys, uint64_t w, uint32_t htsz, uint8_t *htcpu, uint8_t *htgpu);
/* G2 file image: X block then Y block, transposed word order; out holds N*64 bytes, N=T*B*p */
void orc_build_g2(uint8_t *out, uint32_t T, uint32_t B, uint32_t p, uint64_t w, int nthreads);
/* fetch giant k out of a G2 image (inverse of Writeint+toLittleInd32) */
void orc_g2_get(orc_pt *q, const uint8_t *g2, uint32_t T, uint32_t B, uint32_t p, uint64_t k);
/* CPU twin of the kernel probe (ReadHTpack + binary search on tags): 1 = present */
int  orc_probe_htgpu(const uint8_t *htgpu, uint32_t htsz, uint64_t key64);
/* htCPU lookup: all payloads j whose (bucket,tag) match; returns count written (<= cap) */
uint32_t orc_lookup_htcpu(const uint8_t *htcpu, uint32_t htsz, uint64_t key64, uint32_t *js, uint32_t cap);

/* ---- one kernel launch (_test1, 1_9_7File.pb:5181-23979) ----
 * Emits the multiset of (code, idx) records for base point P.  batch = points per inversion
 * (the kernel uses batch = p; the CPU baseline uses 4096 = #helpsize).  Returns the total number
 * of records found (may exceed cap; only cap are stored). */
uint32_t orc_launch(const uint8_t *g2, uint32_t T, uint32_t B, uint32_t p,
                    const uint8_t *htgpu, uint32_t htsz,
                    const orc_pt *P, orc_hit *out, uint32_t cap,
                    uint32_t batch, int nthreads, uint64_t k_begin, uint64_t k_end);

/* ---- window enumeration + hit decoding (1_9_7File.pb:2077-2092, 4689-4765, 5042-5064, 3933-4296) ---- */
typedef struct {
    uint32_t T, B, p, htsz; uint64_t w;
    orc_fe pk;            /* range start */
    orc_pt target;        /* REALPUB */
    orc_pt findpub;       /* FINDPUBG = target - pk*G */
    orc_pt addpubg;       /* -(2w)G */
    orc_pt center;        /* -(p*w)G */
    orc_pt pubaddbig;     /* -(4Nw)G */
    orc_fe prkaddbig;     /* 4Nw */
    orc_fe centerbig;     /* p*w */
    orc_fe globkey;       /* key of next window */
    orc_pt globpub;       /* base point of next window */
} orc_search;
void orc_search_init(orc_search *s, uint32_t T, uint32_t B, uint32_t p, uint64_t w, uint32_t htsz,
                     const orc_fe *pk, const orc_pt *target, const orc_fe *start_key /*NULL => 1*/);
void orc_get_job(orc_search *s, orc_fe *key, orc_pt *P);     /* GetJob */
/* try to turn one record into the private key; returns 1 and writes key on success */
int  orc_check_hit(const orc_search *s, const uint8_t *htcpu, const orc_hit *h,
                   const orc_fe *window_key, const orc_pt *P, orc_fe *key_out);

/* -w parser: value <= 32 means 2^x truncated, else decimal (1_9_7File.pb:1009-1022) */
uint64_t orc_parse_w(const char *arg);
double orc_time_seq_adds(uint32_t n, orc_pt *out);   /* Curve64.pb:3303-3320 */

#ifdef __cplusplus
}
#endif
#endif
