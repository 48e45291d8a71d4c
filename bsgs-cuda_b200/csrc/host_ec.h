// host_ec.h -- small host-side secp256k1 arithmetic used by the library's control path only:
// deriving the per-window constants handed to the kernel (p - Px, 2*Py) and stepping the window
// base point P_{j+1} = P_j + D in bsgs_run (the reference does this step on the host as well:
// GetJob, 1_9_7File.pb:2077-2092).  Not on the hot path; the giant-step arithmetic lives in fp.cuh.
#pragma once
#include <stdint.h>
#include <string.h>

namespace hec {

typedef unsigned __int128 u128;
struct Fe { uint64_t v[4]; };
struct Pt { Fe x, y; };

static const Fe kP = {{0xFFFFFFFEFFFFFC2FULL, ~0ULL, ~0ULL, ~0ULL}};
static const uint64_t kC = 0x1000003D1ULL;

static inline bool is_zero(const Fe& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static inline bool eq(const Fe& a, const Fe& b) { return a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2] && a.v[3] == b.v[3]; }
static inline bool geq_p(const Fe& a) {
    return a.v[3] == ~0ULL && a.v[2] == ~0ULL && a.v[1] == ~0ULL && a.v[0] >= kP.v[0];
}
static inline void load_le(Fe& r, const uint8_t b[32]) { memcpy(r.v, b, 32); }
static inline void store_le(uint8_t b[32], const Fe& a) { memcpy(b, a.v, 32); }

static inline void add(Fe& r, const Fe& a, const Fe& b) {
    u128 c = 0; Fe t;
    for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; t.v[i] = (uint64_t)c; c >>= 64; }
    if (c) { c = (u128)t.v[0] + kC; t.v[0] = (uint64_t)c; c >>= 64; for (int i = 1; i < 4; i++) { c += t.v[i]; t.v[i] = (uint64_t)c; c >>= 64; } }
    if (geq_p(t)) { c = (u128)t.v[0] + kC; t.v[0] = (uint64_t)c; c >>= 64; for (int i = 1; i < 4; i++) { c += t.v[i]; t.v[i] = (uint64_t)c; c >>= 64; } }
    r = t;
}
static inline void neg(Fe& r, const Fe& a) {
    if (is_zero(a)) { r = a; return; }
    uint64_t br = 0; Fe t;
    for (int i = 0; i < 4; i++) { u128 d = (u128)kP.v[i] - a.v[i] - br; t.v[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    r = t;
}
static inline void sub(Fe& r, const Fe& a, const Fe& b) { Fe nb; neg(nb, b); add(r, a, nb); }
static inline void mul(Fe& r, const Fe& a, const Fe& b) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.v[i] * b.v[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
        t[i + 4] = (uint64_t)c;
    }
    u128 c = 0; Fe s;
    for (int i = 0; i < 4; i++) { c += (u128)t[i] + (u128)t[i + 4] * kC; s.v[i] = (uint64_t)c; c >>= 64; }
    uint64_t top = (uint64_t)c;
    c = (u128)s.v[0] + (u128)top * kC; s.v[0] = (uint64_t)c; c >>= 64;
    for (int i = 1; i < 4; i++) { c += s.v[i]; s.v[i] = (uint64_t)c; c >>= 64; }
    if (c) { c = (u128)s.v[0] + kC; s.v[0] = (uint64_t)c; c >>= 64; for (int i = 1; i < 4; i++) { c += s.v[i]; s.v[i] = (uint64_t)c; c >>= 64; } }
    if (geq_p(s)) { c = (u128)s.v[0] + kC; s.v[0] = (uint64_t)c; c >>= 64; for (int i = 1; i < 4; i++) { c += s.v[i]; s.v[i] = (uint64_t)c; c >>= 64; } }
    r = s;
}
static inline void sqr_n(Fe& r, int n) { for (int i = 0; i < n; i++) mul(r, r, r); }
// Fermat inverse with the 255S+15M chain; inv(0) = 0
static inline void inv(Fe& r, const Fe& a) {
    Fe x2, x3, x22, x44, t, u;
    t = a; sqr_n(t, 1); mul(x2, t, a);
    t = x2; sqr_n(t, 1); mul(x3, t, a);
    t = x3; sqr_n(t, 3); mul(u, t, x3);
    t = u; sqr_n(t, 3); mul(u, t, x3);
    t = u; sqr_n(t, 2); mul(u, t, x2);
    t = u; sqr_n(t, 11); mul(x22, t, u);
    t = x22; sqr_n(t, 22); mul(x44, t, x22);
    t = x44; sqr_n(t, 44); mul(u, t, x44);
    t = u; sqr_n(t, 88); mul(u, t, u);
    t = u; sqr_n(t, 44); mul(u, t, x44);
    t = u; sqr_n(t, 3); mul(u, t, x3);
    t = u; sqr_n(t, 23); mul(u, t, x22);
    t = u; sqr_n(t, 5); mul(u, t, a);
    t = u; sqr_n(t, 3); mul(u, t, x2);
    t = u; sqr_n(t, 2); mul(r, t, a);
}
// affine add with the reference's convention: equal x => doubling (m_ADDPTX64, Curve64.pb:2553-2582)
static inline void pt_add(Pt& r, const Pt& a, const Pt& b) {
    Fe lam, t, den;
    if (eq(a.x, b.x)) {
        add(den, a.y, a.y); inv(den, den);
        mul(t, a.x, a.x); add(lam, t, t); add(lam, lam, t); mul(lam, lam, den);
    } else {
        sub(den, a.x, b.x); inv(den, den);
        sub(t, a.y, b.y); mul(lam, t, den);
    }
    Fe x3, y3;
    mul(x3, lam, lam); sub(x3, x3, a.x); sub(x3, x3, b.x);
    sub(t, a.x, x3); mul(y3, lam, t); sub(y3, y3, a.y);
    r.x = x3; r.y = y3;
}

static const Pt kG = {{{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
                      {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}}};
static inline void pt_neg(Pt& r, const Pt& a) { r.x = a.x; neg(r.y, a.y); }
// k*a for 1 <= k < 2^64, LSB-first double-and-add like m_PTMULX64 (Curve64.pb:2586-2619)
static inline void pt_mul_u64(Pt& r, const Pt& a, uint64_t k) {
    Pt acc = a, res; bool have = false;
    while (k) {
        if (k & 1) { if (have) pt_add(res, res, acc); else { res = acc; have = true; } }
        k >>= 1;
        if (k) pt_add(acc, acc, acc);
    }
    r = res;
}

}  // namespace hec
