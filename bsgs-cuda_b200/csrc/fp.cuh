// fp.cuh -- secp256k1 field arithmetic for sm_100a, 8 x u32 little-endian limbs.
//
// Design (B200): every 32x32->64 limb product is ONE IMAD.WIDE.U32 with the 64-bit accumulate and
// the carry folded in (ptxas fuses a `mad.lo.cc / madc.hi.cc` pair on the same operands into
// IMAD.WIDE.U32[.X] with predicate carry-in/out).  Products are accumulated into two 512-bit
// accumulators, E for columns of even weight and O for columns of odd weight, so that every row of
// the schoolbook is an unbroken carry chain of 64-bit slots (the layout used by the fastest public
// GPU big-int code).  The reference kernel instead issues mad.lo + mad.hi (two multiplier issues
// per product, PTX197:L27030-).  Reduction uses 2^256 = 2^32 + 977 (mod p), folded twice
// (cf. lib/Curve64.pb:1330-1434 for the CPU twin of the same identity).
//
// Values are "lazy": any 256-bit value congruent to the intended element; fe_mul/fe_sqr accept lazy
// inputs and produce lazy outputs.  Canonical (< p) values are produced only where the algorithm
// needs them (see comments at the call sites in giant_step.cuh).
//
// The file also compiles for the host (plain C++): the PTX primitives are then emulated with an
// explicit carry flag.  That build is used only by the CPU unit tests of this arithmetic
// (tests/test_fp_hostemu.py); it is not a fallback for any product path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FP_DEV __device__ __forceinline__
#define FP_DEV_NOINLINE __device__ __noinline__
#else
#define FP_DEV static inline
#define FP_DEV_NOINLINE static
#endif

namespace fp {

#if defined(__CUDA_ARCH__)
// (hi:lo) += a*b, carry-out in CC
FP_DEV void madw_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) += a*b + CC, carry-out in CC
FP_DEV void madwc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) = a*b  (no carry in), carry flag untouched afterwards is "no carry"
FP_DEV void mulw(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
FP_DEV void add_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FP_DEV void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FP_DEV void addc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FP_DEV void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FP_DEV void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FP_DEV void subc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
#else
// ---- host emulation of the same primitives (unit tests only) ----
static thread_local uint32_t g_cc = 0;
static uint64_t g_lost_carries = 0;      // carries dropped where the code claims none can occur (checked by the unit tests)
static uint64_t g_rare_calls = 0;        // executions of reduce512_rare (the unit tests make sure their crafted cases reach it)
FP_DEV void madw_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    unsigned __int128 t = (unsigned __int128)a * b + (((uint64_t)hi << 32) | lo);
    lo = (uint32_t)t; hi = (uint32_t)(t >> 32); g_cc = (uint32_t)(t >> 64);
}
FP_DEV void madwc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    unsigned __int128 t = (unsigned __int128)a * b + (((uint64_t)hi << 32) | lo) + g_cc;
    lo = (uint32_t)t; hi = (uint32_t)(t >> 32); g_cc = (uint32_t)(t >> 64);
}
FP_DEV void mulw(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
FP_DEV void add_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; r = (uint32_t)t; g_cc = (uint32_t)(t >> 32); }
FP_DEV void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; r = (uint32_t)t; g_cc = (uint32_t)(t >> 32); }
FP_DEV void addc(uint32_t& r, uint32_t a, uint32_t b) { r = a + b + g_cc; }
FP_DEV void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; r = (uint32_t)t; g_cc = (uint32_t)(t >> 63); }
FP_DEV void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; r = (uint32_t)t; g_cc = (uint32_t)(t >> 63); }
FP_DEV void subc(uint32_t& r, uint32_t a, uint32_t b) { r = a - b - g_cc; }
#endif

#ifndef FP_SQR_LINK
#define FP_SQR_LINK 1
#endif
#ifndef FP_LINK
#define FP_LINK 0
#endif

// p = 2^256 - 2^32 - 977, little-endian limbs
#define FP_P0 0xFFFFFC2Fu
#define FP_P1 0xFFFFFFFEu
#define FP_C0 977u            // 2^256 mod p = 2^32 + 977

// acc[0..7] (four 64-bit slots) += a[0],a[2],a[4],a[6] * b, one unbroken carry chain.
//
// Carry out of the top slot: when the top slot (acc[6], acc[7]) has not been written before ("fresh",
// still 0:0), the last product cannot overflow it: a*b + carry <= (2^32-1)^2 + 1 < 2^64.  Such rows
// need no carry capture at all (CAPTURE = false); fe_mul orders its rows so that 9 of the 16 are of that
// kind.  For the other 7 the carry is stored (not accumulated: each of those words receives exactly
// one carry) in `cw`, a separate word of the weight of the limb above the top slot.
template <bool CAPTURE, bool LINKED>
FP_DEV void mad_row(uint32_t* acc, const uint32_t* a, uint32_t b, uint32_t& cw) {
    if (LINKED) madwc_cc(acc[0], acc[1], a[0], b);     // takes the (zero) carry of the previous row: orders the rows, changes no value
    else madw_cc(acc[0], acc[1], a[0], b);
    madwc_cc(acc[2], acc[3], a[2], b);
    madwc_cc(acc[4], acc[5], a[4], b);
    madwc_cc(acc[6], acc[7], a[6], b);
    if (CAPTURE) addc_cc(cw, 0, 0);                    // cw = carry, and the flag is clear again for a linked successor
#if !defined(__CUDA_ARCH__)
    else g_lost_carries += g_cc;
#endif
}

// Rare tail of reduce512 (about 2^-22 of the calls), out of line: r holds (s mod 2^256) + top*C in limbs
// 0..2 with the carry k into limb 3 still pending, and top2*2^32*C = top2*(977*2^32 + 2^64) not yet added.
// Every wrap of 2^256 is worth another + C.
struct Fe8 { uint32_t v[8]; };
FP_DEV_NOINLINE Fe8 reduce512_rare(Fe8 x, uint32_t k, uint32_t top2) {
    uint32_t* r = x.v;
    uint32_t w, w2;
#if !defined(__CUDA_ARCH__)
    g_rare_calls++;
#endif
    add_cc(r[3], r[3], k);
#pragma unroll
    for (int i = 4; i < 8; i++) addc_cc(r[i], r[i], 0);
    addc(w, 0, 0);
    if (top2) {
        add_cc(r[1], r[1], FP_C0); addc_cc(r[2], r[2], 1);
#pragma unroll
        for (int i = 3; i < 8; i++) addc_cc(r[i], r[i], 0);
        addc(w2, 0, 0);
        w += w2;
    }
    while (w) {                                    // at most twice
        w--;
        add_cc(r[0], r[0], FP_C0); addc_cc(r[1], r[1], 1);
#pragma unroll
        for (int i = 2; i < 8; i++) addc_cc(r[i], r[i], 0);
        addc(w2, 0, 0);
        w += w2;
    }
    return x;
}

// r = v[0..15] mod p (lazy, < 2^256): fold the high half by 2^256 = 2^32 + 977 =: C (mod p) twice.
//
// First fold: s = lo + hi*977 + (hi << 32), computed as two carry chains of IMAD.WIDE (even limbs of
// hi onto lo, odd limbs onto hi << 32) and one 9-limb addition.  Its overflow T = s >> 256 satisfies
// T <= 2^32 + 977 (lo, hi < 2^256), i.e. T = top2*2^32 + top with top2 in {0, 1}.
// Second fold: r = (s mod 2^256) + T*C.  top*C < 2^65 is added to limbs 0..2 only; the carry out of limb 2
// (it needs limb 2 of s to be all ones: 2^-32) and top2 (it needs hi[7] within ~977 of 2^32) are left to
// the out-of-line exact tail above.
FP_DEV void reduce512(uint32_t r[8], uint32_t v[16]) {
    uint32_t t[10];
    t[0] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t[i + 1] = v[8 + i];
    t[9] = 0;
    madw_cc(t[1], t[2], v[9], FP_C0);
    madwc_cc(t[3], t[4], v[11], FP_C0);
    madwc_cc(t[5], t[6], v[13], FP_C0);
    madwc_cc(t[7], t[8], v[15], FP_C0);
    addc(t[9], 0, 0);
    uint32_t l8 = 0;
    madw_cc(v[0], v[1], v[8], FP_C0);
    madwc_cc(v[2], v[3], v[10], FP_C0);
    madwc_cc(v[4], v[5], v[12], FP_C0);
    madwc_cc(v[6], v[7], v[14], FP_C0);
    addc(l8, 0, 0);
    uint32_t top, top2;
    r[0] = v[0];                                   // t[0] == 0
    add_cc(r[1], v[1], t[1]);
#pragma unroll
    for (int i = 2; i < 8; i++) addc_cc(r[i], v[i], t[i]);
    addc_cc(top, l8, t[8]);
    addc(top2, t[9], 0);
    // f = top*977 + (top << 32) as limbs f0..f2, one IMAD.WIDE: (f1:f0) = top*977 + (top << 32 | 0)
    uint32_t f0 = 0, f1 = top, f2, k;
    madw_cc(f0, f1, top, FP_C0);
    addc(f2, 0, 0);
    add_cc(r[0], r[0], f0); addc_cc(r[1], r[1], f1); addc_cc(r[2], r[2], f2);
    addc(k, 0, 0);
#if defined(__CUDA_ARCH__)
    if (__builtin_expect((k | top2) != 0u, 0)) {
#else
    if ((k | top2) != 0u) {
#endif
        Fe8 x;
#pragma unroll
        for (int i = 0; i < 8; i++) x.v[i] = r[i];
        x = reduce512_rare(x, k, top2);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = x.v[i];
    }
}

// r = a*b mod p (lazy in, lazy out).  64 + 9 multiplier issues.
//
// Rows (i = 0, 2, 4, 6), e[k] = limb k of the even-column accumulator, o[k] = limb k + 1 of the odd one:
//   R1  e[i   .. i+7] += a_even * b[i]      top slot e[i+6,i+7]: fresh for i = 0, written by R4(i-2) otherwise -> cw[i]   (limb i+8)
//   R2  o[i   .. i+7] += a_odd  * b[i]      top slot o[i+6,i+7]: always fresh
//   R3  o[i   .. i+7] += a_even * b[i+1]    same slots as R2: never fresh                                    -> cw[i+1] (limb i+9)
//   R4  e[i+2 .. i+9] += a_odd  * b[i+1]    top slot e[i+8,i+9]: always fresh
// i.e. seven captured carries cw[1..7], one per word.
FP_DEV void fe_mul(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t e[16], o[16];
    uint32_t cw[8];
#pragma unroll
    for (int i = 0; i < 16; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i++) cw[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        // FP_LINK 0: rows independent; 1: every row takes the zero carry left by its predecessor (one dependency
        // chain, one live carry predicate; ptxas otherwise interleaves many rows and spills predicates)
        // (FP_LINK 2: only the pairs on the same accumulator are linked: R2 -> R3 and R4 -> R1 of the next round)
        if (i == 0) mad_row<false, false>(e + i, a, b[i], cw[0]);
        else mad_row<true, FP_LINK != 0>(e + i, a, b[i], cw[i]);
        mad_row<false, FP_LINK == 1>(o + i, a + 1, b[i], cw[0]);
        mad_row<true, FP_LINK != 0>(o + i, a, b[i + 1], cw[i + 1]);
        mad_row<false, FP_LINK == 1>(e + i + 2, a + 1, b[i + 1], cw[0]);
    }
    // v = e + (o << 32) + (cw << 256)
    uint32_t v[16];
    v[0] = e[0];
    add_cc(v[1], e[1], o[0]);
#pragma unroll
    for (int i = 2; i < 15; i++) addc_cc(v[i], e[i], o[i - 1]);
    addc(v[15], e[15], o[14]);
    add_cc(v[9], v[9], cw[1]);
#pragma unroll
    for (int i = 2; i < 7; i++) addc_cc(v[8 + i], v[8 + i], cw[i]);
    addc(v[15], v[15], cw[7]);
    reduce512(r, v);
}

// r = a^2 mod p.  36 + 9 multiplier issues: 28 cross products, doubled, plus 8 diagonal squares.
//
// The cross products a_i*a_j (i < j) are taken along NESTED chains of consecutive slots, shortest chain
// first, so that the top slot of every chain is fresh when its last product arrives and no chain can
// carry out (see mad_row): not one carry is captured.  Even columns (slot i+j): 1,2,3,3,2,1 products in
// slots 2..12 = chains [6..8], [4..10], [2..12]; odd columns (o-slot i+j-1): 1,2,3,4,3,2,1 products in
// slots 0..12 = chains [6], [4..8], [2..10], [0..12].
FP_DEV void fe_sqr_wide(uint32_t v[16], const uint32_t a[8]) {      // v = a^2, all 512 bits
    uint32_t e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { e[i] = 0; o[i] = 0; }
#define SQ_O(i, j, first) do { if (first) madw_cc(o[(i) + (j) - 1], o[(i) + (j)], a[i], a[j]); else madwc_cc(o[(i) + (j) - 1], o[(i) + (j)], a[i], a[j]); } while (0)
#define SQ_E(i, j, first) do { if (first) madw_cc(e[(i) + (j)], e[(i) + (j) + 1], a[i], a[j]); else madwc_cc(e[(i) + (j)], e[(i) + (j) + 1], a[i], a[j]); } while (0)
#if defined(__CUDA_ARCH__)
#define SQ_END
#else
#define SQ_END g_lost_carries += g_cc;
#endif
    // Chains of one parity are LINKED: the first product of the next chain takes the (provably zero) carry of
    // the previous one.  That changes no value, but it makes the cross products of one parity a single
    // dependency chain with one live carry; left independent, ptxas interleaves all seven chains and spills
    // carry predicates into a register (two LOP3 per spill).
#define SQ_F (FP_SQR_LINK == 0)     /* first product of a chain after the first: fresh flag (1) or linked (0) */
    SQ_E(2, 4, 1); SQ_E(3, 5, 0); SQ_END                                                                  // [6..8]
    SQ_E(1, 3, SQ_F); SQ_E(1, 5, 0); SQ_E(2, 6, 0); SQ_E(4, 6, 0); SQ_END                                 // [4..10]
    SQ_E(0, 2, SQ_F); SQ_E(0, 4, 0); SQ_E(0, 6, 0); SQ_E(1, 7, 0); SQ_E(3, 7, 0); SQ_E(5, 7, 0); SQ_END   // [2..12]
    SQ_O(3, 4, 1); SQ_END                                                                                 // [6]
    SQ_O(2, 3, SQ_F); SQ_O(2, 5, 0); SQ_O(4, 5, 0); SQ_END                                                // [4..8]
    SQ_O(1, 2, SQ_F); SQ_O(1, 4, 0); SQ_O(1, 6, 0); SQ_O(3, 6, 0); SQ_O(5, 6, 0); SQ_END                  // [2..10]
    SQ_O(0, 1, SQ_F); SQ_O(0, 3, 0); SQ_O(0, 5, 0); SQ_O(0, 7, 0); SQ_O(2, 7, 0); SQ_O(4, 7, 0); SQ_O(6, 7, 0); SQ_END   // [0..12]
#undef SQ_F
#undef SQ_END
#undef SQ_O
#undef SQ_E
    // t = e + (o << 32): limbs 1..15 (e[0] = e[1] = 0; the cross sum is < 2^511)
    uint32_t t[16];
    t[0] = 0; t[1] = o[0];
    add_cc(t[2], e[2], o[1]);
#pragma unroll
    for (int i = 3; i < 15; i++) addc_cc(t[i], e[i], o[i - 1]);
    addc(t[15], 0, 0);
    // v = 2t (funnel shifts: no carry chain), then + diagonal as one unbroken chain of eight 64-bit slots
    v[0] = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) v[i] = (t[i] << 1) | (t[i - 1] >> 31);
    madw_cc(v[0], v[1], a[0], a[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) madwc_cc(v[2 * i], v[2 * i + 1], a[i], a[i]);
}
FP_DEV void fe_sqr(uint32_t r[8], const uint32_t a[8]) {
    uint32_t v[16];
    fe_sqr_wide(v, a);
    reduce512(r, v);
}

// r = a - b mod p for CANONICAL a, b (< p): one conditional +p.  Result canonical.
FP_DEV void fe_sub_canon(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    sub_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) subc_cc(r[i], a[i], b[i]);
    uint32_t br; subc(br, 0, 0);                   // 0 or 0xFFFFFFFF
    // += p  ==  -= (2^32 + 977) modulo 2^256 when a borrow happened
    uint32_t c0 = br & FP_C0, c1 = br & 1u;
    sub_cc(r[0], r[0], c0); subc_cc(r[1], r[1], c1);
#pragma unroll
    for (int i = 2; i < 8; i++) subc_cc(r[i], r[i], 0);
}

// r = a + b for canonical a, b, result lazy (< 2^256) : if the sum wraps 2^256 add 2^32+977.
FP_DEV void fe_add_lazy(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    add_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) addc_cc(r[i], a[i], b[i]);
    uint32_t k; addc(k, 0, 0);
    uint32_t c0 = (0u - k) & FP_C0;
    add_cc(r[0], r[0], c0); addc_cc(r[1], r[1], k);
#pragma unroll
    for (int i = 2; i < 8; i++) addc_cc(r[i], r[i], 0);
}

// in-place canonicalisation of a lazy value (< 2^256): subtract p once if >= p.
FP_DEV void fe_canon(uint32_t r[8]) {
    uint32_t hiall = r[2] & r[3] & r[4] & r[5] & r[6] & r[7];
    if (hiall == 0xFFFFFFFFu && (r[1] == 0xFFFFFFFFu || (r[1] == FP_P1 && r[0] >= FP_P0))) {
        // r - p = r + (2^32 + 977) - 2^256
        add_cc(r[0], r[0], FP_C0); addc_cc(r[1], r[1], 1);
#pragma unroll
        for (int i = 2; i < 8; i++) addc_cc(r[i], r[i], 0);
    }
}

FP_DEV bool fe_is_zero(const uint32_t a[8]) { return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0; }
// same predicate, arranged so that the common case costs one compare and a never-taken branch
FP_DEV bool fe_is_zero_rare(const uint32_t a[8]) {
#if defined(__CUDA_ARCH__)
    if (__builtin_expect(a[0] != 0u, 1)) return false;
#endif
    return fe_is_zero(a);
}
FP_DEV bool fe_eq(const uint32_t a[8], const uint32_t b[8]) {
    return ((a[0] ^ b[0]) | (a[1] ^ b[1]) | (a[2] ^ b[2]) | (a[3] ^ b[3]) | (a[4] ^ b[4]) | (a[5] ^ b[5]) | (a[6] ^ b[6]) | (a[7] ^ b[7])) == 0;
}
FP_DEV void fe_copy(uint32_t r[8], const uint32_t a[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a[i];
}

// Low 64 bits of the CANONICAL value of (a + u) mod p, a lazy (< 2^256), u canonical (< p).
// z = (a + u) mod 2^256 with carry k; the value is z + k*(2^32 + 977).  Its low 64 bits are
// (z mod 2^64) + k*(2^32 + 977) whatever happens above them, so the fast path needs nothing else --
// provided the value is already canonical and did not wrap again, which is certain when the top limb
// of z is not all ones (then z + k*C < 2^256 - 2^224 + 2^33 < p).  Otherwise (probability 2^-32) the
// exact general case runs.
FP_DEV uint64_t fe_add_low64_canon(const uint32_t a[8], const uint32_t u[8]) {
    uint32_t z[8];
    add_cc(z[0], a[0], u[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) addc_cc(z[i], a[i], u[i]);
    uint32_t k; addc(k, 0, 0);
    if (z[7] == 0xFFFFFFFFu) {                     // exact slow path
        uint32_t c0 = (0u - k) & FP_C0;
        add_cc(z[0], z[0], c0); addc_cc(z[1], z[1], k);
#pragma unroll
        for (int i = 2; i < 8; i++) addc_cc(z[i], z[i], 0);
        uint32_t k2; addc(k2, 0, 0);
        if (k2) {                                  // cannot happen for u < p, kept for exactness
            add_cc(z[0], z[0], FP_C0); addc_cc(z[1], z[1], 1);
#pragma unroll
            for (int i = 2; i < 8; i++) addc_cc(z[i], z[i], 0);
        }
        fe_canon(z);
        return ((uint64_t)z[1] << 32) | z[0];
    }
    uint32_t c0 = (0u - k) & FP_C0, l0, l1;
    add_cc(l0, z[0], c0); addc(l1, z[1], k);
    return ((uint64_t)l1 << 32) | l0;
}

// Low 64 bits of the CANONICAL value of (L + A - Q) mod p without forming the 256-bit sum:
// L lazy (< 2^256), A and Q canonical (< p).  (The kernel uses it with L = lambda^2, A = p - Px, Q = Qx:
// X = lambda^2 - Px - Qx.)
//
// Write V = L + A - Q = S*2^224 + R with S = L[7] + A[7] - Q[7] (top limbs, -(2^32-1) <= S <= 2^33-2)
// and R the same expression on the low 224 bits, -2^224 < R < 2^225.  Then
//     S <= -2                      =>  V < 0,                 V mod p = V + p = V + 2^256 - C
//     1 <= S <= 2^32 - 3           =>  0 <= V < 2^256 - 2^224 < p, V mod p = V
//     2^32 + 1 <= S <= 2^33 - 4    =>  2^256 <= V < 2^256 + p - 2^224... , V mod p = V - 2^256 + C
// (C = 2^32 + 977), and 2^256 does not reach the low 64 bits: key = (L + A - Q mod 2^64) + K*C with
// K = -1, 0, +1.  For the seven remaining values of S (probability about 2^-29) the caller takes the
// exact path (fe_sub_canon + fe_add_low64_canon).  t_lo = (A - Q) mod 2^64 and t7 = A[7] - Q[7] are
// shared by the two candidates of a point-window.
FP_DEV uint64_t fe_key_low64(const uint32_t L[8], uint64_t t_lo, int64_t t7, bool& exact_needed) {
    const int64_t S = (int64_t)L[7] + t7;
    const int64_t B = (int64_t)1 << 32;
    const uint64_t lo = ((((uint64_t)L[1]) << 32) | L[0]) + t_lo;
    exact_needed = (S >= -1 && S <= 0) || (S >= B - 2 && S <= B) || (S >= 2 * B - 3);
    const uint64_t C = ((uint64_t)1 << 32) + FP_C0;
    uint64_t key = lo;
    if (S > B) key = lo + C;
    if (S < 0) key = lo - C;
    return key;
}

// Low 64 bits of the CANONICAL value of (v + A - Q) mod p for an UNREDUCED 512-bit v = hi*2^256 + lo (the
// kernel: v = lambda^2, A = p - Px, Q = Qx, both canonical), without reducing v: the candidate X is only ever
// used through these 64 bits, so the squarings of PASS 2 need no reduce512 (9 multiplier issues + ~25
// other instructions each).  t_lo = (A - Q) mod 2^64, t7 = A[7] - Q[7] as for fe_key_low64.
//
// With C = 2^32 + 977 (2^256 = C mod p):  v + A - Q = U := hi*C + lo + A - Q (mod p), -2^256 < U < 2^256 (C + 2).
// Write U = m*2^256 + rho, 0 <= rho < 2^256.  Then U = m*p + (rho + m*C), and whenever the top limb of rho is
// neither 0 nor 2^32 - 1, 0 <= rho + m*C < p, i.e. the canonical value is rho + m*C and
//     key = (rho + m*C) mod 2^64 = lo01 + t_lo + (hi01 + m)*C  (mod 2^64),   x01 = x mod 2^64.
// m and the top limb of rho come from the limbs of weight >= 2^224 alone:
//     U = 2^224 * (M*2^32 + S7) + R,   S7 = lo[7] + t7 + hi[6] + lo32(977 hi[7]) + hi32(977 hi[6]),
//                                      M  = hi[7] + hi32(977 hi[7]),
// where R collects everything below 2^224 of the five terms and -2^224 < R < 5 * 2^224, a carry c in [-1, 4].
// So if s = S7 mod 2^32 lies in [2, 2^32 - 6], m = M + floor(S7 / 2^32) exactly and the top limb of rho is
// s + c in [1, 2^32 - 2].  Otherwise (8 values of s: probability 2^-29) the caller takes its exact path.
// `thr` is the acceptance threshold 2^32 - 8 (a run-time value so that tests can force the exact path: thr = 0
// accepts only s = 2).
FP_DEV uint64_t fe_key_from_wide(const uint32_t v[16], uint64_t t_lo, int64_t t7, uint32_t thr, bool& exact_needed) {
    const uint64_t C = ((uint64_t)1 << 32) + FP_C0;
    const uint64_t lo01 = ((uint64_t)v[1] << 32) | v[0], hi01 = ((uint64_t)v[9] << 32) | v[8];
    const uint64_t w = (uint64_t)v[15] * FP_C0;                        // 977 * hi[7] < 2^42
    const uint32_t g = (uint32_t)(((uint64_t)v[14] * FP_C0) >> 32);    // hi32(977 * hi[6])
    const int64_t S7 = (int64_t)((uint64_t)v[7] + (uint64_t)v[14] + (uint64_t)(uint32_t)w + (uint64_t)g) + t7;
    const uint64_t M = (uint64_t)v[15] + (w >> 32);
    const uint32_t s = (uint32_t)S7;
    exact_needed = (uint32_t)(s - 2u) > thr;
    const uint64_t m = M + (uint64_t)(S7 >> 32);                       // floor division: arithmetic shift
    return lo01 + t_lo + (hi01 + m) * C;
}
#define FP_KEY_THR 0xFFFFFFF8u

// a^(p-2) mod p with the 255-squaring / 15-multiplication addition chain for
// p - 2 = 2^256 - 2^32 - 979 (the reference kernel uses plain right-to-left binary
// exponentiation with 255 + 249 generic multiplications, PTX197:L1319-26062).
FP_DEV_NOINLINE void fe_mul_ni(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) { fe_mul(r, a, b); }
FP_DEV_NOINLINE void fe_sqr_n_ni(uint32_t r[8], int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) { uint32_t t[8]; fe_sqr(t, r); fe_copy(r, t); }
}
FP_DEV_NOINLINE void fe_inv(uint32_t r[8], const uint32_t a[8]) {
    uint32_t x2[8], x3[8], x22[8], x44[8], t[8], u[8];
    fe_copy(t, a); fe_sqr_n_ni(t, 1); fe_mul_ni(x2, t, a);
    fe_copy(t, x2); fe_sqr_n_ni(t, 1); fe_mul_ni(x3, t, a);
    fe_copy(t, x3); fe_sqr_n_ni(t, 3); fe_mul_ni(u, t, x3);          // x6
    fe_copy(t, u); fe_sqr_n_ni(t, 3); fe_mul_ni(u, t, x3);           // x9
    fe_copy(t, u); fe_sqr_n_ni(t, 2); fe_mul_ni(u, t, x2);           // x11
    fe_copy(t, u); fe_sqr_n_ni(t, 11); fe_mul_ni(x22, t, u);         // x22
    fe_copy(t, x22); fe_sqr_n_ni(t, 22); fe_mul_ni(x44, t, x22);     // x44
    fe_copy(t, x44); fe_sqr_n_ni(t, 44); fe_mul_ni(u, t, x44);       // x88
    fe_copy(t, u); fe_sqr_n_ni(t, 88); fe_mul_ni(u, t, u);           // x176  (t = x88^(2^88), u = x88)
    fe_copy(t, u); fe_sqr_n_ni(t, 44); fe_mul_ni(u, t, x44);         // x220
    fe_copy(t, u); fe_sqr_n_ni(t, 3); fe_mul_ni(u, t, x3);           // x223
    fe_copy(t, u); fe_sqr_n_ni(t, 23); fe_mul_ni(u, t, x22);
    fe_copy(t, u); fe_sqr_n_ni(t, 5); fe_mul_ni(u, t, a);
    fe_copy(t, u); fe_sqr_n_ni(t, 3); fe_mul_ni(u, t, x2);
    fe_copy(t, u); fe_sqr_n_ni(t, 2); fe_mul_ni(r, t, a);
}

}  // namespace fp
