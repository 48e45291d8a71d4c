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

// p = 2^256 - 2^32 - 977, little-endian limbs
#define FP_P0 0xFFFFFC2Fu
#define FP_P1 0xFFFFFFFEu
#define FP_C0 977u            // 2^256 mod p = 2^32 + 977

// acc[0..7] (four 64-bit slots) += a[0],a[2],a[4],a[6] * b; the carry out of the top slot is counted in
// `cw`, a separate word of weight acc[8] -- NOT added into acc[8] itself, so that the top slot of the
// next row is still untouched when that row reaches it and its product needs no 64-bit addend
// (an addend (carry, 0) costs a register move on the multiplier pipe for the zero half).
FP_DEV void mad_row(uint32_t* acc, const uint32_t* a, uint32_t b, uint32_t& cw) {
    madw_cc(acc[0], acc[1], a[0], b);
    madwc_cc(acc[2], acc[3], a[2], b);
    madwc_cc(acc[4], acc[5], a[4], b);
    madwc_cc(acc[6], acc[7], a[6], b);
    addc(cw, cw, 0);
}

// r = v[0..15] mod p (lazy, < 2^256): fold the high half by 2^256 = 2^32 + 977 (mod p) twice.
//
// First fold: s = lo + hi*977 + (hi << 32), computed as two carry chains of IMAD.WIDE (even limbs of
// hi onto lo, odd limbs onto hi << 32) and one 9-limb addition.  Its overflow T = s >> 256 satisfies
// T <= 2^32 + 977 (lo, hi < 2^256), i.e. T = top2*2^32 + top with top2 in {0, 1}.
// Second fold: r = (s mod 2^256) + T*(2^32 + 977) = ... + top*977 + (top << 32) + top2*(977 << 32) + (top2 << 64);
// this addend is < 2^67.  If that addition wraps 2^256 once more, the wrapped value is < 2^67, so
// the final "+ 2^32 + 977" touches limbs 0..2 only.
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
    uint32_t f0 = 0, f1 = top, f2;
    madw_cc(f0, f1, top, FP_C0);
    addc(f2, 0, 0);
    // top2 part: (977 << 32) + (1 << 64), only when top2 == 1 (then top <= 977: no further carries)
    uint32_t m = 0u - top2;
    add_cc(f1, f1, m & FP_C0); addc(f2, f2, top2);
    add_cc(r[0], r[0], f0); addc_cc(r[1], r[1], f1); addc_cc(r[2], r[2], f2);
#pragma unroll
    for (int i = 3; i < 8; i++) addc_cc(r[i], r[i], 0);
    uint32_t k; addc(k, 0, 0);
    if (k) {                                       // wrapped once more (probability ~2^-190): r < 2^67 here
        add_cc(r[0], r[0], FP_C0); addc_cc(r[1], r[1], 1); addc(r[2], r[2], 0);
    }
}

// r = a*b mod p (lazy in, lazy out).  64 + 9 multiplier issues.
FP_DEV void fe_mul(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t e[16], o[16];                         // e[k]: limb k ; o[k]: limb k+1
    uint32_t cw[9];                                // row carries: cw[k] has the weight of limb 8 + k (cw[8]: limb 16, always 0)
#pragma unroll
    for (int i = 0; i < 16; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll
    for (int i = 0; i < 9; i++) cw[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mad_row(e + i, a, b[i], cw[i]);                    // a_even * b_even  -> even columns, carry at limb i + 8
        mad_row(o + i, a + 1, b[i], cw[i + 1]);            // a_odd  * b_even  -> odd columns,  carry at limb i + 9
        mad_row(o + i, a, b[i + 1], cw[i + 1]);            // a_even * b_odd   -> odd columns,  carry at limb i + 9
        mad_row(e + i + 2, a + 1, b[i + 1], cw[i + 2]);    // a_odd  * b_odd   -> even columns, carry at limb i + 10
    }
    // v = e + (o << 32) + (cw << 256)
    uint32_t v[16];
    v[0] = e[0];
    add_cc(v[1], e[1], o[0]);
#pragma unroll
    for (int i = 2; i < 15; i++) addc_cc(v[i], e[i], o[i - 1]);
    addc(v[15], e[15], o[14]);
    add_cc(v[8], v[8], cw[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(v[8 + i], v[8 + i], cw[i]);
    addc(v[15], v[15], cw[7]);
    reduce512(r, v);
}

// r = a^2 mod p.  36 + 9 multiplier issues: 28 cross products, doubled, plus 8 diagonal squares.
FP_DEV void fe_sqr(uint32_t r[8], const uint32_t a[8]) {
    uint32_t e[18], o[18];
#pragma unroll
    for (int i = 0; i < 18; i++) { e[i] = 0; o[i] = 0; }
    // cross products a_i*a_j, i<j.  (i+j) odd -> o[i+j-1..], (i+j) even -> e[i+j..]
    // row i, j = i+1, i+3, ...  (odd weight)
#define SQ_O(i, j, first) do { if (first) madw_cc(o[(i) + (j) - 1], o[(i) + (j)], a[i], a[j]); else madwc_cc(o[(i) + (j) - 1], o[(i) + (j)], a[i], a[j]); } while (0)
#define SQ_E(i, j, first) do { if (first) madw_cc(e[(i) + (j)], e[(i) + (j) + 1], a[i], a[j]); else madwc_cc(e[(i) + (j)], e[(i) + (j) + 1], a[i], a[j]); } while (0)
    SQ_O(0, 1, 1); SQ_O(0, 3, 0); SQ_O(0, 5, 0); SQ_O(0, 7, 0); addc(o[8], o[8], 0);
    SQ_E(0, 2, 1); SQ_E(0, 4, 0); SQ_E(0, 6, 0); addc(e[8], e[8], 0);
    SQ_O(1, 2, 1); SQ_O(1, 4, 0); SQ_O(1, 6, 0); addc(o[8], o[8], 0);
    SQ_E(1, 3, 1); SQ_E(1, 5, 0); SQ_E(1, 7, 0); addc(e[10], e[10], 0);
    SQ_O(2, 3, 1); SQ_O(2, 5, 0); SQ_O(2, 7, 0); addc(o[10], o[10], 0);
    SQ_E(2, 4, 1); SQ_E(2, 6, 0); addc(e[10], e[10], 0);
    SQ_O(3, 4, 1); SQ_O(3, 6, 0); addc(o[10], o[10], 0);
    SQ_E(3, 5, 1); SQ_E(3, 7, 0); addc(e[12], e[12], 0);
    SQ_O(4, 5, 1); SQ_O(4, 7, 0); addc(o[12], o[12], 0);
    SQ_E(4, 6, 1); addc(e[12], e[12], 0);
    SQ_O(5, 6, 1); addc(o[12], o[12], 0);
    SQ_E(5, 7, 1); addc(e[14], e[14], 0);
    SQ_O(6, 7, 1); addc(o[14], o[14], 0);
#undef SQ_O
#undef SQ_E
    // v = e + (o << 32), then v = 2v
    uint32_t v[16];
    v[0] = e[0];
    add_cc(v[1], e[1], o[0]);
#pragma unroll
    for (int i = 2; i < 16; i++) addc_cc(v[i], e[i], o[i - 1]);
    add_cc(v[0], v[0], v[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) addc_cc(v[i], v[i], v[i]);
    // + diagonal: one unbroken chain of eight 64-bit slots
    madw_cc(v[0], v[1], a[0], a[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) madwc_cc(v[2 * i], v[2 * i + 1], a[i], a[i]);
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
