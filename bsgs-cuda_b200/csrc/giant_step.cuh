// giant_step.cuh -- the giant-step kernel for sm_100a and its table-layout helpers.
//
// One launch evaluates, for W window base points P_w (W <= 16, a run-time value) and every giant
// point Q_k = (k+1)*(-2w*G), k < N, the candidates X(P_w - Q_k) and X(P_w + Q_k) and probes each in
// the baby-step table -- the work of W launches of the reference kernel `_test1`
// (1_9_7File.pb:5181-23979; PASS 1 PTX197:L149-1297, inversion L1319-26062, PASS 2 L26955-36020).
//
// Structure (fourth iteration; the measurements that led here are in DESIGN.md 4.1 and profiles/):
//   * limb products are single IMAD.WIDE.U32 issues with predicate carries (fp.cuh) instead of the
//     reference's mad.lo + mad.hi pairs; the IMAD.WIDE pipe (measured 30.7 lanes/clk/SM) is the
//     roofline of this kernel;
//   * one Fermat inversion per thread covers PP*W points (reference: p points) and uses the
//     255S+15M addition chain (reference: 255 + 249 generic multiplications);
//   * rows outside, windows inside, ONE loop body: window constants live in shared memory and the
//     hot loop holds exactly 4 inlined multiplications + 2 squarings, so that it stays inside the
//     instruction cache (the first kernel unrolled the windows: 118 KB of code, 35 % of stall
//     samples were instruction-fetch);
//   * the candidate's square is not reduced: X is used only through the low 64 bits of its canonical
//     value, which follow from a few limbs of the 512-bit square (fp::fe_key_from_wide); 2^-29 of the
//     candidates take an out-of-line exact path;
//   * the warps of one launch do not finish together (the first warp of an SM is done after 58 % of the
//     kernel): the library keeps two launches in flight on two streams (bsgs_b200.cu, "lanes"), so this
//     kernel is written to share an SM with CTAs of its successor -- own chain and record buffers per lane;
//   * giant coordinates are 128-bit vectors, SoA across threads: every warp load is one fully
//     coalesced 512-byte transaction;
//   * the CSR table of the reference (offset pair -> binary search: two dependent random reads) is
//     re-laid out at upload into one 32-byte sector per bucket holding fifteen 16-bit fingerprints
//     (from the top half of the 32-bit tag, see fpr16) and a count.  A probe is ONE 256-bit load
//     (LDG.E.256) and eight fp16x2 compares; only a fingerprint match (~2^-14 per entry) or an
//     over-full bucket (> 15 entries, ~1e-5 at load 4) consults the CSR image kept beside it, so
//     the emitted records are exactly the reference's;
//   * probe keys are parked in a per-thread ring in shared memory and resolved in bursts of four
//     independent sector loads, so no probe state is live across the arithmetic;
//   * everything rare (padding rows, a tangent Px == Qx, the ragged last burst, a fingerprint
//     match) goes through one out-of-line exact routine behind a warp vote, so that the hot loop
//     carries no per-slot bookkeeping.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "fp.cuh"

// GS_FP_HALF2 = 1: fingerprints are compared two at a time as fp16 pairs (HSET2.NE, one instruction
// per word instead of three integer ones).  0 keeps the integer zero-halfword test.
#ifndef GS_FP_HALF2
#define GS_FP_HALF2 1
#endif
// GS_ABLATE: timing experiments only (results are wrong): bit 0 = no probe drain, bit 1 = no chain
// traffic, bit 2 = giants read once per thread.  0 in every shipped build.
#ifndef GS_ABLATE
#define GS_ABLATE 0
#endif

// GS_PREFETCH: L2 prefetch of the chain element one iteration ahead (PASS 2 walks the chain backwards)
#ifndef GS_PREFETCH
#define GS_PREFETCH 1
#endif

namespace gs {

constexpr int kMaxW = 16;
#ifndef GS_RING
#define GS_RING 16
#endif
#ifndef GS_GROUP
#define GS_GROUP 4
#endif
constexpr int kRing = GS_RING;             // keys per thread between drains (2 per point-window)
constexpr int kGroup = GS_GROUP;           // independent sector loads in flight per thread during a drain

// per-window constants, all canonical, little-endian limbs
struct WinConst {
    uint32_t px[8], py[8];
    uint32_t npx[8];      // p - Px
    uint32_t py2[8];      // 2*Py mod p  (denominator of the tangent, used when Px == Qx)
};
struct WinParams {                             // 2.1 KB of kernel parameters
    WinConst win[kMaxW];
    uint64_t center[kMaxW];                    // emit mode only: P_w = center[w] * G
};

struct Tables {
    const uint4* gx;          // [PP*2][NT] : limbs 0-3 / 4-7 of Qx, slot (r, t) at (2r+h)*NT + t
    const uint4* gy;
    uint4* chain;             // [PP*W*2][NT]
    const uint32_t* buckets;  // H fingerprint sectors of 32 bytes
    const uint32_t* csr_off;  // H+1 offsets   } the htGPU image as uploaded: exact lookup behind
    const uint32_t* csr_tags; // w tags        } the fingerprints
    uint32_t* hit_count;      // [0] = number of records
    uint32_t* hits;           // records of 3 u32: code, idx, window
    uint32_t hit_cap;
    uint32_t NT, PP, N;       // threads in the grid, points per thread, giants
    uint32_t mask;            // H - 1
    uint32_t exact_thr;       // acceptance threshold of fp::fe_key_from_wide: FP_KEY_THR; 0 forces the exact path (tests)
    uint64_t* keys_out;       // emit mode: keys_out[m - 1] = low 64 bits of X(m*G), m <= nkeys
    uint64_t* trace;          // diagnostics (null in production): 4 words per warp = SM id, globaltimer at start / after pass 1 / at end
    uint64_t nkeys;
};

struct Fe { uint32_t v[8]; };

__device__ __forceinline__ void ld8_stream(uint32_t r[8], const uint4* p0, const uint4* p1) {
    uint4 a = __ldcs(p0), b = __ldcs(p1);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void st8_stream(uint4* p0, uint4* p1, const uint32_t r[8]) {
    __stcs(p0, make_uint4(r[0], r[1], r[2], r[3]));
    __stcs(p1, make_uint4(r[4], r[5], r[6], r[7]));
}

__device__ __forceinline__ void record_hit(const Tables& tb, uint32_t code, uint32_t idx, uint32_t win) {
    uint32_t pos = atomicAdd(tb.hit_count, 1u);
    if (pos < tb.hit_cap) { tb.hits[3 * pos] = code; tb.hits[3 * pos + 1] = idx; tb.hits[3 * pos + 2] = win; }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// CSR fallback == the reference probe (PTX197:L33723-34006; FOUNDINSORTNEW PTX173:L1447-1510)
__device__ __noinline__ bool probe_csr(const uint32_t* off, const uint32_t* tags, uint32_t h, uint32_t tag) {
    uint32_t lo = off[h], hi = off[h + 1];
    while (lo < hi) {
        uint32_t c = lo + ((hi - lo) >> 1), t = tags[c];      // offsets exceed 2^31 at w > 2^31: no lo + hi
        if (tag > t) lo = c + 1; else if (tag < t) hi = c; else return true;
    }
    return false;
}

// exact, slow: used only when Px == Qx (true doubling -> code 4), PTX197:L28996-
__device__ __noinline__ uint64_t double_x_low64(const uint32_t px[8], const uint32_t inv2y[8]) {
    uint32_t xx[8], t3[8], lam[8], X[8], t[8];
    fp::fe_sqr(xx, px); fp::fe_canon(xx);
    fp::fe_add_lazy(t, xx, xx); fp::fe_canon(t);
    fp::fe_add_lazy(t3, t, xx); fp::fe_canon(t3);
    fp::fe_mul(lam, t3, inv2y);
    fp::fe_sqr(X, lam); fp::fe_canon(X);
    fp::fe_sub_canon(t, X, px);
    fp::fe_sub_canon(X, t, px);
    return ((uint64_t)X[1] << 32) | X[0];
}

__device__ __forceinline__ Fe ldg_fe(const uint4* p0, const uint4* p1) {
    uint4 a = __ldg(p0), b = __ldg(p1);
    Fe r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ Fe lds_fe(const uint32_t* s) {
    uint4 a = *reinterpret_cast<const uint4*>(s), b = *reinterpret_cast<const uint4*>(s + 4);
    Fe r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- fingerprint sectors ------------------------------------------------------------------------
// 16 halfwords per bucket: fp[0..14] = fingerprint of the first 15 entries (0 when unused), last
// halfword = entry count, 0xFFFF when the bucket holds more than 15.
// Fingerprint of a 32-bit tag: its top 16 bits; in the half2 variant exponent bits 14..13 are forced
// to 01 so that the pattern, read as an fp16 number, is a normal number (exponent 8..15: never
// Inf/NaN, never zero or denormal, which the comparison may flush) -- bit-equal fingerprints then
// always compare equal and distinct ones never do; empty slots (0) and the count match nothing.
__host__ __device__ __forceinline__ uint32_t fpr16(uint32_t tag) {
#if GS_FP_HALF2
    return ((tag >> 16) & 0x9FFFu) | 0x2000u;
#else
    return tag >> 16;
#endif
}
__global__ void k_build_fp_sectors(const uint32_t* __restrict__ off, const uint32_t* __restrict__ tags,
                                   uint4* __restrict__ sectors, uint64_t H, unsigned long long* overfull) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    uint32_t lo = off[h], n = off[h + 1] - lo;
    uint32_t wbuf[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t a = (uint32_t)(2 * i) < n ? fpr16(tags[lo + 2 * i]) : 0u;
        uint32_t b = (uint32_t)(2 * i + 1) < n && i < 7 ? fpr16(tags[lo + 2 * i + 1]) : 0u;
        wbuf[i] = a | (b << 16);
    }
    wbuf[7] = (wbuf[7] & 0xFFFFu) | ((n <= 15u ? n : 0xFFFFu) << 16);
    if (n > 15u) atomicAdd(overfull, 1ull);
    sectors[2 * h] = make_uint4(wbuf[0], wbuf[1], wbuf[2], wbuf[3]);
    sectors[2 * h + 1] = make_uint4(wbuf[4], wbuf[5], wbuf[6], wbuf[7]);
}

struct FpSector {
    uint4 a, b;
    __device__ __forceinline__ void load(const Tables& tb, uint32_t h) {
        const uint4* p = reinterpret_cast<const uint4*>(tb.buckets) + 2ull * h;
        // One random 32-byte sector, never reused: a single 256-bit load (LDG.E.256) kept out of L1,
        // first in line for L2 eviction, and with the smallest L2 fill size (64 B).  Measured with ncu on
        // config 2: the default fill promotes every miss to the whole 128-byte line (90.7 GB of DRAM reads
        // per 16-window launch), the 64 B hint reads 51.4 GB at the same speed; two 128-bit loads instead
        // of one 256-bit load run 1.5x slower.
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
    }
    // true when the exact lookup is required: some halfword equals the fingerprint (this includes
    // chance matches with the count field or an empty slot -- harmless) or the bucket is over-full
    __device__ __forceinline__ bool maybe(uint32_t tag) const {
        const uint32_t f = fpr16(tag) * 0x00010001u;
#if GS_FP_HALF2
        // ordered "both halves differ" per word; the over-full marker 0xFFFF is a NaN, for which the
        // ordered comparison is false, so such a bucket always reports "maybe" (checked explicitly too)
#define FP_NE(x) __hbne2(*reinterpret_cast<const __half2*>(&(x)), *reinterpret_cast<const __half2*>(&f))
        const bool allne = FP_NE(a.x) & FP_NE(a.y) & FP_NE(a.z) & FP_NE(a.w) & FP_NE(b.x) & FP_NE(b.y) & FP_NE(b.z) & FP_NE(b.w);
#undef FP_NE
        return !allne;
#else
        uint32_t acc = 0;
#define FP_Z(x) { uint32_t y = (x) ^ f; acc |= (y - 0x00010001u) & ~y; }
        FP_Z(a.x) FP_Z(a.y) FP_Z(a.z) FP_Z(a.w) FP_Z(b.x) FP_Z(b.y) FP_Z(b.z) FP_Z(b.w)
#undef FP_Z
        return ((acc & 0x80008000u) != 0u) | (b.w >= 0xFFFF0000u);
#endif
    }
};

// exact, slow: Px == Qx (true doubling -> code 4), PTX197:L28996-
__device__ __noinline__ uint64_t double_x_low64_v(Fe px, Fe inv2y) { return double_x_low64(px.v, inv2y.v); }
// exact, slow: low 64 bits of canonical lambda^2 - Px - Qx for the few top-limb patterns that
// fp::fe_key_low64 does not decide (about 2^-29 of the candidates)
__device__ __noinline__ uint64_t key_exact(Fe l2, Fe npx, Fe qx) {
    uint32_t u[8];
    fp::fe_sub_canon(u, npx.v, qx.v);
    return fp::fe_add_low64_canon(l2.v, u);
}

// exact, slow: the key from lambda itself (full squaring with reduction, then the exact extraction)
__device__ __noinline__ uint64_t key_exact_sqr(Fe lam, Fe npx, Fe qx) {
    Fe l2;
    fp::fe_sqr(l2.v, lam.v);
    uint32_t u[8];
    fp::fe_sub_canon(u, npx.v, qx.v);
    return fp::fe_add_low64_canon(l2.v, u);
}
// The key of one candidate as the hot loop computes it: low 64 bits of canonical lambda^2 + npx - qx.  The
// square is taken WITHOUT reduction and the key read off its low 64 bits and the limbs of weight >= 2^224
// (fp::fe_key_from_wide); the few patterns that does not decide (2^-29 of the candidates, about 100 per second
// per GPU in production, essentially never at test sizes) go through the out-of-line exact path.  `npx`
// points at the eight limbs of p - Px (shared memory in the kernel).  The option `force_exact` (thr = 0)
// sends every candidate through the exact path, so that the whole parity suite can exercise it.
__device__ __forceinline__ uint64_t candidate_key(const Fe& lam, uint64_t t_lo, int64_t t7, const uint32_t* npx, const Fe& qx, uint32_t thr) {
    uint32_t v[16];
    fp::fe_sqr_wide(v, lam.v);
    bool exact_needed;
    uint64_t key = fp::fe_key_from_wide(v, t_lo, t7, thr, exact_needed);
    if (exact_needed) key = key_exact_sqr(lam, lds_fe(npx), qx);
    return key;
}

// Rare half of the drain: everything about ring slot `slot` of a burst is recomputed from its chain
// index cj = ci_first - slot/2 (row cj / W, window cj % W) -- whether the row is padding (giant index
// >= N), whether the pair was a tangent (Px == Qx: slot 2j holds X(2P), code 4, and slot 2j+1 is
// stale), whether the slot lies beyond the ragged last burst -- then the exact CSR lookup alone decides
// (PTX197:L33723-34006; the fingerprint sector is only a filter in front of it).  Reached for a
// fingerprint match (~2^-14 per table entry), an over-full bucket, or when the burst contained any
// of the rare cases above.
static __device__ __noinline__ void drain_slow(const Tables& tb, const uint32_t* s_win, int W, uint32_t t, int ci_first, int slot, int fill, uint64_t key) {
    if (slot >= fill) return;
    const int cj = ci_first - (slot >> 1), rj = cj / W, wj = cj - rj * W;
    const bool plus = slot & 1;
    const uint32_t idx = (uint32_t)rj * tb.NT + t;
    if (idx >= tb.N) return;
    const size_t NT2 = 2 * (size_t)tb.NT;
    const Fe qx = ldg_fe(tb.gx + (size_t)rj * NT2 + t, tb.gx + (size_t)rj * NT2 + tb.NT + t);
    const bool eq = fp::fe_eq(qx.v, s_win + wj * 32);
    if (eq && plus) return;
    const uint32_t h = (uint32_t)key & tb.mask, tag = (uint32_t)(key >> 32);
    if (probe_csr(tb.csr_off, tb.csr_tags, h, tag)) record_hit(tb, plus ? 1u : (eq ? 4u : 2u), idx, (uint32_t)wj);
}

// MODE 0: probe every candidate X in the baby-step table and record hits (the giant-step search).
// MODE 1: the same arithmetic used as a generator: the "giant" array holds (i+1)*G, the windows are
//         centres C_w*G, and the low 64 bits of X((C_w -+ (i+1))*G) are stored as baby-step keys --
//         the reference computes these with w sequential batch additions on the CPU
//         (baby(), 1_9_7File.pb:1162-1235).
template <int BLOCK, int MINB, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB)
k_giant_step(const __grid_constant__ Tables tb, const __grid_constant__ WinParams wp, int W) {
    __shared__ __align__(16) uint32_t s_win[kMaxW * 32];
    __shared__ uint64_t s_ring[MODE == 0 ? kRing * BLOCK : 1];
    const uint32_t t = blockIdx.x * BLOCK + threadIdx.x;
    const uint32_t NT = tb.NT;
    const int PP = (int)tb.PP;
    if (tb.trace && (threadIdx.x & 31) == 0) {
        uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        uint64_t now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        tb.trace[4 * (size_t)(t >> 5)] = smid; tb.trace[4 * (size_t)(t >> 5) + 1] = now;
    }
    for (int i = threadIdx.x; i < W * 32; i += BLOCK) s_win[i] = reinterpret_cast<const uint32_t*>(&wp)[i];
    __syncthreads();
    uint64_t* ring = s_ring + threadIdx.x;

    // code 5: the base point itself (thread 0 of the grid; PTX197:L49-109)
    if (t == 0) {
#pragma unroll 1
        for (int w = 0; w < W; w++) {
            if (MODE == 0) {
                uint32_t h = s_win[w * 32] & tb.mask, tag = s_win[w * 32 + 1];
                if (probe_csr(tb.csr_off, tb.csr_tags, h, tag)) record_hit(tb, 5u, 0u, (uint32_t)w);
            } else if (wp.center[w] <= tb.nkeys) {
                tb.keys_out[wp.center[w] - 1] = ((uint64_t)s_win[w * 32 + 1] << 32) | s_win[w * 32];
            }
        }
    }

    // ---------------- PASS 1: running product of the denominators -> chain ----------------
    // All addresses are running pointers (one 64-bit add per use): recomputing them from (r, w, ci)
    // cost ~60 integer instructions per point-window, and every instruction beside an IMAD.WIDE
    // costs issue time on this part.
    const size_t NT2 = 2 * (size_t)NT;
    Fe acc; acc.v[0] = 1;
#pragma unroll
    for (int i = 1; i < 8; i++) acc.v[i] = 0;
    {
        const uint4* gxp = tb.gx + t;
        uint4* cp = tb.chain + NT2 + t;
        if (!(GS_ABLATE & 2)) st8_stream(cp, cp + NT, acc.v);        // chain element -1 = 1: PASS 2 needs no special case for its last step
        cp += NT2;
#pragma unroll 1
        for (int r = 0; r < PP; r++, gxp += NT2) {
            const Fe qx = ldg_fe((GS_ABLATE & 4) ? tb.gx + t : gxp, (GS_ABLATE & 4) ? tb.gx + t + NT : gxp + NT);
            const uint32_t* sw = s_win;
#pragma unroll 1
            for (int w = 0; w < W; w++, sw += 32, cp += NT2) {
                Fe d;
                { Fe px = lds_fe(sw); fp::fe_sub_canon(d.v, px.v, qx.v); }
                if (fp::fe_is_zero_rare(d.v)) d = lds_fe(sw + 24);                          // 2*Py
                Fe n; fp::fe_mul(n.v, acc.v, d.v); acc = n;
                if (!(GS_ABLATE & 2)) st8_stream(cp, cp + NT, acc.v);
            }
        }
    }

    if (tb.trace && (threadIdx.x & 31) == 0) { uint64_t now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); tb.trace[4 * (size_t)(t >> 5) + 2] = now; }
    // ---------------- one inversion per thread ----------------
    Fe inv; fp::fe_inv(inv.v, acc.v);                                // (a by-value wrapper keeps inv out of local memory but costs 3 spills at 80 registers: measured slower)

    // ---------------- PASS 2: unwind, slopes, X, keys -> ring -> probe bursts ----------------
    // Rows outside, windows inside; loop-carried state is kept small (80 registers, 3 CTAs per SM): the giant
    // index of the row, one pointer into the chain, the chain index and the ring fill.  The rare cases -- a
    // padding row (giant index >= N), a tangent (Px == Qx) or the ragged last burst -- only raise `special`;
    // the drain then classifies every slot from its chain index (drain_slow) instead of the loop tracking
    // per-slot masks.  The chain carries two extra leading rows: element -1 = 1 (so that the last step loads
    // its "previous product" like every other) and one pad row (so that the L2 prefetch needs no guard).
    int fill = 0;
    bool special = false;
    int ci = PP * W - 1;
    uint32_t kidx = (uint32_t)(PP - 1) * NT + t;                    // giant index of the current row
    const uint4* cp = tb.chain + (size_t)(PP * W + 1) * NT2 + t;  // &element[ci]; element[ci-1] is one stride below
    const uint32_t* const sw_top = s_win + (W - 1) * 32;

    // drain: slot 2j / 2j+1 = minus / plus key of chain index ci_first - j
    auto drain = [&](int ci_first) {
#pragma unroll 1
        for (int s0 = 0; s0 < fill; s0 += kGroup) {
            FpSector sec[kGroup]; uint64_t key[kGroup];
#pragma unroll
            for (int j = 0; j < kGroup; j++) {
                key[j] = ring[(s0 + j) * BLOCK];
                sec[j].load(tb, (uint32_t)key[j] & tb.mask);
            }
            uint32_t need = special ? (1u << kGroup) - 1u : 0u;
#pragma unroll
            for (int j = 0; j < kGroup; j++) need |= (sec[j].maybe((uint32_t)(key[j] >> 32)) ? 1u : 0u) << j;
            // One warp-wide vote per group; the rare branch ends in an explicit __syncwarp(): without
            // it lanes that took the out-of-line path were observed (ncu: 31.93 active lanes in the
            // arithmetic, one SM busy 3x longer than the rest) to stay split from their warp for the
            // remainder of the loop.
            if (__any_sync(0xFFFFFFFFu, need != 0u)) {
#pragma unroll 1
                for (int j = 0; j < kGroup; j++)
                    if ((need >> j) & 1u) drain_slow(tb, s_win, W, t, ci_first, s0 + j, fill, ring[(s0 + j) * BLOCK]);    // (key re-read: no indexed register array)
                __syncwarp();
            }
        }
        fill = 0; special = false;
    };

#pragma unroll 1
    for (int r = PP - 1; r >= 0; r--, kidx -= NT) {
        // the row's giant point, held across its windows: slot (r, t) lives at (2r)*NT + t = 2*kidx - t
        const size_t go = (GS_ABLATE & 4) ? (size_t)t : 2 * (size_t)kidx - t;
        const Fe qx = ldg_fe(tb.gx + go, tb.gx + go + NT), qy = ldg_fe(tb.gy + go, tb.gy + go + NT);
        if (MODE == 0 && kidx >= tb.N) special = true;
#pragma unroll 1
        for (const uint32_t* sw = sw_top; sw >= s_win; sw -= 32, ci--) {
            Fe cprev;                                              // element[ci-1]
            cp -= NT2;
            if (!(GS_ABLATE & 2)) {
                ld8_stream(cprev.v, cp, cp + NT);
                if (GS_PREFETCH) { prefetch_l2(cp - NT2); prefetch_l2(cp - NT); }
            } else cprev = inv;
            Fe d;
            { Fe px = lds_fe(sw); fp::fe_sub_canon(d.v, px.v, qx.v); }
            const bool eq = fp::fe_is_zero_rare(d.v);
            if (eq) d = lds_fe(sw + 24);
            Fe invd;
            fp::fe_mul(invd.v, inv.v, cprev.v);                    // 1 / d_ci
            { Fe x; fp::fe_mul(x.v, inv.v, d.v); inv = x; }        // 1 / (d_0 ... d_{ci-1})
            uint64_t* rp = ring + fill * BLOCK;
            if (!eq) {
                // X = lambda^2 + (p - Px) - Qx is never formed, lambda^2 not even reduced: the low 64 bits of X follow
                // from a few limbs of the 512-bit square and of the other two terms (fp::fe_key_from_wide); the part
                // common to both candidates:
                const uint64_t t_lo = *reinterpret_cast<const uint64_t*>(sw + 16) - (((uint64_t)qx.v[1] << 32) | qx.v[0]);
                const int64_t t7 = (int64_t)sw[23] - (int64_t)qx.v[7];
#pragma unroll
                for (int side = 0; side < 2; side++) {              // side 0: P - Q (code 2), slope (Py + Qy)/d ; side 1: P + Q (code 1), slope (Py - Qy)/d
                    Fe py = lds_fe(sw + 8), sum, lam;
                    if (side == 0) fp::fe_add_lazy(sum.v, py.v, qy.v);
                    else fp::fe_sub_canon(sum.v, py.v, qy.v);
                    fp::fe_mul(lam.v, sum.v, invd.v);
                    const uint64_t key = candidate_key(lam, t_lo, t7, sw + 16, qx, tb.exact_thr);
                    if (MODE == 0) rp[side * BLOCK] = key;
                    else if (kidx < tb.N) {                         // multiplier C - (i+1) (side 0) or C + (i+1) (side 1)
                        const uint64_t i1 = (uint64_t)kidx + 1, c = wp.center[(sw - s_win) >> 5];
                        const uint64_t m = side == 0 ? c - i1 : c + i1;
                        if (m <= tb.nkeys) tb.keys_out[m - 1] = key;
                    }
                }
            } else {
                Fe px = lds_fe(sw);
                const uint64_t key = double_x_low64_v(px, invd);    // X(2P)  (code 4); cannot occur in emit mode
                if (MODE == 0) rp[0] = key;
                special = true;
            }
            if (MODE == 0) {
                fill += 2;
                if (fill == kRing) {
                    if (GS_ABLATE & 1) { fill = 0; special = false; }
                    else drain(ci + (kRing >> 1) - 1);
                }
            }
        }
    }
    if (MODE == 0 && fill && !(GS_ABLATE & 1)) {                    // the ragged last burst may end inside a group of four: every slot classified exactly
        special = true;
        drain((fill >> 1) - 1);
    }
    if (tb.trace && (threadIdx.x & 31) == 0) { uint64_t now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); tb.trace[4 * (size_t)(t >> 5) + 3] = now; }
}

// probe-only kernels on the fingerprint layout (measurement + parity of the table layout)
template <int U>
__global__ void __launch_bounds__(256) k_probe_bench_fp(const Tables tb, uint64_t nkeys, uint64_t seed, unsigned long long* nfound) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long found = 0;
    for (; i < nkeys; i += U * stride) {                            // U independent sector reads in flight per thread
        FpSector q[U]; uint64_t k[U]; bool v[U];
#pragma unroll
        for (int j = 0; j < U; j++) { v[j] = i + j * stride < nkeys; k[j] = splitmix64(seed + i + j * stride); if (v[j]) q[j].load(tb, (uint32_t)k[j] & tb.mask); }
#pragma unroll
        for (int j = 0; j < U; j++)
            if (v[j] && q[j].maybe((uint32_t)(k[j] >> 32)) && probe_csr(tb.csr_off, tb.csr_tags, (uint32_t)k[j] & tb.mask, (uint32_t)(k[j] >> 32))) found++;
    }
    if (found) atomicAdd(nfound, found);
}
__global__ void k_probe_keys_fp(const Tables tb, const uint64_t* keys, uint64_t nkeys, uint8_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nkeys) return;
    uint64_t k = keys[i];
    FpSector q; q.load(tb, (uint32_t)k & tb.mask);
    out[i] = q.maybe((uint32_t)(k >> 32)) && probe_csr(tb.csr_off, tb.csr_tags, (uint32_t)k & tb.mask, (uint32_t)(k >> 32)) ? 1 : 0;
}


// ------------------------------------------------------------------------------------------------
// layout conversion kernels (run once per upload)
// ------------------------------------------------------------------------------------------------

// G2 file image (u32 index of word i of giant k = r*8*TB + g + i*TB, word 0 most
// significant; X block then Y block) -> 128-bit SoA layout of the kernel, slot k = r'*NT + t.
__global__ void k_convert_g2(const uint32_t* __restrict__ g2, uint4* __restrict__ gx, uint4* __restrict__ gy,
                             uint32_t T, uint32_t B, uint32_t p, uint32_t N, uint32_t NT, uint32_t PP) {
    size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)NT * PP;
    if (slot >= total) return;
    uint32_t k = slot < N ? (uint32_t)slot : 0u;           // padding slots replicate giant 0 (never reported)
    uint64_t TB = (uint64_t)T * B;
    uint32_t blk = k / (T * p), thr = (k % (T * p)) / p, r = k % p;
    uint64_t g = (uint64_t)blk * T + thr;
    const uint32_t* X = g2 + (size_t)r * 8 * TB + g;
    const uint32_t* Y = X + (size_t)N * 8;
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[7 - i] = X[(size_t)i * TB]; y[7 - i] = Y[(size_t)i * TB]; }
    uint32_t rr = (uint32_t)(slot / NT), tt = (uint32_t)(slot % NT);
    gx[(size_t)(2 * rr) * NT + tt] = make_uint4(x[0], x[1], x[2], x[3]);
    gx[(size_t)(2 * rr + 1) * NT + tt] = make_uint4(x[4], x[5], x[6], x[7]);
    gy[(size_t)(2 * rr) * NT + tt] = make_uint4(y[0], y[1], y[2], y[3]);
    gy[(size_t)(2 * rr + 1) * NT + tt] = make_uint4(y[4], y[5], y[6], y[7]);
}

// diagnostics: element-wise field operations on arrays (parity tests of fp.cuh on the device)
__global__ void k_debug_field(int op, const uint32_t* a, const uint32_t* b, uint32_t* r, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x[8], y[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { x[j] = a[8 * i + j]; y[j] = b[8 * i + j]; z[j] = 0; }
    switch (op) {
        case 0: fp::fe_mul(z, x, y); break;
        case 1: fp::fe_sqr(z, x); break;
        case 2: fp::fe_inv(z, x); break;
        case 3: fp::fe_sub_canon(z, x, y); break;
        case 4: fp::fe_add_lazy(z, x, y); break;
        case 5: fp::fe_copy(z, x); fp::fe_canon(z); break;
        case 6: { uint64_t k = fp::fe_add_low64_canon(x, y); z[0] = (uint32_t)k; z[1] = (uint32_t)(k >> 32); break; }
        case 7: { uint64_t k = double_x_low64(x, y); z[0] = (uint32_t)k; z[1] = (uint32_t)(k >> 32); break; }
        case 8: {                                          // reduce512 of the 512-bit value (y << 256) | x, incl. its out-of-line rare tail
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 8; j++) { v[j] = x[j]; v[8 + j] = y[j]; }
            fp::reduce512(z, v);
            break;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) r[8 * i + j] = z[j];
}

// diagnostics: the hot loop's key extraction on arbitrary inputs, formed exactly as PASS 2 forms them.
// mode 0: in[i] = lambda (8 limbs): candidate_key itself (square, classification, exact path via lambda).
// mode 1: in[i] = an arbitrary 512-bit v (16 limbs) in place of lambda^2: the same classification and fast
//         path (fp::fe_key_from_wide); the exact path reduces v and runs the same exact extraction.
__global__ void k_debug_key(int mode, const uint32_t* in, const uint32_t* A, const uint32_t* Q, uint64_t* keys, uint8_t* exact, uint32_t n, uint32_t thr) {
    __shared__ __align__(16) uint32_t s_a[128 * 8];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fe qx;
    uint32_t* a = s_a + threadIdx.x * 8;
#pragma unroll
    for (int j = 0; j < 8; j++) { qx.v[j] = Q[8 * i + j]; a[j] = A[8 * i + j]; }
    const uint64_t t_lo = *reinterpret_cast<const uint64_t*>(a) - (((uint64_t)qx.v[1] << 32) | qx.v[0]);
    const int64_t t7 = (int64_t)a[7] - (int64_t)qx.v[7];
    uint32_t v[16];
    Fe lam;
    if (mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; j++) lam.v[j] = in[8 * i + j];
        fp::fe_sqr_wide(v, lam.v);
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = in[16 * i + j];
    }
    bool ex;
    uint64_t key = fp::fe_key_from_wide(v, t_lo, t7, thr, ex);
    if (mode == 0) key = candidate_key(lam, t_lo, t7, a, qx, thr);
    else if (ex) {
        Fe l2; fp::reduce512(l2.v, v);
        uint32_t u[8]; fp::fe_sub_canon(u, a, qx.v);
        key = fp::fe_add_low64_canon(l2.v, u);
    }
    keys[i] = key;
    exact[i] = ex ? 1 : 0;
}

}  // namespace gs
