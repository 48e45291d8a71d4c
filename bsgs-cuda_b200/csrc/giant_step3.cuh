// giant_step3.cuh -- giant-step kernel, third structure (see DESIGN.md "kernel history").
//
// What ncu said about the first kernel (profiles/r1_ncu_v1_w4.txt): 35 % of stall samples were
// "no instruction" (7 fully inlined 256-bit multiplications per window = 118 KB of code), 19 % sat
// on the CSR fall-back of the probe (a 32-byte bucket holds 7 x 32-bit tags; at mean load 4 one
// bucket in 20 overflows, i.e. 4 warps in 5 took the two-level dependent lookup), and only 4 warps
// per scheduler were resident (126 registers).  This version keeps the arithmetic inlined (no call
// marshalling) but makes the hot loop small and the probe exact-but-cheap:
//   * one flat loop over (giant row r, window w); window constants live in shared memory, so the
//     number of windows per launch is a run-time value (up to 16) and the code exists once;
//   * the two unwind products (inv*chain, inv*d) are two trips through ONE inlined multiplication,
//     the two sides (P-Q, P+Q) two trips through ONE inlined multiplication + squaring;
//   * probe keys are parked in a per-thread ring in shared memory and resolved in bursts of
//     independent sector loads, so no probe state is live across the arithmetic;
//   * the bucket sector holds fifteen 16-bit fingerprints (top half of the tag) + a count: a probe
//     is one 32-byte read; only a fingerprint match (2^-16 per entry) or an over-full bucket
//     (> 15 entries, ~1e-5 at load 4) goes to the exact CSR lookup, so the emitted records are
//     exactly those of the reference probe (PTX197:L33723-34006).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "fp.cuh"
#include "giant_step.cuh"

namespace gs3 {

using gs::Tables;
using gs::WinConst;
constexpr int kMaxW = 16;
constexpr int kRing = 16;                  // keys per thread between drains (2 per point-window)

struct Fe { uint32_t v[8]; };
struct WinParams { WinConst win[kMaxW]; };     // 2 KB of kernel parameters; rows: px | py | p - px | 2*py

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
// 16 halfwords per bucket: fp[0..14] = tag >> 16 of the first 15 entries (0 when unused), last
// halfword = entry count, 0xFFFF when the bucket holds more than 15.
__global__ void k_build_fp_sectors(const uint32_t* __restrict__ off, const uint32_t* __restrict__ tags,
                                   uint4* __restrict__ sectors, uint64_t H, unsigned long long* overfull) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    uint32_t lo = off[h], n = off[h + 1] - lo;
    uint32_t wbuf[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t a = (uint32_t)(2 * i) < n ? tags[lo + 2 * i] >> 16 : 0u;
        uint32_t b = (uint32_t)(2 * i + 1) < n && i < 7 ? tags[lo + 2 * i + 1] >> 16 : 0u;
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
        a = __ldg(p); b = __ldg(p + 1);
    }
    // true when the exact lookup is required: some halfword equals the fingerprint (this includes
    // chance matches with the count field or an empty slot -- harmless) or the bucket is over-full
    __device__ __forceinline__ bool maybe(uint32_t tag) const {
        const uint32_t f = (tag >> 16) * 0x00010001u;
        uint32_t acc = 0;
#define FP_Z(x) { uint32_t y = (x) ^ f; acc |= (y - 0x00010001u) & ~y; }
        FP_Z(a.x) FP_Z(a.y) FP_Z(a.z) FP_Z(a.w) FP_Z(b.x) FP_Z(b.y) FP_Z(b.z) FP_Z(b.w)
#undef FP_Z
        return ((acc & 0x80008000u) != 0u) | (b.w >= 0xFFFF0000u);
    }
};

// exact, slow: Px == Qx (true doubling -> code 4), PTX197:L28996-
__device__ __noinline__ uint64_t double_x_low64_v(Fe px, Fe inv2y) { return gs::double_x_low64(px.v, inv2y.v); }

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_giant_step3(const Tables tb, const __grid_constant__ WinParams wp, int W) {
    __shared__ __align__(16) uint32_t s_win[kMaxW * 32];
    __shared__ uint64_t s_ring[kRing * BLOCK];
    const uint32_t t = blockIdx.x * BLOCK + threadIdx.x;
    const uint32_t NT = tb.NT;
    const int PP = (int)tb.PP;
    for (int i = threadIdx.x; i < W * 32; i += BLOCK) s_win[i] = reinterpret_cast<const uint32_t*>(&wp)[i];
    __syncthreads();
    uint64_t* ring = s_ring + threadIdx.x;

    // code 5: the base point itself (thread 0 of the grid; PTX197:L49-109)
    if (t == 0) {
#pragma unroll 1
        for (int w = 0; w < W; w++) {
            uint32_t h = s_win[w * 32] & tb.mask, tag = s_win[w * 32 + 1];
            if (gs::probe_csr(tb.csr_off, tb.csr_tags, h, tag)) gs::record_hit(tb, 5u, 0u, (uint32_t)w);
        }
    }

    // ---------------- PASS 1: running product of the denominators -> chain ----------------
    Fe acc; acc.v[0] = 1;
#pragma unroll
    for (int i = 1; i < 8; i++) acc.v[i] = 0;
    {
        int r = 0, w = 0;
        const int total = PP * W;
#pragma unroll 1
        for (int ci = 0; ci < total; ci++) {
            Fe qx = ldg_fe(tb.gx + (size_t)(2 * r) * NT + t, tb.gx + (size_t)(2 * r + 1) * NT + t);
            Fe d;
            { Fe px = lds_fe(s_win + w * 32); fp::fe_sub_canon(d.v, px.v, qx.v); }
            if (fp::fe_is_zero(d.v)) d = lds_fe(s_win + w * 32 + 24);          // 2*Py
            Fe n; fp::fe_mul(n.v, acc.v, d.v); acc = n;
            size_t co = (size_t)ci * 2 * NT + t;
            gs::st8_stream(tb.chain + co, tb.chain + co + NT, acc.v);
            if (++w == W) { w = 0; r++; }
        }
    }

    // ---------------- one inversion per thread ----------------
    Fe inv; fp::fe_inv(inv.v, acc.v);

    // ---------------- PASS 2: unwind, slopes, X, keys -> ring -> probe bursts ----------------
    int fill = 0, ci_first = 0; uint32_t eqmask = 0, deadmask = 0;
    int r = PP - 1, w = W - 1;
#pragma unroll 1
    for (int ci = PP * W - 1; ci >= 0; ci--) {
        const bool live = (uint32_t)r * NT + t < tb.N;
        const uint4* gxp = tb.gx + (size_t)(2 * r) * NT + t;
        const uint32_t* sw = s_win + w * 32;
        Fe opnd;                                                   // chain[ci-1], or 1 for the very first element
        if (ci > 0) {
            size_t cp = (size_t)(ci - 1) * 2 * NT + t;
            gs::ld8_stream(opnd.v, tb.chain + cp, tb.chain + cp + NT);
            if (ci > 1) { prefetch_l2(tb.chain + cp - 2 * (size_t)NT); prefetch_l2(tb.chain + cp - (size_t)NT); }
        } else {
            opnd.v[0] = 1;
#pragma unroll
            for (int i = 1; i < 8; i++) opnd.v[i] = 0;
        }
        Fe d;
        { Fe qx = ldg_fe(gxp, gxp + NT); Fe px = lds_fe(sw); fp::fe_sub_canon(d.v, px.v, qx.v); }
        const bool eq = fp::fe_is_zero(d.v);
        if (eq) d = lds_fe(sw + 24);
        // invd = inv * chain[ci-1] ; inv = inv * d   -- two trips through one multiplication
        Fe invd;
#pragma unroll 1
        for (int j = 0; j < 2; j++) {
            Fe x; fp::fe_mul(x.v, inv.v, opnd.v);
            if (j == 0) { invd = x; opnd = d; } else inv = x;
        }
        if (!eq) {
            Fe u, s;
            { Fe qx = ldg_fe(gxp, gxp + NT); Fe npx = lds_fe(sw + 16); fp::fe_sub_canon(u.v, npx.v, qx.v); }   // -(Px + Qx)
            {
                const uint4* gyp = tb.gy + (size_t)(2 * r) * NT + t;
                s = ldg_fe(gyp, gyp + NT);                          // Qy: minus side uses Py + Qy (code 2)
            }
#pragma unroll 1
            for (int side = 0; side < 2; side++) {
                Fe py = lds_fe(sw + 8), sum, lam, l2;
                fp::fe_add_lazy(sum.v, py.v, s.v);
                fp::fe_mul(lam.v, sum.v, invd.v);
                fp::fe_sqr(l2.v, lam.v);
                ring[(fill + side) * BLOCK] = fp::fe_add_low64_canon(l2.v, u.v);
                // plus side uses Py - Qy = Py + (p - Qy) (code 1); Qy != 0 on this curve
                uint32_t z[8] = {FP_P0, FP_P1, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
                fp::fe_sub_canon(s.v, z, s.v);
            }
        } else {
            Fe px = lds_fe(sw);
            ring[fill * BLOCK] = double_x_low64_v(px, invd);        // X(2P)  (code 4)
        }
        if (fill == 0) ci_first = ci;
        eqmask |= (eq ? 1u : 0u) << (fill >> 1);
        deadmask |= (live ? 0u : 1u) << (fill >> 1);
        fill += 2;
        if (fill == kRing || ci == 0) {
            // drain: slot 2j / 2j+1 = minus / plus key of chain index ci_first - j
#pragma unroll 1
            for (int s0 = 0; s0 < fill; s0 += 4) {
                FpSector sec[4]; uint64_t key[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    key[j] = ring[(s0 + j) * BLOCK];
                    sec[j].load(tb, (uint32_t)key[j] & tb.mask);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int pair = (s0 + j) >> 1; const bool plus = (s0 + j) & 1;
                    const bool eqj = (eqmask >> pair) & 1u, dead = (deadmask >> pair) & 1u;
                    if (s0 + j >= fill || dead || (plus && eqj)) continue;
                    const uint32_t tag = (uint32_t)(key[j] >> 32);
                    if (sec[j].maybe(tag) && gs::probe_csr(tb.csr_off, tb.csr_tags, (uint32_t)key[j] & tb.mask, tag)) {
                        const int cj = ci_first - pair, rj = cj / W, wj = cj - rj * W;
                        gs::record_hit(tb, plus ? 1u : (eqj ? 4u : 2u), (uint32_t)rj * NT + t, (uint32_t)wj);
                    }
                }
            }
            fill = 0; eqmask = 0; deadmask = 0;
        }
        if (--w < 0) { w = W - 1; r--; }
    }
}

// probe-only kernels on the fingerprint layout (measurement + parity of the table layout)
__global__ void __launch_bounds__(256) k_probe_bench_fp(const Tables tb, uint64_t nkeys, uint64_t seed, unsigned long long* nfound) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long found = 0;
    for (; i < nkeys; i += 4 * stride) {                            // 4 independent sector reads in flight per thread
        FpSector q[4]; uint64_t k[4]; bool v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { v[j] = i + j * stride < nkeys; k[j] = gs::splitmix64(seed + i + j * stride); if (v[j]) q[j].load(tb, (uint32_t)k[j] & tb.mask); }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (v[j] && q[j].maybe((uint32_t)(k[j] >> 32)) && gs::probe_csr(tb.csr_off, tb.csr_tags, (uint32_t)k[j] & tb.mask, (uint32_t)(k[j] >> 32))) found++;
    }
    if (found) atomicAdd(nfound, found);
}
__global__ void k_probe_keys_fp(const Tables tb, const uint64_t* keys, uint64_t nkeys, uint8_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nkeys) return;
    uint64_t k = keys[i];
    FpSector q; q.load(tb, (uint32_t)k & tb.mask);
    out[i] = q.maybe((uint32_t)(k >> 32)) && gs::probe_csr(tb.csr_off, tb.csr_tags, (uint32_t)k & tb.mask, (uint32_t)(k >> 32)) ? 1 : 0;
}

}  // namespace gs3
