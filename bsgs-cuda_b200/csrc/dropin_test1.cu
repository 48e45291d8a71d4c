// dropin_test1.cu -- bit-level drop-in for the reference's embedded kernel ("Option A").
//
// The reference host loads a PTX text from its DataSection (1_9_7File.pb:5181-23979, de-obfuscated at
// 4739-4741) with cuModuleLoadData (2194), looks up the entry `_test1` (2198) and the constant block
// `_A` (2278-2283) and launches (B,1,1) x (T,1,1) with ONE u64 parameter, the base of its single
// device allocation (2260-2275, 2450).  This file is a from-scratch sm_100a kernel with exactly that
// interface and memory map, so the module built from it (dropin_test1.ptx / .cubin) can replace the
// DataSection bytes (scripts/make_datasection.py writes them) without touching the host:
//
//   _A (30 x u32, filled 1_9_7File.pb:2312-2325, 2442):  +4 w, +8 p, +12 N, +32..63 Px, +64..95 Py
//        (8 words each, most significant first), +96 puboffset (u64), +104 H+1, +112 H-1
//   base + 0            u32 hit counter;  base + 128 + 8i : record i = {u32 code, u32 idx}
//   base + 2048         X of the giants: word k (0 = most significant) of giant g*p + r (g = global
//                       thread id) at u32 index r*8*TB + k*TB + g, TB = T*B (Writeint, 1831-1903)
//   base + 2048 + 32N   Y, same order;   base + 2048 + 64N : N*32 bytes of scratch (the chain)
//   base + puboffset    htGPU image: u32 off[H+1], u32 tags[w] (packHTGPUFile, 3337-3444)
//
// Output semantics are the reference kernel's (PTX197:L49-109, L26955-36020): code 2 = X(P - Q_k) found,
// 1 = X(P + Q_k), 4 = X(2P) when Px == Qx_k, 5 = Px itself (thread 0; its idx word is left unwritten
// as in the reference), idx = k.  One deliberate difference: records beyond the 240 that fit below
// byte 2048 are counted but not stored (the reference overwrites its giant array, 1_9_7File.pb:2337).
//
// The arithmetic is this repository's (fp.cuh: one IMAD.WIDE per limb product, 255S+15M inversion);
// what the memory map dictates is kept: one window per launch, one inversion per thread per p points,
// the two-level CSR probe.  Probes are parked in a shared-memory ring and resolved four at a time so
// that the offset reads of different keys overlap.
#include <stdint.h>
#include "fp.cuh"

extern "C" {
__constant__ uint32_t _A[30];
}

namespace {

constexpr int kMaxT = 512;            // registers allow 512 threads per CTA (the reference's own -t 512 is its README setting)
constexpr int kRingKeys = 8;          // keys per thread between drains (4 giant points)

struct Fe { uint32_t v[8]; };

__device__ __forceinline__ Fe ld_words(const uint32_t* p, uint32_t TB) {          // 8 coalesced words, most significant first
    Fe r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[7 - k] = __ldg(p + (size_t)k * TB);
    return r;
}
// a random 4-byte table read: cached in L1 as usual (the neighbouring word is read next), but only a
// 64-byte fill from DRAM into L2 instead of the default 128-byte line
__device__ __forceinline__ uint32_t ld_tbl(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ Fe lds_fe(const uint32_t* s) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = s[i];
    return r;
}

__device__ __forceinline__ void record(uint32_t* base, uint32_t code, uint32_t idx, bool with_idx) {
    uint32_t pos = atomicAdd(base, 1u);
    if (pos < 240u) { base[32 + 2 * pos] = code; if (with_idx) base[33 + 2 * pos] = idx; }
}

// the reference probe (PTX197:L33723-34006): bucket = low bits of the low word, binary search of the
// ascending tag run
__device__ __noinline__ bool probe_csr(const uint32_t* off, const uint32_t* tags, uint32_t h, uint32_t tag) {
    uint32_t lo = off[h], hi = off[h + 1];
    while (lo < hi) {
        uint32_t c = lo + ((hi - lo) >> 1), t = tags[c];
        if (tag > t) lo = c + 1; else if (tag < t) hi = c; else return true;
    }
    return false;
}

// exact, slow: true doubling of P when Px == Qx (PTX197:L28996-)
__device__ __noinline__ uint64_t double_x_low64(Fe px, Fe inv2y) {
    uint32_t xx[8], t3[8], lam[8], X[8], t[8];
    fp::fe_sqr(xx, px.v); fp::fe_canon(xx);
    fp::fe_add_lazy(t, xx, xx); fp::fe_canon(t);
    fp::fe_add_lazy(t3, t, xx); fp::fe_canon(t3);
    fp::fe_mul(lam, t3, inv2y.v);
    fp::fe_sqr(X, lam); fp::fe_canon(X);
    fp::fe_sub_canon(t, X, px.v);
    fp::fe_sub_canon(X, t, px.v);
    return ((uint64_t)X[1] << 32) | X[0];
}

}  // namespace

extern "C" __global__ void __launch_bounds__(kMaxT, 1) _test1(uint32_t* base) {
    __shared__ uint32_t s_c[32];                     // Px | Py | p - Px | 2Py, little-endian limbs
    __shared__ uint64_t s_ring[kRingKeys * kMaxT];
    const uint32_t T = blockDim.x, TB = T * gridDim.x, g = blockIdx.x * T + threadIdx.x;
    const uint32_t p = _A[2], N = _A[3];
    const uint32_t mask = _A[28];
    const uint32_t* off = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(base) + *reinterpret_cast<const uint64_t*>(&_A[24]));
    const uint32_t* tags = off + _A[26];
    const uint32_t* X = base + 512;
    const uint32_t* Y = X + (size_t)8 * N;
    uint4* CH = reinterpret_cast<uint4*>(base + 512 + (size_t)16 * N);      // private scratch: two 128-bit halves per giant, [r][half][g]

    if (threadIdx.x == 0) {
        uint32_t px[8], py[8], z[8], r[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { px[i] = _A[15 - i]; py[i] = _A[23 - i]; z[i] = 0; }
        fp::fe_sub_canon(r, z, px);                  // p - Px
#pragma unroll
        for (int i = 0; i < 8; i++) { s_c[i] = px[i]; s_c[8 + i] = py[i]; s_c[16 + i] = r[i]; }
        fp::fe_add_lazy(r, py, py); fp::fe_canon(r);
#pragma unroll
        for (int i = 0; i < 8; i++) s_c[24 + i] = r[i];
    }
    __syncthreads();
    if (g == 0 && probe_csr(off, tags, s_c[0] & mask, s_c[1])) record(base, 5u, 0u, false);      // PTX197:L49-109
    if (T > kMaxT) return;                           // cannot be launched anyway (register file)

    // PASS 1 (PTX197:L149-1297): running product of the denominators; the chain block is private
    // scratch, so it is laid out as 128-bit vectors (two coalesced 512-byte warp transactions per element)
    Fe acc; acc.v[0] = 1;
#pragma unroll
    for (int i = 1; i < 8; i++) acc.v[i] = 0;
    const size_t rowstride = (size_t)8 * TB;
    {
        const uint32_t* xp = X + g;
        uint4* cp = CH + g;
#pragma unroll 1
        for (uint32_t r = 0; r < p; r++, xp += rowstride, cp += 2 * (size_t)TB) {
            Fe qx = ld_words(xp, TB), d;
            { Fe px = lds_fe(s_c); fp::fe_sub_canon(d.v, px.v, qx.v); }
            if (fp::fe_is_zero_rare(d.v)) d = lds_fe(s_c + 24);
            Fe n; fp::fe_mul(n.v, acc.v, d.v); acc = n;
            cp[0] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
            cp[TB] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
        }
    }
    Fe inv; fp::fe_inv(inv.v, acc.v);               // one inversion per thread (PTX197:L1319-26062 uses 504 multiplications)

    // PASS 2 (PTX197:L26955-36020)
    uint64_t* ring = s_ring + threadIdx.x;
    int fill = 0; uint32_t eqmask = 0;
    const uint32_t* xp = X + g + (size_t)(p - 1) * rowstride;
    const uint4* cp = CH + g + (size_t)(p - 1) * 2 * TB;
    const ptrdiff_t yoff = Y - X;
#pragma unroll 1
    for (int r = (int)p - 1; r >= 0; r--, xp -= rowstride) {
        Fe cprev;
        cp -= 2 * (size_t)TB;
        if (r > 0) {
            const uint4 c0 = __ldcs(cp), c1 = __ldcs(cp + TB);
            cprev.v[0] = c0.x; cprev.v[1] = c0.y; cprev.v[2] = c0.z; cprev.v[3] = c0.w;
            cprev.v[4] = c1.x; cprev.v[5] = c1.y; cprev.v[6] = c1.z; cprev.v[7] = c1.w;
        } else {
            cprev.v[0] = 1;
#pragma unroll
            for (int i = 1; i < 8; i++) cprev.v[i] = 0;
        }
        const Fe qx = ld_words(xp, TB), qy = ld_words(xp + yoff, TB);
        Fe d;
        { Fe px = lds_fe(s_c); fp::fe_sub_canon(d.v, px.v, qx.v); }
        const bool eq = fp::fe_is_zero_rare(d.v);
        if (eq) d = lds_fe(s_c + 24);
        Fe invd;
        fp::fe_mul(invd.v, inv.v, cprev.v);
        { Fe x; fp::fe_mul(x.v, inv.v, d.v); inv = x; }
        if (!eq) {
            Fe u;
            { Fe npx = lds_fe(s_c + 16); fp::fe_sub_canon(u.v, npx.v, qx.v); }
#pragma unroll
            for (int side = 0; side < 2; side++) {   // side 0: P - Q (code 2); side 1: P + Q (code 1)
                Fe py = lds_fe(s_c + 8), sum, lam, l2;
                if (side == 0) fp::fe_add_lazy(sum.v, py.v, qy.v); else fp::fe_sub_canon(sum.v, py.v, qy.v);
                fp::fe_mul(lam.v, sum.v, invd.v);
                fp::fe_sqr(l2.v, lam.v);
                ring[(fill + side) * kMaxT] = fp::fe_add_low64_canon(l2.v, u.v);
            }
        } else {
            ring[fill * kMaxT] = double_x_low64(lds_fe(s_c), invd);
        }
        eqmask |= (eq ? 1u : 0u) << (fill >> 1);
        fill += 2;
        if (fill == kRingKeys || r == 0) {
            // slot 2j / 2j+1 = minus / plus key of point r_first - j, r_first = r + fill/2 - 1.
            // Four offset pairs in flight, then the four searches.  (Measured on config 2: batching the
            // first tag of all eight buckets, L2 prefetch of the next rows, and no-allocate 64-byte-fill
            // loads of the offsets were all slower, 8.5-10.7 vs 11.4 G steps/s -- the reference layout
            // makes this kernel bound by L1 request throughput: 8 word loads per coordinate and ~5
            // scattered 4-byte reads per probe.)
            const int r_first = r + (fill >> 1) - 1;
#pragma unroll 1
            for (int s0 = 0; s0 < fill; s0 += 4) {
                uint64_t key[4]; uint32_t lo[4], hi[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    key[j] = ring[(s0 + j) * kMaxT];
                    const uint32_t h = (uint32_t)key[j] & mask;
                    lo[j] = ld_tbl(off + h); hi[j] = ld_tbl(off + h + 1);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int pair = (s0 + j) >> 1; const bool plus = (s0 + j) & 1;
                    const bool eqj = (eqmask >> pair) & 1u;
                    if (s0 + j >= fill || (plus && eqj)) continue;
                    const uint32_t tag = (uint32_t)(key[j] >> 32);
                    uint32_t a = lo[j], b = hi[j]; bool found = false;
                    while (a < b) {
                        uint32_t c = a + ((b - a) >> 1), t = ld_tbl(tags + c);
                        if (tag > t) a = c + 1; else if (tag < t) b = c; else { found = true; break; }
                    }
                    if (found) record(base, plus ? 1u : (eqj ? 4u : 2u), g * p + (uint32_t)(r_first - pair), true);
                }
            }
            __syncwarp();
            fill = 0; eqmask = 0;
        }
    }
}
