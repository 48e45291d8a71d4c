// giant_step.cuh -- the giant-step kernel for sm_100a and its table-layout helpers.
//
// One launch evaluates, for W consecutive window base points P_w and every giant point
// Q_k = (k+1)*(-2w*G), k < N, the candidates X(P_w - Q_k) and X(P_w + Q_k) and probes each in the
// baby-step table -- the work of W launches of the reference kernel `_test1`
// (1_9_7File.pb:5181-23979; PASS 1 PTX197:L149-1297, inversion L1319-26062, PASS 2 L26955-36020).
//
// B200-first differences from the reference kernel (results are identical, see tests):
//   * limb products are single IMAD.WIDE.U32 issues (fp.cuh) instead of mad.lo + mad.hi pairs;
//   * one Fermat inversion per thread covers PP*W points (reference: p points) and uses the
//     255S+15M addition chain (reference: 255 + 249 generic multiplications);
//   * giant coordinates are stored as 128-bit vectors, SoA across threads, so each warp load is one
//     fully coalesced 512-byte transaction; each Q is loaded once for all W windows;
//   * the CSR table (offset pair -> binary search, two dependent random reads) is re-laid out at
//     upload into one 32-byte (or 64-byte) sector per bucket, so a probe is ONE random sector read
//     with no dependent access; over-full buckets fall back to the CSR image kept beside it;
//   * probe sectors are requested one iteration ahead and consumed after the next point's
//     arithmetic, so the random HBM latency is hidden behind ~10^3 integer instructions.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "fp.cuh"

namespace gs {

constexpr int kMaxW = 8;
constexpr int kBlock = 256;

// per-window constants, all canonical, little-endian limbs
struct WinConst {
    uint32_t px[8], py[8];
    uint32_t npx[8];      // p - Px
    uint32_t py2[8];      // 2*Py mod p  (denominator of the tangent, used when Px == Qx)
};
template <int W> struct StepParams { WinConst win[W]; };

struct Tables {
    const uint4* gx;          // [PP*2][NT] : limbs 0-3 / 4-7 of Qx, slot (r, t) at (2r+h)*NT + t
    const uint4* gy;
    uint4* chain;             // [PP*W*2][NT]
    const uint32_t* buckets;  // H buckets of bucket_words u32: tags[bucket_words-1], meta
    const uint32_t* csr_off;  // H+1 offsets   } the htGPU image as uploaded; fallback for
    const uint32_t* csr_tags; // w tags        } over-full buckets
    uint32_t* hit_count;      // [0] = number of records
    uint32_t* hits;           // records of 3 u32: code, idx, window
    uint32_t hit_cap;
    uint32_t NT, PP, N;       // threads in the grid, points per thread, giants
    uint32_t mask;            // H - 1
};

__device__ __forceinline__ void ld8(uint32_t r[8], const uint4* p0, const uint4* p1) {
    uint4 a = __ldg(p0), b = __ldg(p1);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
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

// CSR fallback == the reference probe (PTX197:L33723-34006; FOUNDINSORTNEW PTX173:L1447-1510)
__device__ __noinline__ bool probe_csr(const uint32_t* off, const uint32_t* tags, uint32_t h, uint32_t tag) {
    uint32_t lo = off[h], hi = off[h + 1];
    while (lo < hi) {
        uint32_t c = (lo + hi) >> 1, t = tags[c];
        if (tag > t) lo = c + 1; else if (tag < t) hi = c; else return true;
    }
    return false;
}

// A probe whose sector has been requested but not yet examined.
template <int BW> struct Pending;
template <> struct Pending<8> {
    uint4 a, b; uint32_t h, tag;
    __device__ __forceinline__ void issue(const Tables& tb, uint64_t key) {
        h = (uint32_t)key & tb.mask; tag = (uint32_t)(key >> 32);
        const uint4* p = reinterpret_cast<const uint4*>(tb.buckets) + 2ull * h;
        a = __ldg(p); b = __ldg(p + 1);
    }
    __device__ __forceinline__ bool resolve(const Tables& tb) const {
        uint32_t meta = b.w;
        uint32_t m = (a.x == tag ? 1u : 0u) | (a.y == tag ? 2u : 0u) | (a.z == tag ? 4u : 0u) | (a.w == tag ? 8u : 0u) |
                     (b.x == tag ? 16u : 0u) | (b.y == tag ? 32u : 0u) | (b.z == tag ? 64u : 0u);
        if (meta <= 7u) return (m & ((1u << meta) - 1u)) != 0u;
        return probe_csr(tb.csr_off, tb.csr_tags, h, tag);
    }
};
template <> struct Pending<16> {
    uint4 a, b, c, d; uint32_t h, tag;
    __device__ __forceinline__ void issue(const Tables& tb, uint64_t key) {
        h = (uint32_t)key & tb.mask; tag = (uint32_t)(key >> 32);
        const uint4* p = reinterpret_cast<const uint4*>(tb.buckets) + 4ull * h;
        a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3);
    }
    __device__ __forceinline__ bool resolve(const Tables& tb) const {
        uint32_t meta = d.w;
        uint32_t m = (a.x == tag ? 1u : 0u) | (a.y == tag ? 2u : 0u) | (a.z == tag ? 4u : 0u) | (a.w == tag ? 8u : 0u) |
                     (b.x == tag ? 16u : 0u) | (b.y == tag ? 32u : 0u) | (b.z == tag ? 64u : 0u) | (b.w == tag ? 128u : 0u) |
                     (c.x == tag ? 256u : 0u) | (c.y == tag ? 512u : 0u) | (c.z == tag ? 1024u : 0u) | (c.w == tag ? 2048u : 0u) |
                     (d.x == tag ? 4096u : 0u) | (d.y == tag ? 8192u : 0u) | (d.z == tag ? 16384u : 0u);
        if (meta <= 15u) return (m & ((1u << meta) - 1u)) != 0u;
        return probe_csr(tb.csr_off, tb.csr_tags, h, tag);
    }
};

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

template <int W, int BW>
__global__ void __launch_bounds__(kBlock, 2)
k_giant_step(const Tables tb, const __grid_constant__ StepParams<W> sp) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t NT = tb.NT;
    const int PP = (int)tb.PP;

    // code 5: the base point itself (thread 0 of the grid; PTX197:L49-109)
    if (t == 0) {
#pragma unroll
        for (int w = 0; w < W; w++) {
            Pending<BW> pr;
            pr.issue(tb, ((uint64_t)sp.win[w].px[1] << 32) | sp.win[w].px[0]);
            if (pr.resolve(tb)) record_hit(tb, 5u, 0u, (uint32_t)w);
        }
    }

    // ---------------- PASS 1: running product of the denominators -> chain ----------------
    uint32_t acc[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    {
        uint32_t qx[8], nq[8];
        ld8(qx, tb.gx + t, tb.gx + NT + t);
#pragma unroll 1
        for (int r = 0; r < PP; r++) {
            if (r + 1 < PP) ld8(nq, tb.gx + (size_t)(2 * r + 2) * NT + t, tb.gx + (size_t)(2 * r + 3) * NT + t);
#pragma unroll
            for (int w = 0; w < W; w++) {
                uint32_t d[8];
                fp::fe_sub_canon(d, sp.win[w].px, qx);
                if (fp::fe_is_zero(d)) fp::fe_copy(d, sp.win[w].py2);
                uint32_t n[8];
                fp::fe_mul(n, acc, d);
                fp::fe_copy(acc, n);
                size_t ci = (size_t)(r * W + w) * 2 * NT + t;
                st8_stream(tb.chain + ci, tb.chain + ci + NT, acc);
            }
            fp::fe_copy(qx, nq);
        }
    }

    // ---------------- one inversion per thread ----------------
    uint32_t inv[8];
    fp::fe_inv(inv, acc);

    // ---------------- PASS 2: unwind, slopes, X, probe ----------------
    Pending<BW> pm, pp;            // probes of the previous point (minus / plus side)
    uint32_t pend_idx = 0, pend_win = 0;
    bool pend = false;

    uint32_t qx[8], qy[8];
    ld8(qx, tb.gx + (size_t)(2 * (PP - 1)) * NT + t, tb.gx + (size_t)(2 * (PP - 1) + 1) * NT + t);
    ld8(qy, tb.gy + (size_t)(2 * (PP - 1)) * NT + t, tb.gy + (size_t)(2 * (PP - 1) + 1) * NT + t);
#pragma unroll 1
    for (int r = PP - 1; r >= 0; r--) {
        uint32_t nqx[8], nqy[8];
        if (r > 0) {
            ld8(nqx, tb.gx + (size_t)(2 * r - 2) * NT + t, tb.gx + (size_t)(2 * r - 1) * NT + t);
            ld8(nqy, tb.gy + (size_t)(2 * r - 2) * NT + t, tb.gy + (size_t)(2 * r - 1) * NT + t);
        }
        const uint32_t k = (uint32_t)r * NT + t;
        const bool live = k < tb.N;
#pragma unroll
        for (int w = W - 1; w >= 0; w--) {
            const int ci = r * W + w;
            uint32_t invd[8];
            if (ci > 0) {
                uint32_t cprev[8];
                size_t cp = (size_t)(ci - 1) * 2 * NT + t;
                ld8_stream(cprev, tb.chain + cp, tb.chain + cp + NT);
                fp::fe_mul(invd, inv, cprev);
            } else {
                fp::fe_copy(invd, inv);
            }
            uint32_t d[8];
            fp::fe_sub_canon(d, sp.win[w].px, qx);
            const bool eq = fp::fe_is_zero(d);
            if (eq) fp::fe_copy(d, sp.win[w].py2);
            if (ci > 0) { uint32_t n[8]; fp::fe_mul(n, inv, d); fp::fe_copy(inv, n); }

            uint64_t keym = 0, keyp = 0;
            if (!eq) {
                uint32_t u[8], s[8], lam[8], l2[8];
                fp::fe_sub_canon(u, sp.win[w].npx, qx);              // -(Px + Qx)
                fp::fe_add_lazy(s, sp.win[w].py, qy);                // minus side: P - Q   (code 2)
                fp::fe_mul(lam, s, invd);
                fp::fe_sqr(l2, lam);
                keym = fp::fe_add_low64_canon(l2, u);
                fp::fe_sub_canon(s, sp.win[w].py, qy);               // plus side: P + Q    (code 1)
                fp::fe_mul(lam, s, invd);
                fp::fe_sqr(l2, lam);
                keyp = fp::fe_add_low64_canon(l2, u);
            } else {
                keym = double_x_low64(sp.win[w].px, invd);           // X(2P)              (code 4)
            }
            // examine the sectors requested one point ago, then request this point's
            if (pend) {
                if (pm.resolve(tb)) record_hit(tb, pend_idx >> 31 ? 4u : 2u, pend_idx & 0x7FFFFFFFu, pend_win);
                if (!(pend_idx >> 31) && pp.resolve(tb)) record_hit(tb, 1u, pend_idx, pend_win);
            }
            pend = live;
            if (live) {
                pm.issue(tb, keym);
                if (!eq) pp.issue(tb, keyp);
                pend_idx = k | (eq ? 0x80000000u : 0u);
                pend_win = (uint32_t)w;
            }
        }
        if (r > 0) { fp::fe_copy(qx, nqx); fp::fe_copy(qy, nqy); }
    }
    if (pend) {
        if (pm.resolve(tb)) record_hit(tb, pend_idx >> 31 ? 4u : 2u, pend_idx & 0x7FFFFFFFu, pend_win);
        if (!(pend_idx >> 31) && pp.resolve(tb)) record_hit(tb, 1u, pend_idx, pend_win);
    }
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

// htGPU CSR image -> one sector per bucket: (BW-1) tags + meta (count, or 0xFFFFFFFF = over-full)
template <int BW>
__global__ void k_build_buckets(const uint32_t* __restrict__ off, const uint32_t* __restrict__ tags,
                                uint32_t* __restrict__ buckets, uint64_t H, unsigned long long* overfull) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    uint32_t lo = off[h], n = off[h + 1] - lo;
    uint32_t wbuf[BW];
#pragma unroll
    for (int i = 0; i < BW - 1; i++) wbuf[i] = (uint32_t)i < n ? tags[lo + i] : 0u;
    wbuf[BW - 1] = n <= (uint32_t)(BW - 1) ? n : 0xFFFFFFFFu;
    if (n > (uint32_t)(BW - 1)) atomicAdd(overfull, 1ull);
    uint4* dst = reinterpret_cast<uint4*>(buckets) + h * (BW / 4);
#pragma unroll
    for (int i = 0; i < BW / 4; i++) dst[i] = make_uint4(wbuf[4 * i], wbuf[4 * i + 1], wbuf[4 * i + 2], wbuf[4 * i + 3]);
}

// probe-only kernels (measurement + parity of the table layout)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
template <int BW>
__global__ void __launch_bounds__(256) k_probe_bench(const Tables tb, uint64_t nkeys, uint64_t seed, unsigned long long* nfound) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long found = 0;
    // 4 independent probes in flight per thread
    for (; i < nkeys; i += 4 * stride) {
        Pending<BW> q[4]; bool v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { v[j] = i + j * stride < nkeys; if (v[j]) q[j].issue(tb, splitmix64(seed + i + j * stride)); }
#pragma unroll
        for (int j = 0; j < 4; j++) if (v[j] && q[j].resolve(tb)) found++;
    }
    if (found) atomicAdd(nfound, found);
}
template <int BW>
__global__ void k_probe_keys(const Tables tb, const uint64_t* keys, uint64_t nkeys, uint8_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nkeys) return;
    Pending<BW> q; q.issue(tb, keys[i]);
    out[i] = q.resolve(tb) ? 1 : 0;
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
    }
#pragma unroll
    for (int j = 0; j < 8; j++) r[8 * i + j] = z[j];
}

}  // namespace gs

// ================================================================================================
// v2: compact-code kernel.  ncu on v1 (profiles/r1a) showed 36 % of all stall cycles in
// "no instruction" (the fully inlined body of 7 field multiplications per window overflows the
// instruction caches) and only 4 warps per scheduler at 126 registers.  Here the 256-bit
// multiplication and squaring are real functions with register-passed operands, the window loop is
// a run-time loop (window constants in shared memory), and probe keys go through a per-thread ring
// in shared memory that is drained in bursts of independent sector loads, so that no probe state
// is live across the arithmetic.
// ================================================================================================
namespace gs2 {

using gs::Tables;
using gs::WinConst;
constexpr int kMaxW = 16;
constexpr int kBlock = 256;
constexpr int kRing = 16;                  // keys per thread between drains (2 per point-window)

struct Fe { uint32_t v[8]; };
struct WinParams { WinConst win[kMaxW]; };     // 2 KB of kernel parameters; rows: px | py | p - px | 2*py

__device__ __noinline__ Fe mulf(Fe a, Fe b) { Fe r; fp::fe_mul(r.v, a.v, b.v); return r; }
__device__ __noinline__ Fe sqrf(Fe a) { Fe r; fp::fe_sqr(r.v, a.v); return r; }

// a^(p-2), 255 S + 15 M, through the two functions above (keeps the kernel's code small)
__device__ __noinline__ Fe sqrnf(Fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) a = sqrf(a);
    return a;
}
__device__ __noinline__ Fe invf(Fe a) {
    Fe x2 = mulf(sqrnf(a, 1), a);
    Fe x3 = mulf(sqrnf(x2, 1), a);
    Fe x6 = mulf(sqrnf(x3, 3), x3);
    Fe x9 = mulf(sqrnf(x6, 3), x3);
    Fe x11 = mulf(sqrnf(x9, 2), x2);
    Fe x22 = mulf(sqrnf(x11, 11), x11);
    Fe x44 = mulf(sqrnf(x22, 22), x22);
    Fe x88 = mulf(sqrnf(x44, 44), x44);
    Fe x176 = mulf(sqrnf(x88, 88), x88);
    Fe x220 = mulf(sqrnf(x176, 44), x44);
    Fe x223 = mulf(sqrnf(x220, 3), x3);
    Fe t = mulf(sqrnf(x223, 23), x22);
    t = mulf(sqrnf(t, 5), a);
    t = mulf(sqrnf(t, 3), x2);
    return mulf(sqrnf(t, 2), a);
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

template <int BW> struct Sector;
template <> struct Sector<8> {
    uint4 a, b;
    __device__ __forceinline__ void load(const Tables& tb, uint32_t h) {
        const uint4* p = reinterpret_cast<const uint4*>(tb.buckets) + 2ull * h;
        a = __ldg(p); b = __ldg(p + 1);
    }
    __device__ __forceinline__ bool match(const Tables& tb, uint32_t h, uint32_t tag) const {
        uint32_t meta = b.w;
        uint32_t m = (a.x == tag ? 1u : 0u) | (a.y == tag ? 2u : 0u) | (a.z == tag ? 4u : 0u) | (a.w == tag ? 8u : 0u) |
                     (b.x == tag ? 16u : 0u) | (b.y == tag ? 32u : 0u) | (b.z == tag ? 64u : 0u);
        if (meta <= 7u) return (m & ((1u << meta) - 1u)) != 0u;
        return gs::probe_csr(tb.csr_off, tb.csr_tags, h, tag);
    }
};
template <> struct Sector<16> {
    uint4 a, b, c, d;
    __device__ __forceinline__ void load(const Tables& tb, uint32_t h) {
        const uint4* p = reinterpret_cast<const uint4*>(tb.buckets) + 4ull * h;
        a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3);
    }
    __device__ __forceinline__ bool match(const Tables& tb, uint32_t h, uint32_t tag) const {
        uint32_t meta = d.w;
        uint32_t m = (a.x == tag ? 1u : 0u) | (a.y == tag ? 2u : 0u) | (a.z == tag ? 4u : 0u) | (a.w == tag ? 8u : 0u) |
                     (b.x == tag ? 16u : 0u) | (b.y == tag ? 32u : 0u) | (b.z == tag ? 64u : 0u) | (b.w == tag ? 128u : 0u) |
                     (c.x == tag ? 256u : 0u) | (c.y == tag ? 512u : 0u) | (c.z == tag ? 1024u : 0u) | (c.w == tag ? 2048u : 0u) |
                     (d.x == tag ? 4096u : 0u) | (d.y == tag ? 8192u : 0u) | (d.z == tag ? 16384u : 0u);
        if (meta <= 15u) return (m & ((1u << meta) - 1u)) != 0u;
        return gs::probe_csr(tb.csr_off, tb.csr_tags, h, tag);
    }
};

// Drain `cnt` (<= kRing, even) keys of this thread's ring.  Slot 2j / 2j+1 hold the minus / plus
// key of the point-window with chain index ci0 - j (the unwind runs downwards); bit j of eqmask
// marks a doubling (code 4, plus slot unused), bit j of deadmask a padding point.
template <int BW>
__device__ __noinline__ void drain(const Tables& tb, const uint64_t* ring, int cnt, int ci0, int W, uint32_t t,
                                   uint32_t eqmask, uint32_t deadmask) {
#pragma unroll 1
    for (int s0 = 0; s0 < cnt; s0 += 4) {
        Sector<BW> sec[4]; uint64_t key[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (s0 + j >= cnt) continue;
            key[j] = ring[(s0 + j) * kBlock];
            sec[j].load(tb, (uint32_t)key[j] & tb.mask);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int pair = (s0 + j) >> 1; bool plus = (s0 + j) & 1;
            bool eq = (eqmask >> pair) & 1u, dead = (deadmask >> pair) & 1u;
            if (s0 + j >= cnt || dead || (plus && eq)) continue;
            if (sec[j].match(tb, (uint32_t)key[j] & tb.mask, (uint32_t)(key[j] >> 32))) {
                int ci = ci0 - pair, r = ci / W, w = ci - r * W;
                gs::record_hit(tb, plus ? 1u : (eq ? 4u : 2u), (uint32_t)r * tb.NT + t, (uint32_t)w);
            }
        }
    }
}

template <int BW, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
k_giant_step2(const Tables tb, const __grid_constant__ WinParams wp, int W) {
    __shared__ __align__(16) uint32_t s_win[kMaxW * 32];
    __shared__ uint64_t s_ring[kRing * kBlock];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t NT = tb.NT;
    const int PP = (int)tb.PP;
    for (int i = threadIdx.x; i < W * 32; i += kBlock) s_win[i] = reinterpret_cast<const uint32_t*>(&wp)[i];
    __syncthreads();
    uint64_t* ring = s_ring + threadIdx.x;

    // code 5: the base point itself (thread 0 of the grid; PTX197:L49-109)
    if (t == 0) {
#pragma unroll 1
        for (int w = 0; w < W; w++) {
            uint64_t key = ((uint64_t)s_win[w * 32 + 1] << 32) | s_win[w * 32];
            Sector<BW> sec; sec.load(tb, (uint32_t)key & tb.mask);
            if (sec.match(tb, (uint32_t)key & tb.mask, (uint32_t)(key >> 32))) gs::record_hit(tb, 5u, 0u, (uint32_t)w);
        }
    }

    // ---------------- PASS 1: running product of the denominators -> chain ----------------
    Fe acc; acc.v[0] = 1;
#pragma unroll
    for (int i = 1; i < 8; i++) acc.v[i] = 0;
#pragma unroll 1
    for (int r = 0; r < PP; r++) {
        Fe qx = ldg_fe(tb.gx + (size_t)(2 * r) * NT + t, tb.gx + (size_t)(2 * r + 1) * NT + t);
#pragma unroll 1
        for (int w = 0; w < W; w++) {
            Fe px = lds_fe(s_win + w * 32), d;
            fp::fe_sub_canon(d.v, px.v, qx.v);
            if (fp::fe_is_zero(d.v)) d = lds_fe(s_win + w * 32 + 24);          // 2*Py
            acc = mulf(acc, d);
            size_t ci = (size_t)(r * W + w) * 2 * NT + t;
            gs::st8_stream(tb.chain + ci, tb.chain + ci + NT, acc.v);
        }
    }

    // ---------------- one inversion per thread ----------------
    Fe inv = invf(acc);

    // ---------------- PASS 2: unwind, slopes, X, keys -> ring -> probe bursts ----------------
    int fill = 0, ci_first = 0; uint32_t eqmask = 0, deadmask = 0;
#pragma unroll 1
    for (int r = PP - 1; r >= 0; r--) {
        const bool live = (uint32_t)r * NT + t < tb.N;
#pragma unroll 1
        for (int w = W - 1; w >= 0; w--) {
            const int ci = r * W + w;
            Fe qx = ldg_fe(tb.gx + (size_t)(2 * r) * NT + t, tb.gx + (size_t)(2 * r + 1) * NT + t);
            Fe invd = inv;
            Fe d;
            { Fe px = lds_fe(s_win + w * 32); fp::fe_sub_canon(d.v, px.v, qx.v); }
            const bool eq = fp::fe_is_zero(d.v);
            if (eq) d = lds_fe(s_win + w * 32 + 24);
            if (ci > 0) {
                size_t cp = (size_t)(ci - 1) * 2 * NT + t;
                Fe cprev; gs::ld8_stream(cprev.v, tb.chain + cp, tb.chain + cp + NT);
                invd = mulf(inv, cprev);
                inv = mulf(inv, d);
            }
            uint64_t keym, keyp = 0;
            if (!eq) {
                Fe qy = ldg_fe(tb.gy + (size_t)(2 * r) * NT + t, tb.gy + (size_t)(2 * r + 1) * NT + t);
                Fe py = lds_fe(s_win + w * 32 + 8), u, s;
                { Fe npx = lds_fe(s_win + w * 32 + 16); fp::fe_sub_canon(u.v, npx.v, qx.v); }     // -(Px + Qx)
                fp::fe_add_lazy(s.v, py.v, qy.v);                                                  // P - Q  (code 2)
                Fe l2 = sqrf(mulf(s, invd));
                keym = fp::fe_add_low64_canon(l2.v, u.v);
                fp::fe_sub_canon(s.v, py.v, qy.v);                                                 // P + Q  (code 1)
                l2 = sqrf(mulf(s, invd));
                keyp = fp::fe_add_low64_canon(l2.v, u.v);
            } else {
                Fe px = lds_fe(s_win + w * 32);
                keym = gs::double_x_low64(px.v, invd.v);                                           // X(2P)  (code 4)
            }
            if (fill == 0) ci_first = ci;
            ring[fill * kBlock] = keym; ring[(fill + 1) * kBlock] = keyp;
            eqmask |= (eq ? 1u : 0u) << (fill >> 1);
            deadmask |= (live ? 0u : 1u) << (fill >> 1);
            fill += 2;
            if (fill == kRing) { drain<BW>(tb, ring, fill, ci_first, W, t, eqmask, deadmask); fill = 0; eqmask = 0; deadmask = 0; }
        }
    }
    if (fill) drain<BW>(tb, ring, fill, ci_first, W, t, eqmask, deadmask);
}

}  // namespace gs2
