// verify.cuh -- device-side verification of hit records (reference: checkerThread,
// 1_9_7File.pb:3933-4296, which does this on one CPU thread with a table look-up in the htCPU file
// and several scalar multiplications per record).
//
// The GPU table matches htsz + 32 bits of X, so at w = 2^30 / htsz = 28 about one probe in 2^30 is a
// false positive: ~30 records per second per GPU at 24 G giant-steps/s.  Here every record is
// re-examined where it was produced: the candidate point R = P_w +- Q_idx is recomputed exactly,
// its bucket is read from the CSR image (tags + payloads j), and for every entry with the same tag
// the baby point (j+1)G is recomputed (32-bit scalar, Jacobian double-and-add) and compared with R
// on all 256 bits of X.  Only records that survive are handed to the host, with the payload j.
// One thread per record; a few hundred microseconds each, only on launches that reported something.
#pragma once
#include <stdint.h>
#include "fp.cuh"
#include "giant_step.cuh"

namespace vfy {

using gs::Fe;

__device__ __noinline__ Fe mulc(Fe a, Fe b) { Fe r; fp::fe_mul(r.v, a.v, b.v); fp::fe_canon(r.v); return r; }
__device__ __forceinline__ Fe sqrc(Fe a) { return mulc(a, a); }
__device__ __noinline__ Fe subc(Fe a, Fe b) { Fe r; fp::fe_sub_canon(r.v, a.v, b.v); return r; }
__device__ __noinline__ Fe addc(Fe a, Fe b) { Fe r; fp::fe_add_lazy(r.v, a.v, b.v); fp::fe_canon(r.v); return r; }
__device__ __noinline__ Fe invc(Fe a) { Fe r; fp::fe_inv(r.v, a.v); fp::fe_canon(r.v); return r; }
__device__ __forceinline__ Fe dblc(Fe a) { return addc(a, a); }
__device__ __forceinline__ Fe load_fe(const uint32_t* p) { Fe r; for (int i = 0; i < 8; i++) r.v[i] = p[i]; return r; }

struct Jac { Fe X, Y, Z; };

// dbl-2009-l (a = 0)
__device__ __noinline__ Jac jac_double(Jac p) {
    Fe A = sqrc(p.X), B = sqrc(p.Y), C = sqrc(B);
    Fe t = sqrc(addc(p.X, B));
    Fe D = dblc(subc(subc(t, A), C));
    Fe E = addc(dblc(A), A), F = sqrc(E);
    Jac r;
    r.X = subc(F, dblc(D));
    Fe C8 = dblc(dblc(dblc(C)));
    r.Y = subc(mulc(E, subc(D, r.X)), C8);
    r.Z = dblc(mulc(p.Y, p.Z));
    return r;
}
// madd-2007-bl (Z2 = 1); the operands never have equal x here (see kG_x)
__device__ __noinline__ Jac jac_add_affine(Jac p, Fe x2, Fe y2) {
    Fe Z1Z1 = sqrc(p.Z), U2 = mulc(x2, Z1Z1), S2 = mulc(mulc(y2, p.Z), Z1Z1);
    Fe H = subc(U2, p.X), HH = sqrc(H), I = dblc(dblc(HH)), J = mulc(H, I);
    Fe rr = dblc(subc(S2, p.Y)), V = mulc(p.X, I);
    Jac r;
    r.X = subc(subc(sqrc(rr), J), dblc(V));
    r.Y = subc(mulc(rr, subc(V, r.X)), dblc(mulc(p.Y, J)));
    r.Z = subc(subc(sqrc(addc(p.Z, H)), Z1Z1), HH);
    return r;
}

// canonical affine X of k*G, 1 <= k < 2^33.  MSB-first: the accumulator is m*G with m >= 2 whenever G
// is added, so the addition never degenerates.
__device__ __noinline__ Fe kG_x(uint64_t k) {
    const Fe gx = {{0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu, 0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu}};
    const Fe gy = {{0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u, 0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u}};
    if (k == 1) return gx;
    Jac a; a.X = gx; a.Y = gy; a.Z = Fe{{1, 0, 0, 0, 0, 0, 0, 0}};
    int top = 63 - __clzll((long long)k);
#pragma unroll 1
    for (int b = top - 1; b >= 0; b--) {
        a = jac_double(a);
        if ((k >> b) & 1) a = jac_add_affine(a, gx, gy);
    }
    Fe zi = invc(a.Z);
    return mulc(a.X, sqrc(zi));
}

// records: 3 u32 each (code, idx, window); out: per record 0xFFFFFFFF = rejected, else payload j
// (codes 1, 2, 5) or 0xFFFFFFFE (code 4: accepted without a table entry, as the reference's checker
// handles it, 1_9_7File.pb:3974-4019).
__global__ void __launch_bounds__(32) k_verify_hits(const gs::Tables tb, const __grid_constant__ gs::WinParams wp, const uint32_t* __restrict__ js,
                                                    uint32_t cnt, uint32_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const uint32_t code = tb.hits[3 * i], idx = tb.hits[3 * i + 1], win = tb.hits[3 * i + 2];
    if (win >= (uint32_t)gs::kMaxW) { out[i] = 0xFFFFFFFFu; return; }
    if (code == 4) { out[i] = 0xFFFFFFFEu; return; }
    Fe px = load_fe(wp.win[win].px), py = load_fe(wp.win[win].py), X;
    if (code == 5) X = px;
    else {
        if (idx >= tb.N) { out[i] = 0xFFFFFFFFu; return; }
        const uint32_t r = idx / tb.NT, t = idx - r * tb.NT;
        Fe qx = gs::ldg_fe(tb.gx + (size_t)(2 * r) * tb.NT + t, tb.gx + (size_t)(2 * r + 1) * tb.NT + t);
        Fe qy = gs::ldg_fe(tb.gy + (size_t)(2 * r) * tb.NT + t, tb.gy + (size_t)(2 * r + 1) * tb.NT + t);
        Fe num = code == 1 ? subc(py, qy) : addc(py, qy);          // P + Q : P - Q
        Fe lam = mulc(num, invc(subc(px, qx)));
        X = subc(subc(sqrc(lam), px), qx);
    }
    const uint32_t h = X.v[0] & tb.mask, tag = X.v[1];
    uint32_t res = 0xFFFFFFFFu;
    for (uint32_t pos = tb.csr_off[h]; pos < tb.csr_off[h + 1] && res == 0xFFFFFFFFu; pos++) {
        if (tb.csr_tags[pos] != tag) continue;
        const uint32_t j = js[pos];
        Fe bx = kG_x((uint64_t)j + 1);
        if (fp::fe_eq(bx.v, X.v)) res = j;
    }
    out[i] = res;
}

// htCPU item stream {tag, j} -> tags[] and js[] (upload of an htCPU file image)
__global__ void k_split_items(const uint2* __restrict__ items, uint32_t* __restrict__ tags, uint32_t* __restrict__ js, uint64_t w) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < w; i += stride) { uint2 it = items[i]; tags[i] = it.x; js[i] = it.y; }
}

}  // namespace vfy
