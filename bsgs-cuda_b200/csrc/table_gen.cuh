// table_gen.cuh -- device-side generation of the giant array Q_k = (k+1)*ADDPUBG, ADDPUBG = -(2w)G
// (reference: giant(0) single-threaded on the CPU, 1_9_7File.pb:1418-1488 with fillarrayN,
// lib/Curve64.pb:3033-3064; minutes at N ~ 2*10^7) directly in the kernel's 128-bit SoA layout, and
// the export back into the reference's G2 file order (Writeint, 1_9_7File.pb:1831-1903).
//
// Doubling schedule, the same idea as fillarrayN: once Q_0..Q_{m-1} exist, Q_{m+i} = Q_i + m*D for
// i < m-1 are independent affine additions with one common addend (one thread each), and
// Q_{2m-1} = 2*(m*D) is the next addend, which the host derives by one doubling.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "fp.cuh"

namespace tgen {

struct PtConst { uint32_t x[8], y[8]; };                 // canonical, little-endian limbs

__device__ __forceinline__ void ldslot(uint32_t r[8], const uint4* base, uint32_t NT, uint64_t k) {
    uint64_t row = k / NT, col = k % NT;
    uint4 a = base[(2 * row) * NT + col], b = base[(2 * row + 1) * NT + col];
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void stslot(uint4* base, uint32_t NT, uint64_t k, const uint32_t r[8]) {
    uint64_t row = k / NT, col = k % NT;
    base[(2 * row) * NT + col] = make_uint4(r[0], r[1], r[2], r[3]);
    base[(2 * row + 1) * NT + col] = make_uint4(r[4], r[5], r[6], r[7]);
}

__global__ void k_set_slot(uint4* gx, uint4* gy, uint32_t NT, uint64_t k, const __grid_constant__ PtConst A) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { stslot(gx, NT, k, A.x); stslot(gy, NT, k, A.y); }
}

// slots [m, m+cnt) <- slots [0, cnt) + A      (A.x differs from every source x: cnt <= m-1)
__global__ void __launch_bounds__(128) k_giants_extend(uint4* gx, uint4* gy, uint32_t NT, uint64_t m, uint64_t cnt,
                                                      const __grid_constant__ PtConst A) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    uint32_t x1[8], y1[8], d[8], inv[8], s[8], lam[8], x3[8], y3[8], t[8];
    ldslot(x1, gx, NT, i); ldslot(y1, gy, NT, i);
    fp::fe_sub_canon(d, A.x, x1);
    fp::fe_inv(inv, d);
    fp::fe_sub_canon(s, A.y, y1);
    fp::fe_mul(lam, s, inv);
    fp::fe_sqr(x3, lam); fp::fe_canon(x3);
    fp::fe_sub_canon(t, x3, x1); fp::fe_sub_canon(x3, t, A.x);
    fp::fe_sub_canon(t, x1, x3);
    fp::fe_mul(y3, lam, t); fp::fe_canon(y3);
    fp::fe_sub_canon(t, y3, y1);
    stslot(gx, NT, m + i, x3); stslot(gy, NT, m + i, t);
}

// padding slots [N, total) replicate giant 0 (they are computed but never reported)
__global__ void k_pad_slots(uint4* gx, uint4* gy, uint32_t NT, uint64_t N, uint64_t total) {
    uint64_t k = N + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    uint32_t x[8], y[8];
    ldslot(x, gx, NT, 0); ldslot(y, gy, NT, 0);
    stslot(gx, NT, k, x); stslot(gy, NT, k, y);
}

// kernel layout -> G2 file image (inverse of gs::k_convert_g2)
__global__ void k_export_g2(uint32_t* __restrict__ g2, const uint4* __restrict__ gx, const uint4* __restrict__ gy,
                            uint32_t T, uint32_t B, uint32_t p, uint32_t N, uint32_t NT) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    uint32_t x[8], y[8];
    ldslot(x, gx, NT, k); ldslot(y, gy, NT, k);
    uint64_t TB = (uint64_t)T * B;
    uint32_t blk = (uint32_t)(k / ((uint64_t)T * p)), thr = (uint32_t)((k % ((uint64_t)T * p)) / p), r = (uint32_t)(k % p);
    uint64_t g = (uint64_t)blk * T + thr;
    uint32_t* X = g2 + (size_t)r * 8 * TB + g;
    uint32_t* Y = X + (size_t)N * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) { X[(size_t)i * TB] = x[7 - i]; Y[(size_t)i * TB] = y[7 - i]; }
}

}  // namespace tgen
