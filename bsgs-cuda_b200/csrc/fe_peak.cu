// fe_peak.cu -- what the field arithmetic of fp.cuh delivers on its own: back-to-back fe_mul / fe_sqr
// on register operands, no memory traffic, no probe, the kernel's own geometry (3 CTAs x 256 threads
// per SM at <= 80 registers, 2 independent chains per thread).  One giant step of the search costs
// 2.5 multiplications + 1 squaring, so
//     ceiling_steps_per_s = 1 / (2.5 * t_mul + t_sqr)
// is the speed the giant-step kernel would reach if everything around the multiplications (field
// additions, loads, the probe, loop control) were free.  Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp.cuh"

#define ITER 2000

template <int OP>
__global__ void __launch_bounds__(256, 3) spin(uint32_t* out, uint32_t seed) {
    uint32_t a[8], b[8], c[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * 0x9E3779B9u + threadIdx.x * 977u + i * 0x85EBCA6Bu; b[i] = a[i] * 3u + blockIdx.x; c[i] = a[i] ^ 0x5bd1e995u; d[i] = b[i] + 17u * i; }
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        uint32_t r[8], s[8];
        if (OP == 0) { fp::fe_mul(r, a, b); fp::fe_mul(s, c, d); }      // two independent products per iteration
        else { fp::fe_sqr(r, a); fp::fe_sqr(s, c); }
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = r[i]; c[i] = s[i]; }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i] ^ c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// The arithmetic of one point-window exactly as the giant-step kernel does it (pass 1 + pass 2:
// 5 multiplications, 2 unreduced squarings, 3 canonical subtractions, 1 lazy addition, 2 key extractions from the wide square),
// window constants in shared memory, giant coordinates and chain values in registers -- no global
// memory, no probe, no ring.  Two giant steps per iteration.
__global__ void __launch_bounds__(256, 3) spin_pw(uint32_t* out, uint32_t seed) {
    __shared__ __align__(16) uint32_t s_win[32];
    if (threadIdx.x < 32) s_win[threadIdx.x] = seed * 0x9E3779B9u + threadIdx.x * 0x85EBCA6Bu;
    __syncthreads();
    uint32_t qx[8], qy[8], inv[8], acc[8], cprev[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { qx[i] = seed + threadIdx.x * 977u + i * 0x85EBCA6Bu; qy[i] = qx[i] * 3u + blockIdx.x; inv[i] = qx[i] ^ 0x5bd1e995u; acc[i] = qy[i] + 17u * i; cprev[i] = acc[i] ^ qx[i]; }
    qx[7] &= 0x7FFFFFFFu; qy[7] &= 0x7FFFFFFFu;                 // canonical inputs for the subtractions
    uint64_t sink = 0;
#pragma unroll 1
    for (int it = 0; it < ITER / 4; it++) {
        uint32_t px[8], py[8], npx[8], d[8], invd[8], t[8], sum[8], lam[8], l2[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { px[i] = s_win[i]; py[i] = s_win[8 + i]; npx[i] = s_win[16 + i]; }
        px[7] &= 0x7FFFFFFFu; py[7] &= 0x7FFFFFFFu; npx[7] &= 0x7FFFFFFFu;
        fp::fe_sub_canon(d, px, qx);
        fp::fe_mul(t, acc, d);                                   // pass 1
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = t[i];
        fp::fe_sub_canon(d, px, cprev);                          // pass 2 forms d again (a different operand keeps the compiler from sharing it)
        fp::fe_mul(invd, inv, cprev);
        fp::fe_mul(t, inv, d);
#pragma unroll
        for (int i = 0; i < 8; i++) inv[i] = t[i];
        const uint64_t t_lo = ((((uint64_t)npx[1]) << 32) | npx[0]) - ((((uint64_t)qx[1]) << 32) | qx[0]);
        const int64_t t7 = (int64_t)npx[7] - (int64_t)qx[7];
        uint32_t v[16]; bool ex;
        fp::fe_add_lazy(sum, py, qy);
        fp::fe_mul(lam, sum, invd); fp::fe_sqr_wide(v, lam);
        sink ^= fp::fe_key_from_wide(v, t_lo, t7, FP_KEY_THR, ex);
        fp::fe_sub_canon(sum, py, qy);
        fp::fe_mul(lam, sum, invd); fp::fe_sqr_wide(v, lam);
        sink += fp::fe_key_from_wide(v, t_lo, t7, FP_KEY_THR, ex) + (ex ? 1 : 0);
#pragma unroll
        for (int i = 0; i < 8; i++) l2[i] = v[i] ^ v[8 + i];
#pragma unroll
        for (int i = 0; i < 7; i++) { qx[i] ^= l2[i]; qy[i] += lam[i]; cprev[i] ^= invd[i]; }      // next "giant"
        cprev[7] = (cprev[7] ^ invd[7]) & 0x7FFFFFFFu;
    }
    uint32_t x = (uint32_t)sink ^ (uint32_t)(sink >> 32);
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= acc[i] ^ inv[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
static double run_pw(int sms, uint32_t* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    spin_pw<<<sms * 3, 256>>>(out, 1); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); spin_pw<<<sms * 3, 256>>>(out, 2 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return (double)sms * 3 * 256 * (ITER / 4) * 2 / (best * 1e-3);   // giant steps per second, whole GPU
}

template <int OP>
static double run(int sms, uint32_t* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    spin<OP><<<sms * 3, 256>>>(out, 1); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); spin<OP><<<sms * 3, 256>>>(out, 2 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return (double)sms * 3 * 256 * ITER * 2 / (best * 1e-3);          // operations per second, whole GPU
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * sms * 3 * 256);
    double mul = run<0>(sms, out), sqr = run<1>(sms, out);
    double ceil_steps = 1.0 / (2.5 / mul + 1.0 / sqr);
    double pw = run_pw(sms, out);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fe_mul_per_s\": %.4g, \"fe_sqr_per_s\": %.4g, \"giant_step_ceiling_per_s\": %.4g, "
           "\"point_window_arithmetic_steps_per_s\": %.4g, "
           "\"how\": \"2.5 fe_mul + 1 fe_sqr per giant step, 3 CTAs x 256 threads per SM, registers only; point_window_arithmetic = the whole field arithmetic of one point-window (5 mul, 2 unreduced sqr, 3 sub, 1 add, 2 keys from the wide square) without memory or probe, ONE launch (its uneven tail included)\", \"err\": \"%s\"}\n",
           pr.name, sms, mul, sqr, ceil_steps, pw, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
