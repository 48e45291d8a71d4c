// pipe_mix.cu -- how much other work fits beside IMAD.WIDE?  Measures cycles per inner step for
// mixes of A x IMAD.WIDE.U32 (64-bit accumulate) and B x {IADD3 | LOP3 | IMAD.LO | MOV-like} per step,
// 8 independent chains per thread, 1024 threads per SM (8 warps per scheduler).  Prints JSON.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define CH 8
__constant__ uint32_t cb[64];

// KIND: 0 IADD3, 1 SHF (funnel shift), 2 IMAD.LO, 3 SETP+SEL, 8 LOP3 with 3 register reads; KIND2 additionally 4 FFMA, 5 DFMA, 6 HSET2 (fp16x2 compare)
template <int NW, int NB, int KIND, int NB2 = 0, int KIND2 = 0>
__global__ void __launch_bounds__(1024, 1) mix(uint64_t* out, uint32_t seed, long long* cycles) {
    uint64_t acc[CH]; uint32_t x[CH], y[CH], mm[CH]; float f[CH]; double df[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { acc[i] = (uint64_t)(i + seed) * 0x9E3779B97F4A7C15ull + threadIdx.x; x[i] = i * 7 + seed + threadIdx.x * 31; y[i] = x[i] * 3; f[i] = 1.0f + i; df[i] = 1.0 + i; mm[i] = (i + 3) * 0x85EBCA6Bu + threadIdx.x * 977 + seed; }
    uint32_t a = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITER; it++) {
        uint32_t b = cb[it & 63];
#pragma unroll
        for (int i = 0; i < CH; i++) {
#pragma unroll
            for (int k = 0; k < NW; k++)
                asm volatile("{ .reg .u32 l, h; mov.b64 {l,h}, %0; mad.lo.cc.u32 l, %1, %2, l; madc.hi.u32 h, %1, %2, h; mov.b64 %0, {l,h}; }" : "+l"(acc[i]) : "r"(mm[i]), "r"(b));
#pragma unroll
            for (int k = 0; k < NB; k++) {
                if (KIND == 0) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(mm[(i + 1) % CH]), "r"(a));
                if (KIND == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(mm[(i + 1) % CH]));
                if (KIND == 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(mm[(i + 1) % CH]), "r"(b));
                if (KIND == 3) asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; selp.u32 %0, %0, %1, p; }" : "+r"(x[i]) : "r"(mm[(i + 1) % CH]), "r"(b));
                if (KIND == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(mm[(i + 1) % CH]), "r"(mm[(i + 2) % CH]));   // three register reads
            }
#pragma unroll
            for (int k = 0; k < NB2; k++) {
                if (KIND2 == 0) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(mm[(i + 2) % CH]), "r"(a));
                if (KIND2 == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(y[i]) : "r"(mm[(i + 2) % CH]));
                if (KIND2 == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f));
                if (KIND2 == 5) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(df[i]) : "d"(1.0001), "d"(0.5));
                if (KIND2 == 6) asm volatile("set.ne.u32.f16x2 %0, %0, %1;" : "+r"(y[i]) : "r"(mm[(i + 2) % CH]));
            }
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i] + x[i] + y[i] + (uint64_t)f[i] + (uint64_t)df[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NW, int NB, int KIND, int NB2 = 0, int KIND2 = 0>
static void run(const char* name, int sms, uint64_t* out, long long* cyc, bool last) {
    mix<NW, NB, KIND, NB2, KIND2><<<sms, 1024>>>(out, 1, cyc); cudaDeviceSynchronize();
    mix<NW, NB, KIND, NB2, KIND2><<<sms, 1024>>>(out, 2, cyc); cudaDeviceSynchronize();
    long long* h = new long long[sms]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += (double)h[i]; avg /= sms;
    // per scheduler: 8 warps x CH chains x ITER steps; cycles per (warp-step) on one scheduler
    double per_step = avg / (8.0 * CH * ITER);
    printf("  \"%s\": {\"wide\": %d, \"other\": %d, \"sched_cycles_per_step\": %.3f}%s\n", name, NW, NB, per_step, last ? "" : ",");
    delete[] h;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    uint64_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint64_t) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
    uint32_t hcb[64]; for (int i = 0; i < 64; i++) hcb[i] = 0x9E3779B9u * (i + 1) | 1u;
    cudaMemcpyToSymbol(cb, hcb, sizeof hcb);
    printf("{\n");
    run<1, 0, 0>("wide1", sms, out, cyc, false);
    run<0, 1, 0>("iadd3_1", sms, out, cyc, false);
    run<0, 1, 1>("shf_1", sms, out, cyc, false);
    run<0, 1, 2>("imadlo_1", sms, out, cyc, false);
    run<1, 1, 0>("wide1_iadd3_1", sms, out, cyc, false);
    run<1, 2, 0>("wide1_iadd3_2", sms, out, cyc, false);
    run<1, 3, 0>("wide1_iadd3_3", sms, out, cyc, false);
    run<1, 4, 0>("wide1_iadd3_4", sms, out, cyc, false);
    run<1, 1, 1>("wide1_shf_1", sms, out, cyc, false);
    run<1, 2, 1>("wide1_shf_2", sms, out, cyc, false);
    run<1, 3, 1>("wide1_shf_3", sms, out, cyc, false);
    run<1, 1, 2>("wide1_imadlo_1", sms, out, cyc, false);
    run<1, 2, 2>("wide1_imadlo_2", sms, out, cyc, false);
    run<1, 2, 3>("wide1_setp_sel_2", sms, out, cyc, false);
    run<0, 2, 2>("imadlo_2", sms, out, cyc, false);
    run<0, 1, 3>("setp_sel_1", sms, out, cyc, false);
    run<0, 1, 2, 1, 0>("imadlo_1_iadd3_1", sms, out, cyc, false);
    run<0, 1, 2, 2, 0>("imadlo_1_iadd3_2", sms, out, cyc, false);
    run<0, 2, 2, 2, 0>("imadlo_2_iadd3_2", sms, out, cyc, false);
    run<0, 1, 2, 1, 4>("imadlo_1_ffma_1", sms, out, cyc, false);
    run<0, 1, 0, 1, 4>("iadd3_1_ffma_1", sms, out, cyc, false);
    run<0, 0, 0, 1, 4>("ffma_1", sms, out, cyc, false);
    run<0, 0, 0, 2, 4>("ffma_2", sms, out, cyc, false);
    run<1, 0, 0, 1, 4>("wide1_ffma_1", sms, out, cyc, false);
    run<1, 0, 0, 2, 4>("wide1_ffma_2", sms, out, cyc, false);
    run<2, 3, 0>("wide2_iadd3_3", sms, out, cyc, false);
    run<1, 1, 8>("wide1_lop3r_1", sms, out, cyc, false);
    run<1, 2, 8>("wide1_lop3r_2", sms, out, cyc, false);
    run<1, 3, 8>("wide1_lop3r_3", sms, out, cyc, false);
    run<0, 0, 0, 1, 5>("dfma_1", sms, out, cyc, false);
    run<1, 0, 0, 1, 5>("wide1_dfma_1", sms, out, cyc, false);
    run<1, 0, 0, 2, 5>("wide1_dfma_2", sms, out, cyc, false);
    run<1, 2, 0, 1, 5>("wide1_iadd3_2_dfma_1", sms, out, cyc, false);
    run<0, 2, 0, 1, 5>("iadd3_2_dfma_1", sms, out, cyc, false);
    run<0, 0, 0, 1, 6>("hset2_1", sms, out, cyc, false);
    run<1, 0, 0, 1, 6>("wide1_hset2_1", sms, out, cyc, false);
    run<1, 0, 0, 2, 6>("wide1_hset2_2", sms, out, cyc, true);
    printf("}\n");
    return 0;
}
