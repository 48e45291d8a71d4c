// imad_peak.cu -- measures the integer-pipe issue rates that bound the giant-step kernel.
//
// BASELINE.md section 3 defines the roofline denominator as
//     peak_M32/s = SMs x (IMAD lanes / clk / SM) x f_SM
// and requires the lane rate to be MEASURED with a dependency-free IMAD.WIDE.U32 micro-kernel on
// the box.  This tool times several instruction mixes (8 independent chains per thread, 1024
// threads per SM, one CTA per SM) and prints one JSON object; bench.py stores it beside the
// bench line and profiles/ keeps a copy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 8

enum Mix { WIDE = 0, WIDE_X, LO, HI, IADD3, WIDE_IADD3, DFMA, FFMA, NMIX };
static const char* kNames[NMIX] = {"imad_wide_u32", "imad_wide_u32_x_carry", "imad_lo", "imad_hi_u32", "iadd3", "imad_wide_plus_iadd3", "dfma", "ffma"};
// ops counted per inner step per chain
static const int kOps[NMIX] = {1, 1, 1, 1, 1, 2, 1, 1};

__constant__ uint32_t cb[64];

template <int MIX>
__global__ void __launch_bounds__(1024, 1) spin(uint64_t* out, uint32_t seed, long long* cycles) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1 + (threadIdx.x << 3);
    uint32_t lo[CHAINS], hi[CHAINS], x[CHAINS], mm[CHAINS]; uint64_t acc[CHAINS];
    double d[CHAINS]; float f[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        lo[i] = (i + seed) * 0x9E3779B9u + threadIdx.x; acc[i] = (uint64_t)lo[i] * 0x9E3779B97F4A7C15ull; hi[i] = (i + seed) * 0x7F4A7C15u + threadIdx.x * 5;
        x[i] = i * 7 + seed + threadIdx.x * 31; mm[i] = (i + 3) * 0x85EBCA6Bu + threadIdx.x * 977 + seed; d[i] = i + 1.0 + threadIdx.x * 1e-9; f[i] = i + 1.0f + threadIdx.x * 1e-7f;
    }
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < ITER; it++) {
        // the multiplier changes every iteration (uniform load from constant memory) so ptxas cannot
        // hoist the products; the multiplicands are per-thread, per-chain registers
        b = cb[it & 63];
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            uint32_t m = mm[i], xm = mm[(i + 1) % CHAINS];
#define WIDE_ASM(first, last)                                                                                   \
    asm volatile("{ .reg .u32 l, h; mov.b64 {l,h}, %0; " first " l, %1, %2, l; " last " h, %1, %2, h; mov.b64 %0, {l,h}; }" \
                 : "+l"(acc[i]) : "r"(m), "r"(b))
            // lo/hi pair == one 64-bit multiply-accumulate: ptxas emits IMAD.WIDE.U32 (.X when carry-chained)
            if (MIX == WIDE || MIX == WIDE_IADD3) WIDE_ASM("mad.lo.cc.u32", "madc.hi.u32");
            if (MIX == WIDE_X) {                        // carry chained through the 8 accumulators
                if (i == 0) WIDE_ASM("mad.lo.cc.u32", "madc.hi.cc.u32");
                else if (i < CHAINS - 1) WIDE_ASM("madc.lo.cc.u32", "madc.hi.cc.u32");
                else WIDE_ASM("madc.lo.cc.u32", "madc.hi.u32");
            }
            if (MIX == LO) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(xm), "r"(b));
            if (MIX == HI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(xm), "r"(b));
            if (MIX == IADD3 || MIX == WIDE_IADD3) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(xm), "r"(a));   // one IADD3
            if (MIX == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0 + b * 1e-12), "d"(0.5));
            if (MIX == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0f + b * 1e-12f), "f"(0.5f));
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i] + lo[i] + hi[i] + x[i] + (uint64_t)d[i] + (uint64_t)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX>
static void run(int sms, uint64_t* out, long long* cyc, int clock_khz, bool last) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    spin<MIX><<<sms, 1024>>>(out, 1, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; r++) spin<MIX><<<sms, 1024>>>(out, r + 2, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[sms]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += (double)h[i]; avg /= sms;
    double lane_ops_per_sm = 1024.0 * CHAINS * ITER * kOps[MIX];
    double per_clk = lane_ops_per_sm / avg;                                   // lanes/clk/SM by the SM's own cycle counter
    double gops = lane_ops_per_sm * sms * reps / (ms * 1e-3) / 1e9;           // whole chip, wall clock
    printf("  \"%s\": {\"lanes_per_clk_per_sm\": %.2f, \"gops\": %.1f, \"implied_mhz\": %.0f}%s\n", kNames[MIX], per_clk, gops,
           gops * 1e9 / (per_clk * sms) / 1e6, last ? "" : ",");
    delete[] h; (void)clock_khz;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint64_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint64_t) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
    uint32_t hcb[64]; for (int i = 0; i < 64; i++) hcb[i] = 0x9E3779B9u * (i + 1) | 1u;
    cudaMemcpyToSymbol(cb, hcb, sizeof hcb);
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"attr_clock_mhz\": %d,\n", pr.name, sms, khz / 1000);
    run<WIDE>(sms, out, cyc, khz, false); run<WIDE_X>(sms, out, cyc, khz, false); run<LO>(sms, out, cyc, khz, false);
    run<HI>(sms, out, cyc, khz, false); run<IADD3>(sms, out, cyc, khz, false); run<WIDE_IADD3>(sms, out, cyc, khz, false);
    run<DFMA>(sms, out, cyc, khz, false); run<FFMA>(sms, out, cyc, khz, true);
    printf("}\n");
    return 0;
}
