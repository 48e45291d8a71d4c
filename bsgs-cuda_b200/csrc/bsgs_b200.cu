// bsgs_b200.cu -- C ABI of libbsgs_b200.so (see include/bsgs_b200.h for the contract and the
// reference interfaces each entry point replaces).
#include "../../include/bsgs_b200.h"
#include "giant_step.cuh"
#include "table_build.cuh"
#include "table_gen.cuh"
#include "verify.cuh"
#include "host_ec.h"
#include <cuda_runtime.h>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct bsgs_ctx {
    int dev = 0, last_cuda = 0, sm_count = 0;
    uint32_t T = 0, B = 0, p = 0, htsz = 0;
    uint64_t w = 0, N = 0, H = 0;
    // geometry
    uint32_t blocks_per_sm = 3, grid = 0, NT = 0, PP = 0, max_windows = 16, block = 256;
    // device memory
    void* d_g2_stage = nullptr;      // raw G2 image (freed after commit)
    uint4 *d_gx = nullptr, *d_gy = nullptr, *d_chain = nullptr;
    void* d_ht = nullptr;            // raw htGPU image: off[H+1] | tags[w]  (kept: CSR fallback)
    uint32_t* d_buckets = nullptr;
    uint32_t* d_js = nullptr;        // payload j of every CSR item (htCPU image = tags interleaved with these)
    uint32_t* d_hits = nullptr;      // [0] counter, [4..] records (3 u32 each)
    unsigned long long* d_scratch = nullptr;
    uint32_t* d_verdict = nullptr;   // per record: payload j / accepted / rejected (device-side verification)
    uint32_t* h_verdict = nullptr;
    bool verify = false;
    bool force_exact = false;        // tests: every candidate key through the out-of-line exact path
    uint64_t* d_trace = nullptr;     // diagnostics: per-warp SM id and timestamps of the last launch (option "trace")
    gs::WinParams last_wp;           // window constants of the launch whose records are being collected
    uint32_t hit_cap = 1u << 16;
    uint32_t* h_hits = nullptr;      // pinned mirror
    size_t chain_bytes = 0, giants_bytes = 0, table_bytes = 0;
    bool giants_ready = false, ht_ready = false;
    unsigned long long overfull = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    // Second launch lane (bsgs_run / bsgs_time_launches): its own stream, chain, record buffers.  Launches alternate
    // between the lanes so that two are in flight: the warps of a launch do not finish together (the hardware
    // arbiter favours some warps: the first warp of an SM is done after 58 % of the kernel), and
    // the CTAs of the next launch take the slots they leave instead of the SM draining at every launch boundary.
    bool overlap = true;
    cudaStream_t stream2 = nullptr;
    uint4* d_chain2 = nullptr;
    uint32_t *d_hits2 = nullptr, *h_hits2 = nullptr, *d_verdict2 = nullptr, *h_verdict2 = nullptr;
    gs::WinParams last_wp2;
    cudaEvent_t e2 = nullptr;
};

struct Lane { cudaStream_t st; uint4* chain; uint32_t *d_hits, *h_hits, *d_verdict, *h_verdict; gs::WinParams* wp; };
static Lane lane_of(bsgs_ctx* c, int i) {
    if (i == 0) return Lane{c->stream, c->d_chain, c->d_hits, c->h_hits, c->d_verdict, c->h_verdict, &c->last_wp};
    return Lane{c->stream2, c->d_chain2, c->d_hits2, c->h_hits2, c->d_verdict2, c->h_verdict2, &c->last_wp2};
}

#define CU(call)                                                         \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) {                                         \
            ctx->last_cuda = (int)_e;                                    \
            cudaGetLastError();                                          \
            return _e == cudaErrorMemoryAllocation ? BSGS_E_NOMEM : BSGS_E_CUDA; \
        }                                                                \
    } while (0)

static int alloc_giants(bsgs_ctx* ctx);
static int ensure_lane2(bsgs_ctx* ctx);
static gs::Tables make_tables(const bsgs_ctx* c, int lane = 0);
static void fill_win(gs::WinConst& wc, const uint8_t* pxy);
static size_t ht_image_bytes(const bsgs_ctx* c) { return (size_t)c->H * 4 + 4 + (size_t)c->w * 4; }

extern "C" const char* bsgs_strerror(int s) {
    switch (s) {
        case BSGS_OK: return "ok";
        case BSGS_E_ARG: return "bad argument or size mismatch";
        case BSGS_E_CUDA: return "CUDA runtime error";
        case BSGS_E_NOMEM: return "out of memory";
        case BSGS_E_STATE: return "tables not uploaded";
        case BSGS_E_LIMIT: return "parameter outside structural limits";
        case BSGS_E_OVERFLOW: return "hit buffer overflow";
        default: return "unknown status";
    }
}
extern "C" int bsgs_last_cuda_error(const bsgs_ctx* ctx) { return ctx ? ctx->last_cuda : 0; }

static void compute_geometry(bsgs_ctx* c) {
    c->grid = (uint32_t)c->sm_count * c->blocks_per_sm;
    // never launch more threads than giants: shrink the grid for tiny problems (tests)
    while (c->grid > 1 && (uint64_t)(c->grid - 1) * c->block >= c->N) c->grid--;
    c->NT = c->grid * c->block;
    c->PP = (uint32_t)((c->N + c->NT - 1) / c->NT);
}

extern "C" int bsgs_create(bsgs_ctx** out, int device, uint32_t T, uint32_t B, uint32_t p, uint64_t w, uint32_t htsz) {
    if (!out) return BSGS_E_ARG;
    *out = nullptr;
    if (!T || !B || !p || !w) return BSGS_E_ARG;
    uint64_t N = (uint64_t)T * B * p;
    if (N >= (1ull << 31) || htsz > 31 || htsz < 1 || w >= (1ull << 32)) return BSGS_E_LIMIT;   // u32 indices/tags (F197:2584, 4412-4416)
    bsgs_ctx* ctx = new (std::nothrow) bsgs_ctx();
    if (!ctx) return BSGS_E_NOMEM;
    ctx->dev = device; ctx->T = T; ctx->B = B; ctx->p = p; ctx->w = w; ctx->htsz = htsz; ctx->N = N; ctx->H = 1ull << htsz;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->e0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->e1);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_hits, 16 + (size_t)ctx->hit_cap * 12);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_scratch, 64);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_verdict, (size_t)ctx->hit_cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_verdict, (size_t)ctx->hit_cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_hits, 16 + (size_t)ctx->hit_cap * 12);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_hits, 0, 16 + (size_t)ctx->hit_cap * 12, ctx->stream);      // whole buffer: the eager read-back copies record slots beyond the count
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { int rc = e == cudaErrorMemoryAllocation ? BSGS_E_NOMEM : BSGS_E_CUDA; bsgs_destroy(ctx); cudaGetLastError(); return rc; }
    compute_geometry(ctx);
    *out = ctx;
    return BSGS_OK;
}

extern "C" void bsgs_destroy(bsgs_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->dev);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_g2_stage); cudaFree(ctx->d_gx); cudaFree(ctx->d_gy); cudaFree(ctx->d_chain);
    cudaFree(ctx->d_ht); cudaFree(ctx->d_buckets); cudaFree(ctx->d_js); cudaFree(ctx->d_hits); cudaFree(ctx->d_scratch); cudaFree(ctx->d_verdict); cudaFree(ctx->d_trace);
    cudaFree(ctx->d_chain2); cudaFree(ctx->d_hits2); cudaFree(ctx->d_verdict2);
    if (ctx->h_hits2) cudaFreeHost(ctx->h_hits2);
    if (ctx->h_verdict2) cudaFreeHost(ctx->h_verdict2);
    if (ctx->e2) cudaEventDestroy(ctx->e2);
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->h_verdict) cudaFreeHost(ctx->h_verdict);
    if (ctx->h_hits) cudaFreeHost(ctx->h_hits);
    if (ctx->e0) cudaEventDestroy(ctx->e0);
    if (ctx->e1) cudaEventDestroy(ctx->e1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
}

extern "C" int bsgs_set_option(bsgs_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return BSGS_E_ARG;
    if (!strcmp(name, "verify")) { ctx->verify = value != 0; return BSGS_OK; }      // may be toggled at any time
    if (!strcmp(name, "force_exact")) { ctx->force_exact = value != 0; return BSGS_OK; }
    if (!strcmp(name, "overlap")) { ctx->overlap = value != 0; return BSGS_OK; }    // 0: one launch at a time (bsgs_run, bsgs_time_launches)
    if (!strcmp(name, "trace")) {                                                  // per-warp timestamps of every following launch
        CU(cudaSetDevice(ctx->dev));
        if (value && !ctx->d_trace) CU(cudaMalloc(&ctx->d_trace, (size_t)ctx->NT / 32 * 32));
        if (!value && ctx->d_trace) { cudaFree(ctx->d_trace); ctx->d_trace = nullptr; }
        return BSGS_OK;
    }
    if (ctx->d_gx || ctx->d_buckets) return BSGS_E_STATE;       // geometry is baked into the layouts
    if (!strcmp(name, "blocks_per_sm")) { if (value < 1 || value > 3) return BSGS_E_ARG; ctx->blocks_per_sm = (uint32_t)value; compute_geometry(ctx); return BSGS_OK; }
    if (!strcmp(name, "block")) {                               // threads per CTA: 256 (x blocks_per_sm), 320 (x 2), 128 (x 6) or 64 (x 12)
        if (value != 256 && value != 320 && value != 128 && value != 64) return BSGS_E_ARG;
        ctx->block = (uint32_t)value;
        if (value == 320) ctx->blocks_per_sm = 2;
        if (value == 128) ctx->blocks_per_sm = 6;
        if (value == 64) ctx->blocks_per_sm = 12;
        compute_geometry(ctx); return BSGS_OK;
    }
    if (!strcmp(name, "max_windows")) { if (value < 1 || value > BSGS_MAX_WINDOWS) return BSGS_E_ARG; ctx->max_windows = (uint32_t)value; return BSGS_OK; }
    if (!strcmp(name, "l2_fetch")) {                            // device-wide: DRAM->L2 fetch granularity for the random sector reads
        if (value != 32 && value != 64 && value != 128) return BSGS_E_ARG;
        CU(cudaSetDevice(ctx->dev)); CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value)); return BSGS_OK;
    }
    return BSGS_E_ARG;
}

extern "C" int bsgs_device_count(int* count) {
    if (!count) return BSGS_E_ARG;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return BSGS_E_CUDA; }
    return BSGS_OK;
}
extern "C" int bsgs_device_query(int device, bsgs_devinfo* out) {
    if (!out) return BSGS_E_ARG;
    memset(out, 0, sizeof *out);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return BSGS_E_CUDA; }
    if (device < 0 || device >= n) return BSGS_E_ARG;
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) { cudaGetLastError(); return BSGS_E_CUDA; }
    strncpy(out->name, pr.name, sizeof out->name - 1);
    out->cc_major = pr.major; out->cc_minor = pr.minor; out->sm_count = pr.multiProcessorCount;
    size_t fr = 0, tot = 0;
    int cur = 0; cudaGetDevice(&cur);
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); cudaSetDevice(cur); return BSGS_E_CUDA; }
    cudaSetDevice(cur);
    out->free_bytes = fr; out->total_bytes = tot;
    return BSGS_OK;
}
extern "C" int bsgs_get_info(const bsgs_ctx* ctx, bsgs_info* info) {
    if (!ctx || !info) return BSGS_E_ARG;
    info->grid = ctx->grid; info->block = ctx->block; info->points_per_thread = ctx->PP; info->sm_count = (uint32_t)ctx->sm_count;
    info->n_giants = ctx->N; info->bytes_giants = ctx->giants_bytes; info->bytes_chain = ctx->chain_bytes; info->bytes_table = ctx->table_bytes;
    return BSGS_OK;
}

// ------------------------------------------------------------------------------------------------ tables
extern "C" int bsgs_giants_staging(bsgs_ctx* ctx, void** dev_ptr, size_t* bytes) {
    if (!ctx || !dev_ptr || !bytes) return BSGS_E_ARG;
    CU(cudaSetDevice(ctx->dev));
    if (!ctx->d_g2_stage) CU(cudaMalloc(&ctx->d_g2_stage, (size_t)ctx->N * 64));
    *dev_ptr = ctx->d_g2_stage; *bytes = (size_t)ctx->N * 64;
    return BSGS_OK;
}
extern "C" int bsgs_giants_commit(bsgs_ctx* ctx) {
    if (!ctx || !ctx->d_g2_stage) return BSGS_E_STATE;
    CU(cudaSetDevice(ctx->dev));
    size_t slots = (size_t)ctx->NT * ctx->PP;
    { int rc = alloc_giants(ctx); if (rc) return rc; }
    uint32_t blocks = (uint32_t)((slots + 255) / 256);
    gs::k_convert_g2<<<blocks, 256, 0, ctx->stream>>>((const uint32_t*)ctx->d_g2_stage, ctx->d_gx, ctx->d_gy, ctx->T, ctx->B, ctx->p,
                                                     (uint32_t)ctx->N, ctx->NT, ctx->PP);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(ctx->d_g2_stage)); ctx->d_g2_stage = nullptr;
    ctx->giants_ready = true;
    return BSGS_OK;
}
extern "C" int bsgs_upload_giants(bsgs_ctx* ctx, const void* g2, size_t bytes) {
    if (!ctx || !g2 || bytes != (size_t)ctx->N * 64) return BSGS_E_ARG;
    void* d; size_t n; int rc = bsgs_giants_staging(ctx, &d, &n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(d, g2, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return bsgs_giants_commit(ctx);
}

// ---- device-side giant array (replaces giant(0) + Writeint, F197:1418-1488, 1831-1903)
static int alloc_giants(bsgs_ctx* ctx) {
    size_t slots = (size_t)ctx->NT * ctx->PP;
    if (!ctx->d_gx) {
        CU(cudaMalloc(&ctx->d_gx, slots * 32));
        CU(cudaMalloc(&ctx->d_gy, slots * 32));
        ctx->chain_bytes = slots * 32 * ctx->max_windows + 2 * (size_t)ctx->NT * 32;      // two leading rows: pad + the element "1" (giant_step.cuh)
        CU(cudaMalloc(&ctx->d_chain, ctx->chain_bytes));
        ctx->giants_bytes = slots * 64;
    }
    return BSGS_OK;
}
static void to_const(tgen::PtConst& c, const hec::Pt& p) { memcpy(c.x, p.x.v, 32); memcpy(c.y, p.y.v, 32); }
// slots [0, N) of (gx, gy) <- (k+1)*D, padding slots replicate slot 0
static int generate_multiples(bsgs_ctx* ctx, uint4* gx, uint4* gy, hec::Pt A) {
    tgen::PtConst c; to_const(c, A);
    tgen::k_set_slot<<<1, 32, 0, ctx->stream>>>(gx, gy, ctx->NT, 0, c);
    for (uint64_t m = 1; m < ctx->N; m *= 2) {
        uint64_t cnt = m - 1 < ctx->N - m ? m - 1 : ctx->N - m;
        if (cnt) tgen::k_giants_extend<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(gx, gy, ctx->NT, m, cnt, c);
        hec::pt_add(A, A, A);                                    // (2m)*D = slot 2m-1
        to_const(c, A);
        if (2 * m - 1 < ctx->N) tgen::k_set_slot<<<1, 32, 0, ctx->stream>>>(gx, gy, ctx->NT, 2 * m - 1, c);
    }
    uint64_t total = (uint64_t)ctx->NT * ctx->PP;
    if (total > ctx->N) tgen::k_pad_slots<<<(unsigned)((total - ctx->N + 255) / 256), 256, 0, ctx->stream>>>(gx, gy, ctx->NT, ctx->N, total);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return BSGS_OK;
}
extern "C" int bsgs_giants_generate(bsgs_ctx* ctx) {
    if (!ctx) return BSGS_E_ARG;
    CU(cudaSetDevice(ctx->dev));
    int rc = alloc_giants(ctx);
    if (rc) return rc;
    hec::Pt A;                                                   // D = ADDPUBG = -(2w)G  (F197:4689-4698)
    hec::pt_mul_u64(A, hec::kG, 2 * ctx->w);
    hec::pt_neg(A, A);
    rc = generate_multiples(ctx, ctx->d_gx, ctx->d_gy, A);
    if (rc) return rc;
    ctx->giants_ready = true;
    return BSGS_OK;
}
extern "C" int bsgs_giants_export(bsgs_ctx* ctx, void* g2, size_t bytes) {
    if (!ctx || !g2 || bytes != (size_t)ctx->N * 64) return BSGS_E_ARG;
    if (!ctx->giants_ready) return BSGS_E_STATE;
    CU(cudaSetDevice(ctx->dev));
    uint32_t* d = nullptr;
    CU(cudaMalloc(&d, bytes));
    tgen::k_export_g2<<<(unsigned)((ctx->N + 255) / 256), 256, 0, ctx->stream>>>(d, ctx->d_gx, ctx->d_gy, ctx->T, ctx->B, ctx->p, (uint32_t)ctx->N, ctx->NT);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(g2, d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CU(e);
    return BSGS_OK;
}

extern "C" int bsgs_ht_staging(bsgs_ctx* ctx, void** dev_ptr, size_t* bytes) {
    if (!ctx || !dev_ptr || !bytes) return BSGS_E_ARG;
    CU(cudaSetDevice(ctx->dev));
    if (!ctx->d_ht) CU(cudaMalloc(&ctx->d_ht, ht_image_bytes(ctx) + 64));
    *dev_ptr = ctx->d_ht; *bytes = ht_image_bytes(ctx);
    return BSGS_OK;
}
extern "C" int bsgs_ht_commit(bsgs_ctx* ctx) {
    if (!ctx || !ctx->d_ht) return BSGS_E_STATE;
    CU(cudaSetDevice(ctx->dev));
    size_t bbytes = (size_t)ctx->H * 32;                         // one 32-byte fingerprint sector per bucket
    if (!ctx->d_buckets) CU(cudaMalloc(&ctx->d_buckets, bbytes));
    ctx->table_bytes = bbytes + ht_image_bytes(ctx);
    const uint32_t* off = (const uint32_t*)ctx->d_ht;
    const uint32_t* tags = off + ctx->H + 1;
    CU(cudaMemsetAsync(ctx->d_scratch, 0, 8, ctx->stream));
    uint32_t blocks = (uint32_t)((ctx->H + 255) / 256);
    gs::k_build_fp_sectors<<<blocks, 256, 0, ctx->stream>>>(off, tags, (uint4*)ctx->d_buckets, ctx->H, ctx->d_scratch);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&ctx->overfull, ctx->d_scratch, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ht_ready = true;
    return BSGS_OK;
}
extern "C" int bsgs_upload_ht(bsgs_ctx* ctx, const void* htgpu, size_t bytes) {
    if (!ctx || !htgpu || bytes != ht_image_bytes(ctx)) return BSGS_E_ARG;
    void* d; size_t n; int rc = bsgs_ht_staging(ctx, &d, &n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(d, htgpu, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_js) { cudaFree(ctx->d_js); ctx->d_js = nullptr; }   // payloads of a previous device build are stale
    return bsgs_ht_commit(ctx);
}

extern "C" int bsgs_bcast_ht(bsgs_ctx* const* ctxs, int n) {
    if (!ctxs || n < 1 || !ctxs[0]) return BSGS_E_ARG;
    bsgs_ctx* src = ctxs[0];
    if (!src->ht_ready) return BSGS_E_STATE;
    const size_t img = ht_image_bytes(src), js = (size_t)src->w * 4;
    for (int i = 1; i < n; i++)
        if (!ctxs[i] || ctxs[i] == src || ctxs[i]->w != src->w || ctxs[i]->htsz != src->htsz) return BSGS_E_ARG;
    // phase 1: allocate, enable peer access where the devices differ, start every copy
    for (int i = 1; i < n; i++) {
        bsgs_ctx* ctx = ctxs[i];                                   // (CU() records errors in `ctx`)
        void* d; size_t nb; int rc = bsgs_ht_staging(ctx, &d, &nb);
        if (rc) return rc;
        if (src->d_js && !ctx->d_js) CU(cudaMalloc(&ctx->d_js, js + 16));
        if (!src->d_js && ctx->d_js) { cudaFree(ctx->d_js); ctx->d_js = nullptr; }
        if (ctx->dev != src->dev) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, ctx->dev, src->dev));
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(src->dev, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e); cudaGetLastError(); }
        }
        CU(cudaMemcpyPeerAsync(d, ctx->dev, src->d_ht, src->dev, img, ctx->stream));      // staged through the host when peers cannot map each other
        if (src->d_js) CU(cudaMemcpyPeerAsync(ctx->d_js, ctx->dev, src->d_js, src->dev, js, ctx->stream));
    }
    // phase 2: wait and build the fingerprint sectors on every destination
    for (int i = 1; i < n; i++) {
        bsgs_ctx* ctx = ctxs[i];
        CU(cudaSetDevice(ctx->dev));
        CU(cudaStreamSynchronize(ctx->stream));
        int rc = bsgs_ht_commit(ctx);
        if (rc) return rc;
    }
    return BSGS_OK;
}

// htCPU file image (4H + 4 + 8w bytes): the same table with the payloads j, which device-side
// verification needs.  Replaces bsgs_upload_ht for hosts that loaded <...>_htCPUv0.BIN.
extern "C" int bsgs_upload_htcpu(bsgs_ctx* ctx, const void* htcpu, size_t bytes) {
    size_t head = ((size_t)(ctx ? ctx->H : 0) + 1) * 4;
    if (!ctx || !htcpu || bytes != head + (size_t)ctx->w * 8) return BSGS_E_ARG;
    void* d; size_t n; int rc = bsgs_ht_staging(ctx, &d, &n);
    if (rc) return rc;
    if (!ctx->d_js) CU(cudaMalloc(&ctx->d_js, (size_t)ctx->w * 4 + 16));
    uint2* items = nullptr;
    CU(cudaMalloc(&items, (size_t)ctx->w * 8 + 16));
    cudaError_t e = cudaMemcpyAsync(d, htcpu, head, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(items, (const uint8_t*)htcpu + head, (size_t)ctx->w * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && ctx->w) {
        vfy::k_split_items<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(items, (uint32_t*)d + ctx->H + 1, ctx->d_js, ctx->w);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(items);
    CU(e);
    return bsgs_ht_commit(ctx);
}

// ---- device-side table construction (replaces HashTableInsert + sortHashTable32bit + packHT*File,
// F197:2555-2622, 2771-2820, 3232-3444, for tables that never have to exist on the host)
static int build_from_device_keys(bsgs_ctx* ctx, const uint64_t* d_keys) {
    void* d; size_t n; int rc = bsgs_ht_staging(ctx, &d, &n);
    if (rc) return rc;
    if (!ctx->d_js) CU(cudaMalloc(&ctx->d_js, (size_t)ctx->w * 4 + 16));
    uint32_t* cursor = nullptr;
    CU(cudaMalloc(&cursor, (size_t)ctx->H * 4));
    uint32_t* off = (uint32_t*)ctx->d_ht;
    cudaError_t e = tbl::build_csr(d_keys, ctx->w, ctx->H, off, off + ctx->H + 1, ctx->d_js, cursor, ctx->sm_count, ctx->stream);
    cudaFree(cursor);
    CU(e);
    return bsgs_ht_commit(ctx);
}
extern "C" int bsgs_ht_build_from_keys(bsgs_ctx* ctx, const uint64_t* keys, uint64_t nkeys) {
    if (!ctx || !keys || nkeys != ctx->w) return BSGS_E_ARG;
    CU(cudaSetDevice(ctx->dev));
    uint64_t* dk = nullptr;
    CU(cudaMalloc(&dk, nkeys * 8));
    cudaError_t e = cudaMemcpyAsync(dk, keys, nkeys * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    int rc = e == cudaSuccess ? build_from_device_keys(ctx, dk) : BSGS_E_CUDA;
    if (e != cudaSuccess) { ctx->last_cuda = (int)e; cudaGetLastError(); }
    cudaFree(dk);
    return rc;
}
// The real baby-step table: key j = low 64 bits of X((j+1)G), j < w (baby(), F197:1162-1235; index base
// F197:1221), produced by the giant-step kernel in emit mode: a temporary array of (i+1)G, i < N,
// and centres C_c = c(2N+1) + N + 1, so that centre c yields the multipliers c(2N+1)+1 .. (c+1)(2N+1).
template <int BLOCK, int MINB>
static cudaError_t launch_emit_k(const bsgs_ctx* c, const gs::Tables& tb, const gs::WinParams& wp, int nwin) {
    gs::k_giant_step<BLOCK, MINB, 1><<<c->grid, BLOCK, 0, c->stream>>>(tb, wp, nwin);
    return cudaGetLastError();
}
static cudaError_t launch_emit(const bsgs_ctx* c, const gs::Tables& tb, const gs::WinParams& wp, int nwin) {
    if (c->block == 320) return launch_emit_k<320, 2>(c, tb, wp, nwin);
    if (c->block == 128) return launch_emit_k<128, 6>(c, tb, wp, nwin);
    if (c->block == 64) return launch_emit_k<64, 12>(c, tb, wp, nwin);
    return c->blocks_per_sm >= 3 ? launch_emit_k<256, 3>(c, tb, wp, nwin) : launch_emit_k<256, 2>(c, tb, wp, nwin);
}
static int generate_baby_keys(bsgs_ctx* ctx, uint64_t* d_keys) {
    size_t slots = (size_t)ctx->NT * ctx->PP;
    const uint32_t W = BSGS_MAX_WINDOWS;
    uint4 *gx = nullptr, *gy = nullptr, *chain = nullptr;
    auto cleanup = [&]() { cudaFree(gx); cudaFree(gy); cudaFree(chain); };
    cudaError_t e = cudaMalloc(&gx, slots * 32);
    if (e == cudaSuccess) e = cudaMalloc(&gy, slots * 32);
    if (e == cudaSuccess) e = cudaMalloc(&chain, slots * 32 * W + 2 * (size_t)ctx->NT * 32);
    if (e != cudaSuccess) { cleanup(); ctx->last_cuda = (int)e; cudaGetLastError(); return BSGS_E_NOMEM; }
    int rc = generate_multiples(ctx, gx, gy, hec::kG);
    if (rc) { cleanup(); return rc; }
    gs::Tables tb = make_tables(ctx);
    tb.gx = gx; tb.gy = gy; tb.chain = chain; tb.keys_out = d_keys; tb.nkeys = ctx->w;
    const uint64_t span = 2 * ctx->N + 1;
    const uint64_t ncenters = (ctx->w + span - 1) / span;
    hec::Pt P, S;
    hec::pt_mul_u64(P, hec::kG, ctx->N + 1);                     // first centre
    hec::pt_mul_u64(S, hec::kG, span);                           // centre stride
    for (uint64_t c0 = 0; c0 < ncenters && rc == BSGS_OK; c0 += W) {
        uint32_t n = (uint32_t)(ncenters - c0 < W ? ncenters - c0 : W);
        gs::WinParams wp; memset(&wp, 0, sizeof wp);
        for (uint32_t i = 0; i < n; i++) {
            uint8_t pxy[64]; hec::store_le(pxy, P.x); hec::store_le(pxy + 32, P.y);
            fill_win(wp.win[i], pxy);
            wp.center[i] = (c0 + i) * span + ctx->N + 1;
            hec::pt_add(P, P, S);
        }
        e = launch_emit(ctx, tb, wp, (int)n);
        if (e != cudaSuccess) { ctx->last_cuda = (int)e; cudaGetLastError(); rc = BSGS_E_CUDA; }
    }
    e = cudaStreamSynchronize(ctx->stream);
    if (rc == BSGS_OK && e != cudaSuccess) { ctx->last_cuda = (int)e; cudaGetLastError(); rc = BSGS_E_CUDA; }
    cleanup();
    return rc;
}
extern "C" int bsgs_ht_generate(bsgs_ctx* ctx) {
    if (!ctx) return BSGS_E_ARG;
    CU(cudaSetDevice(ctx->dev));
    uint64_t* dk = nullptr;
    CU(cudaMalloc(&dk, (size_t)ctx->w * 8));
    int rc = generate_baby_keys(ctx, dk);
    if (rc == BSGS_OK) rc = build_from_device_keys(ctx, dk);
    cudaFree(dk);
    return rc;
}
extern "C" int bsgs_ht_build_synthetic(bsgs_ctx* ctx, uint64_t seed) {
    if (!ctx) return BSGS_E_ARG;
    CU(cudaSetDevice(ctx->dev));
    uint64_t* dk = nullptr;
    CU(cudaMalloc(&dk, (size_t)ctx->w * 8));
    tbl::k_fill_synth<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dk, ctx->w, seed);
    cudaError_t e = cudaGetLastError();
    int rc = e == cudaSuccess ? build_from_device_keys(ctx, dk) : BSGS_E_CUDA;
    if (e != cudaSuccess) { ctx->last_cuda = (int)e; cudaGetLastError(); }
    cudaFree(dk);
    return rc;
}
extern "C" int bsgs_ht_export(bsgs_ctx* ctx, int kind, void* out, size_t bytes) {
    if (!ctx || !out) return BSGS_E_ARG;
    if (!ctx->ht_ready) return BSGS_E_STATE;
    CU(cudaSetDevice(ctx->dev));
    size_t head = ((size_t)ctx->H + 1) * 4;
    if (kind == 0) {                                             // htGPU image: the device buffer as it is
        if (bytes != ht_image_bytes(ctx)) return BSGS_E_ARG;
        CU(cudaMemcpyAsync(out, ctx->d_ht, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return BSGS_OK;
    }
    if (kind != 1 || bytes != head + (size_t)ctx->w * 8) return BSGS_E_ARG;
    if (!ctx->d_js) return BSGS_E_STATE;                         // table was uploaded as htGPU: payloads unknown
    uint2* items = nullptr;
    CU(cudaMalloc(&items, (size_t)ctx->w * 8 + 16));
    tbl::k_interleave<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const uint32_t*)ctx->d_ht + ctx->H + 1, ctx->d_js, items, ctx->w);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, ctx->d_ht, head, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync((uint8_t*)out + head, items, (size_t)ctx->w * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(items);
    CU(e);
    return BSGS_OK;
}
extern "C" int bsgs_ht_device_image(bsgs_ctx* ctx, void** dev_ptr, size_t* bytes) {
    if (!ctx || !dev_ptr || !bytes) return BSGS_E_ARG;
    if (!ctx->ht_ready) return BSGS_E_STATE;
    *dev_ptr = ctx->d_ht; *bytes = ht_image_bytes(ctx);
    return BSGS_OK;
}

// ------------------------------------------------------------------------------------------------ hot path
static gs::Tables make_tables(const bsgs_ctx* c, int lane) {
    gs::Tables tb;
    tb.gx = c->d_gx; tb.gy = c->d_gy; tb.chain = lane ? c->d_chain2 : c->d_chain; tb.buckets = c->d_buckets;
    tb.csr_off = (const uint32_t*)c->d_ht; tb.csr_tags = tb.csr_off + c->H + 1;
    tb.hit_count = lane ? c->d_hits2 : c->d_hits; tb.hits = tb.hit_count + 4; tb.hit_cap = c->hit_cap;
    tb.NT = c->NT; tb.PP = c->PP; tb.N = (uint32_t)c->N; tb.mask = (uint32_t)(c->H - 1);
    tb.exact_thr = c->force_exact ? 0u : FP_KEY_THR;
    tb.keys_out = nullptr; tb.nkeys = 0;
    tb.trace = c->d_trace;
    return tb;
}

static void fill_win(gs::WinConst& wc, const uint8_t* pxy) {
    hec::Fe x, y, nx, y2;
    hec::load_le(x, pxy); hec::load_le(y, pxy + 32);
    hec::neg(nx, x); hec::add(y2, y, y);
    memcpy(wc.px, x.v, 32); memcpy(wc.py, y.v, 32); memcpy(wc.npx, nx.v, 32); memcpy(wc.py2, y2.v, 32);
}

template <int BLOCK, int MINB>
static cudaError_t launch_k(bsgs_ctx* c, uint32_t nwin, const uint8_t* pxy, int lane) {
    Lane L = lane_of(c, lane);
    gs::WinParams& wp = *L.wp;
    memset(&wp, 0, sizeof wp);
    for (uint32_t i = 0; i < nwin; i++) fill_win(wp.win[i], pxy + 64 * i);
    gs::Tables tb = make_tables(c, lane);
    gs::k_giant_step<BLOCK, MINB, 0><<<c->grid, BLOCK, 0, L.st>>>(tb, wp, (int)nwin);
    return cudaGetLastError();
}
static cudaError_t launch_any(bsgs_ctx* c, uint32_t nwin, const uint8_t* pxy, int lane = 0) {
    if (nwin < 1 || nwin > BSGS_MAX_WINDOWS) return cudaErrorInvalidValue;
    if (c->block == 320) return launch_k<320, 2>(c, nwin, pxy, lane);                  // 5 warps per scheduler, 96 registers
    if (c->block == 128) return launch_k<128, 6>(c, nwin, pxy, lane);                  // same 24 warps per SM in smaller CTAs: slots free sooner
    if (c->block == 64) return launch_k<64, 12>(c, nwin, pxy, lane);
    switch (c->blocks_per_sm) {
        case 1: case 2: return launch_k<256, 2>(c, nwin, pxy, lane);
        case 3: return launch_k<256, 3>(c, nwin, pxy, lane);
        default: return cudaErrorInvalidValue;
    }
}

// second lane, allocated on first use (bsgs_run / bsgs_time_launches with more than one launch).  When the
// device has no room for a second chain (very large N) the context quietly stays with one launch at a time.
static int ensure_lane2(bsgs_ctx* ctx) {
    if (ctx->stream2) return BSGS_OK;
    CU(cudaSetDevice(ctx->dev));
    cudaStream_t st = nullptr;
    cudaError_t e = cudaMalloc(&ctx->d_chain2, ctx->chain_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_hits2, 16 + (size_t)ctx->hit_cap * 12);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_verdict2, (size_t)ctx->hit_cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_verdict2, (size_t)ctx->hit_cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_hits2, 16 + (size_t)ctx->hit_cap * 12);
    if (e == cudaSuccess && !ctx->e2) e = cudaEventCreate(&ctx->e2);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_hits2, 0, 16 + (size_t)ctx->hit_cap * 12, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(ctx->d_chain2); cudaFree(ctx->d_hits2); cudaFree(ctx->d_verdict2);
        if (ctx->h_verdict2) cudaFreeHost(ctx->h_verdict2);
        if (ctx->h_hits2) cudaFreeHost(ctx->h_hits2);
        if (st) cudaStreamDestroy(st);
        ctx->d_chain2 = nullptr; ctx->d_hits2 = ctx->d_verdict2 = ctx->h_verdict2 = ctx->h_hits2 = nullptr;
        cudaGetLastError();
        if (e == cudaErrorMemoryAllocation) { ctx->overlap = false; return BSGS_OK; }     // no room: one lane
        ctx->last_cuda = (int)e;
        return BSGS_E_CUDA;
    }
    ctx->stream2 = st;
    return BSGS_OK;
}

// read back counter + records of the launch on `lane`, clear the counter; appends to out[*nhits...] with window offset
static int collect_hits(bsgs_ctx* ctx, uint32_t window_base, bsgs_hit* out, uint32_t cap, uint32_t* nhits, int lane = 0) {
    Lane L = lane_of(ctx, lane);
    const uint32_t eager = 64;                                   // one D2H covers the common case
    CU(cudaMemcpyAsync(L.h_hits, L.d_hits, 16 + eager * 12, cudaMemcpyDeviceToHost, L.st));
    CU(cudaStreamSynchronize(L.st));
    uint32_t cnt = L.h_hits[0];
    if (cnt == 0) return BSGS_OK;
    uint32_t stored = cnt < ctx->hit_cap ? cnt : ctx->hit_cap;
    if (stored > eager) {
        CU(cudaMemcpyAsync(L.h_hits + 4 + eager * 3, L.d_hits + 4 + eager * 3, (size_t)(stored - eager) * 12, cudaMemcpyDeviceToHost, L.st));
    }
    const bool verify = ctx->verify && ctx->d_js;
    if (verify) {                                                // re-examine every record where it was produced
        gs::Tables tb = make_tables(ctx, lane);
        vfy::k_verify_hits<<<(stored + 31) / 32, 32, 0, L.st>>>(tb, *L.wp, ctx->d_js, stored, L.d_verdict);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(L.h_verdict, L.d_verdict, (size_t)stored * 4, cudaMemcpyDeviceToHost, L.st));
    }
    CU(cudaMemsetAsync(L.d_hits, 0, 4, L.st));                   // F197:2500-2504
    CU(cudaStreamSynchronize(L.st));
    int rc = cnt > ctx->hit_cap ? BSGS_E_OVERFLOW : BSGS_OK;
    for (uint32_t i = 0; i < stored; i++) {
        uint32_t payload = 0xFFFFFFFFu;
        if (verify) {
            payload = L.h_verdict[i];
            if (payload == 0xFFFFFFFFu) continue;                // false positive of the 32-bit-tag table: dropped
        }
        uint32_t pos = *nhits;
        if (pos < cap && out) {
            out[pos].code = L.h_hits[4 + 3 * i]; out[pos].idx = L.h_hits[4 + 3 * i + 1];
            out[pos].window = window_base + L.h_hits[4 + 3 * i + 2]; out[pos].reserved = payload;
        } else rc = BSGS_E_OVERFLOW;
        (*nhits)++;
    }
    return rc;
}

extern "C" int bsgs_step_multi(bsgs_ctx* ctx, uint32_t nwin, const uint8_t* pxy_le, bsgs_hit* out, uint32_t cap, uint32_t* nhits) {
    if (!ctx || !pxy_le || !nhits || nwin < 1 || nwin > ctx->max_windows) return BSGS_E_ARG;
    if (!ctx->giants_ready || !ctx->ht_ready) return BSGS_E_STATE;
    *nhits = 0;
    CU(cudaSetDevice(ctx->dev));
    CU(launch_any(ctx, nwin, pxy_le));
    return collect_hits(ctx, 0, out, cap, nhits);
}
extern "C" int bsgs_step(bsgs_ctx* ctx, const uint8_t px_le[32], const uint8_t py_le[32], bsgs_hit* out, uint32_t cap, uint32_t* nhits) {
    if (!px_le || !py_le) return BSGS_E_ARG;
    uint8_t pxy[64]; memcpy(pxy, px_le, 32); memcpy(pxy + 32, py_le, 32);
    return bsgs_step_multi(ctx, 1, pxy, out, cap, nhits);
}

extern "C" int bsgs_run(bsgs_ctx* ctx, const uint8_t p0xy_le[64], const uint8_t dxy_le[64], uint64_t nwindows, uint32_t batch,
                        uint32_t stop_after_hits, bsgs_hit* out, uint32_t cap, uint32_t* nhits, uint64_t* windows_done) {
    if (!ctx || !p0xy_le || !dxy_le || !nhits) return BSGS_E_ARG;
    if (!ctx->giants_ready || !ctx->ht_ready) return BSGS_E_STATE;
    if (batch < 1 || batch > ctx->max_windows) return BSGS_E_ARG;
    *nhits = 0;
    CU(cudaSetDevice(ctx->dev));
    hec::Pt P, D;
    hec::load_le(P.x, p0xy_le); hec::load_le(P.y, p0xy_le + 32);
    hec::load_le(D.x, dxy_le); hec::load_le(D.y, dxy_le + 32);
    uint64_t done = 0; int rc = BSGS_OK;
    uint8_t cur[BSGS_MAX_WINDOWS * 64], nxt[BSGS_MAX_WINDOWS * 64];
    auto fill = [&](uint8_t* buf, uint32_t n) {                 // GetJob (F197:2077-2092), n windows ahead
        for (uint32_t i = 0; i < n; i++) { hec::store_le(buf + 64 * i, P.x); hec::store_le(buf + 64 * i + 32, P.y); hec::pt_add(P, P, D); }
    };
    bool two = ctx->overlap && nwindows > batch;                   // two launches in flight, alternating lanes
    if (two) { int r = ensure_lane2(ctx); if (r) return r; two = ctx->overlap; }
    uint32_t n = (uint32_t)(nwindows < batch ? nwindows : batch);
    uint64_t issued = 0; int lane = 0;
    if (n) { fill(cur, n); CU(launch_any(ctx, n, cur, 0)); issued = n; }
    while (done < nwindows) {
        // the next batch: base points derived on the host while the kernel runs; with two lanes it is launched BEFORE
        // the records of the current one are collected
        uint64_t left = nwindows - issued;
        uint32_t n2 = (uint32_t)(left < batch ? left : batch);
        if (n2) {
            fill(nxt, n2);
            if (two) { CU(launch_any(ctx, n2, nxt, lane ^ 1)); issued += n2; }
        }
        int r2 = collect_hits(ctx, (uint32_t)done, out, cap, nhits, lane);
        done += n;
        bool stop = (r2 && r2 != BSGS_E_OVERFLOW) || (stop_after_hits && *nhits >= stop_after_hits);
        if (r2 == BSGS_E_OVERFLOW) rc = r2; else if (r2) rc = r2;
        if (stop) {
            if (two && n2) {                                        // one more launch is already running: its windows count too
                int r3 = collect_hits(ctx, (uint32_t)done, out, cap, nhits, lane ^ 1);
                done += n2;
                if (r3 && rc == BSGS_OK) rc = r3;
            }
            break;
        }
        if (!n2) break;
        if (two) lane ^= 1;
        else { CU(launch_any(ctx, n2, nxt, 0)); issued += n2; }
        n = n2;
    }
    if (windows_done) *windows_done = done;
    return rc;
}

// ------------------------------------------------------------------------------------------------ measurement
extern "C" int bsgs_time_launches(bsgs_ctx* ctx, uint32_t nwin, const uint8_t* pxy_le, uint32_t iters, float* ms) {
    if (!ctx || !pxy_le || !ms || nwin < 1 || nwin > ctx->max_windows || !iters) return BSGS_E_ARG;
    if (!ctx->giants_ready || !ctx->ht_ready) return BSGS_E_STATE;
    CU(cudaSetDevice(ctx->dev));
    bool two = ctx->overlap && iters > 1;
    if (two) { int r = ensure_lane2(ctx); if (r) return r; two = ctx->overlap; }
    CU(cudaEventRecord(ctx->e0, ctx->stream));
    if (two) CU(cudaStreamWaitEvent(ctx->stream2, ctx->e0, 0));      // nothing of the timed region starts before e0
    for (uint32_t i = 0; i < iters; i++) CU(launch_any(ctx, nwin, pxy_le + (size_t)i * nwin * 64, two ? (int)(i & 1) : 0));
    if (two) { CU(cudaEventRecord(ctx->e2, ctx->stream2)); CU(cudaStreamWaitEvent(ctx->stream, ctx->e2, 0)); }    // e1 after both lanes
    CU(cudaEventRecord(ctx->e1, ctx->stream));
    CU(cudaEventSynchronize(ctx->e1));
    CU(cudaEventElapsedTime(ms, ctx->e0, ctx->e1));
    CU(cudaMemsetAsync(ctx->d_hits, 0, 4, ctx->stream));
    if (two) { CU(cudaMemsetAsync(ctx->d_hits2, 0, 4, ctx->stream2)); CU(cudaStreamSynchronize(ctx->stream2)); }
    CU(cudaStreamSynchronize(ctx->stream));
    return BSGS_OK;
}

extern "C" int bsgs_probe_bench(bsgs_ctx* ctx, uint64_t nkeys, uint64_t seed, uint32_t iters, float* ms, uint64_t* nfound) {
    if (!ctx || !ms || !iters) return BSGS_E_ARG;
    if (!ctx->ht_ready) return BSGS_E_STATE;
    CU(cudaSetDevice(ctx->dev));
    gs::Tables tb = make_tables(ctx);
    CU(cudaMemsetAsync(ctx->d_scratch, 0, 8, ctx->stream));
    const char* env_u = getenv("BSGS_PROBE_UNROLL"); const char* env_g = getenv("BSGS_PROBE_CTAS");   // measurement knobs
    const int unroll = env_u ? atoi(env_u) : 8;
    uint32_t grid = (uint32_t)ctx->sm_count * (env_g ? (uint32_t)atoi(env_g) : 16u);
    CU(cudaEventRecord(ctx->e0, ctx->stream));
    for (uint32_t i = 0; i < iters; i++) {
        if (unroll >= 8) gs::k_probe_bench_fp<8><<<grid, 256, 0, ctx->stream>>>(tb, nkeys, seed + i * nkeys, ctx->d_scratch);
        else if (unroll <= 2) gs::k_probe_bench_fp<2><<<grid, 256, 0, ctx->stream>>>(tb, nkeys, seed + i * nkeys, ctx->d_scratch);
        else gs::k_probe_bench_fp<4><<<grid, 256, 0, ctx->stream>>>(tb, nkeys, seed + i * nkeys, ctx->d_scratch);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->e1, ctx->stream));
    CU(cudaEventSynchronize(ctx->e1));
    CU(cudaEventElapsedTime(ms, ctx->e0, ctx->e1));
    unsigned long long f = 0;
    CU(cudaMemcpyAsync(&f, ctx->d_scratch, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nfound) *nfound = f;
    return BSGS_OK;
}

extern "C" int bsgs_probe_keys(bsgs_ctx* ctx, const uint64_t* keys, uint64_t nkeys, uint8_t* found) {
    if (!ctx || !keys || !found) return BSGS_E_ARG;
    if (!ctx->ht_ready) return BSGS_E_STATE;
    if (!nkeys) return BSGS_OK;
    CU(cudaSetDevice(ctx->dev));
    uint64_t* dk = nullptr; uint8_t* df = nullptr;
    CU(cudaMalloc(&dk, nkeys * 8));
    cudaError_t e = cudaMalloc(&df, nkeys);
    if (e != cudaSuccess) { cudaFree(dk); ctx->last_cuda = e; cudaGetLastError(); return BSGS_E_NOMEM; }
    gs::Tables tb = make_tables(ctx);
    e = cudaMemcpyAsync(dk, keys, nkeys * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        uint32_t blocks = (uint32_t)((nkeys + 255) / 256);
        gs::k_probe_keys_fp<<<blocks, 256, 0, ctx->stream>>>(tb, dk, nkeys, df);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(found, df, nkeys, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dk); cudaFree(df);
    if (e != cudaSuccess) { ctx->last_cuda = e; cudaGetLastError(); return BSGS_E_CUDA; }
    return BSGS_OK;
}

// ------------------------------------------------------------------------------------------------ diagnostics
extern "C" int bsgs_debug_field(int device, int op, const uint32_t* a, const uint32_t* b, uint32_t* r, uint32_t n) {
    if (!a || !b || !r || !n) return BSGS_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BSGS_E_CUDA;
    uint32_t *da = nullptr, *db = nullptr, *dr = nullptr; size_t bytes = (size_t)n * 32;
    cudaError_t e = cudaMalloc(&da, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&db, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dr, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { gs::k_debug_field<<<(n + 127) / 128, 128>>>(op, da, db, dr, n); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(r, dr, bytes, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dr);
    if (e != cudaSuccess) { cudaGetLastError(); return BSGS_E_CUDA; }
    return BSGS_OK;
}

// the hot loop's key extraction: keys[i] = low 64 bits of canonical (in[i]^2 + A[i] - Q[i]) mod p (mode 0, in[i] =
// lambda, 8 limbs) or of (in[i] + A[i] - Q[i]) mod p (mode 1, in[i] = an arbitrary 512-bit value, 16 limbs);
// exact[i] = 1 when the classification asked for the out-of-line exact path; force != 0 sends everything there
extern "C" int bsgs_debug_key(int device, int mode, const uint32_t* in, const uint32_t* A, const uint32_t* Q, uint64_t* keys, uint8_t* exact, uint32_t n, int force) {
    if (!in || !A || !Q || !keys || !exact || !n || mode < 0 || mode > 1) return BSGS_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BSGS_E_CUDA;
    uint32_t *dl = nullptr, *da = nullptr, *dq = nullptr; uint64_t* dk = nullptr; uint8_t* de = nullptr; size_t bytes = (size_t)n * 32;
    const size_t inbytes = bytes * (mode ? 2 : 1);
    cudaError_t e = cudaMalloc(&dl, inbytes);
    if (e == cudaSuccess) e = cudaMalloc(&da, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dq, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dk, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&de, n);
    if (e == cudaSuccess) e = cudaMemcpy(dl, in, inbytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(da, A, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dq, Q, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { gs::k_debug_key<<<(n + 127) / 128, 128>>>(mode, dl, da, dq, dk, de, n, force ? 0u : FP_KEY_THR); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(keys, dk, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(exact, de, n, cudaMemcpyDeviceToHost);
    cudaFree(dl); cudaFree(da); cudaFree(dq); cudaFree(dk); cudaFree(de);
    if (e != cudaSuccess) { cudaGetLastError(); return BSGS_E_CUDA; }
    return BSGS_OK;
}

// diagnostics: per-warp {SM id, globaltimer ns at kernel start, after pass 1, at the end} of the last launch
// (option "trace" = 1 before the launch); out holds 4 words per warp, *nwarps warps
extern "C" int bsgs_debug_trace(bsgs_ctx* ctx, uint64_t* out, uint32_t cap_warps, uint32_t* nwarps) {
    if (!ctx || !out || !nwarps) return BSGS_E_ARG;
    if (!ctx->d_trace) return BSGS_E_STATE;
    uint32_t n = ctx->NT / 32; *nwarps = n;
    if (cap_warps < n) return BSGS_E_ARG;
    CU(cudaSetDevice(ctx->dev));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, ctx->d_trace, (size_t)n * 32, cudaMemcpyDeviceToHost));
    return BSGS_OK;
}
