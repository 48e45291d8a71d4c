/*
 * ref_harness.c -- runs the REFERENCE's own kernel (TEST INFRASTRUCTURE ONLY).
 *
 * The reference's GPU path is the PTX kernel `_test1` embedded in 1_9_7File.pb:5181-23979; the
 * host that drives it is PureBasic (not buildable here).  oracle/Makefile decodes that PTX where
 * it lies and assembles it with ptxas for sm_100a into oracle/_ref/ref_test1.cubin.  This file is
 * our own ~200-line driver-API runner that reproduces what the reference host `cuda()` does
 * (1_9_7File.pb:2181-2509):
 *     one allocation, 64-byte aligned; bytes 0..2047 header (u32 hit counter at 0, records
 *     {code,idx} from byte 128); giants X|Y at +2048 (N*64 bytes); chain block (N*32);
 *     htGPU image at puboffset = align64(N*96+64)+2048; const `_A` (30 x u32) filled as
 *     1_9_7File.pb:2312-2325; per launch 64 bytes (Px,Py as 8 MSW-first words) to _A+32,
 *     launch (B,1,1)x(T,1,1) with the base pointer, sync, read counter + records, clear counter.
 * It is the bit-exact GPU oracle for the hit records and the "reference on the same box" timing.
 * libcuda is resolved with dlopen so this library loads on machines without a GPU.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <dlfcn.h>

typedef int CUresult; typedef int CUdevice; typedef void *CUcontext, *CUmodule, *CUfunction, *CUevent, *CUstream;
typedef unsigned long long CUdeviceptr;

static struct {
    void *h;
    CUresult (*Init)(unsigned);
    CUresult (*DeviceGet)(CUdevice *, int);
    CUresult (*PrimaryRetain)(CUcontext *, CUdevice);
    CUresult (*PrimaryRelease)(CUdevice);
    CUresult (*CtxPush)(CUcontext);
    CUresult (*CtxPop)(CUcontext *);
    CUresult (*ModuleLoadData)(CUmodule *, const void *);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
    CUresult (*ModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *);
    CUresult (*FuncSetCacheConfig)(CUfunction, int);
    CUresult (*MemAlloc)(CUdeviceptr *, size_t);
    CUresult (*MemFree)(CUdeviceptr);
    CUresult (*MemcpyHtoD)(CUdeviceptr, const void *, size_t);
    CUresult (*MemcpyDtoH)(void *, CUdeviceptr, size_t);
    CUresult (*MemsetD8)(CUdeviceptr, unsigned char, size_t);
    CUresult (*MemcpyDtoD)(CUdeviceptr, CUdeviceptr, size_t);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
    CUresult (*CtxSynchronize)(void);
    CUresult (*EventCreate)(CUevent *, unsigned);
    CUresult (*EventRecord)(CUevent, CUstream);
    CUresult (*EventSynchronize)(CUevent);
    CUresult (*EventElapsedTime)(float *, CUevent, CUevent);
    CUresult (*EventDestroy)(CUevent);
} cu;

static int load_cuda(void) {
    if (cu.h) return 0;
    cu.h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!cu.h) cu.h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!cu.h) return -1;
#define SYM(field, name) do { *(void **)(&cu.field) = dlsym(cu.h, name); if (!cu.field) { fprintf(stderr, "ref_harness: missing %s\n", name); return -2; } } while (0)
    SYM(Init, "cuInit"); SYM(DeviceGet, "cuDeviceGet");
    SYM(PrimaryRetain, "cuDevicePrimaryCtxRetain"); SYM(PrimaryRelease, "cuDevicePrimaryCtxRelease_v2");
    SYM(CtxPush, "cuCtxPushCurrent_v2"); SYM(CtxPop, "cuCtxPopCurrent_v2");
    SYM(ModuleLoadData, "cuModuleLoadData"); SYM(ModuleUnload, "cuModuleUnload");
    SYM(ModuleGetFunction, "cuModuleGetFunction"); SYM(ModuleGetGlobal, "cuModuleGetGlobal_v2");
    SYM(FuncSetCacheConfig, "cuFuncSetCacheConfig");
    SYM(MemAlloc, "cuMemAlloc_v2"); SYM(MemFree, "cuMemFree_v2");
    SYM(MemcpyHtoD, "cuMemcpyHtoD_v2"); SYM(MemcpyDtoH, "cuMemcpyDtoH_v2"); SYM(MemsetD8, "cuMemsetD8_v2"); SYM(MemcpyDtoD, "cuMemcpyDtoD_v2");
    SYM(LaunchKernel, "cuLaunchKernel"); SYM(CtxSynchronize, "cuCtxSynchronize");
    SYM(EventCreate, "cuEventCreate"); SYM(EventRecord, "cuEventRecord");
    SYM(EventSynchronize, "cuEventSynchronize"); SYM(EventElapsedTime, "cuEventElapsedTime");
    SYM(EventDestroy, "cuEventDestroy_v2");
#undef SYM
    return 0;
}

typedef struct ref_ctx {
    CUdevice dev; CUcontext ctx; CUmodule mod; CUfunction fn;
    CUdeviceptr raw, base, constA;
    uint32_t T, B, p, htsz; uint64_t w, N, H, puboffset; size_t total;
    CUevent e0, e1;
    int last_err;
} ref_ctx;

#define CK(call) do { CUresult _e = (call); if (_e) { c->last_err = _e; fprintf(stderr, "ref_harness: %s -> %d\n", #call, _e); rc = -100 - (int)_e; goto out; } } while (0)
#define ALIGN 64ull         /* #alignMemoryGpu, 1_9_7File.pb:15 */
#define PARAMSIZE 2048ull   /* paramsize, 1_9_7File.pb:2211 */

int ref_available(void) { return load_cuda() == 0; }

int ref_create(ref_ctx **out, int dev, const char *cubin_path, uint32_t T, uint32_t B, uint32_t p, uint64_t w, uint32_t htsz) {
    int rc = 0; *out = NULL;
    if (load_cuda()) return -1;
    ref_ctx *c = (ref_ctx *)calloc(1, sizeof *c);
    c->T = T; c->B = B; c->p = p; c->w = w; c->htsz = htsz; c->N = (uint64_t)T * B * p; c->H = 1ull << htsz;
    if (c->N * 8 >= (1ull << 31)) { free(c); return -3; }     /* kernel indexes with signed 32-bit (PTX197:L153-156) */
    FILE *f = fopen(cubin_path, "rb"); if (!f) { free(c); return -2; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    char *img = (char *)malloc(sz + 1); if (fread(img, 1, sz, f) != (size_t)sz) { fclose(f); free(img); free(c); return -2; } img[sz] = 0; fclose(f);
    CK(cu.Init(0)); CK(cu.DeviceGet(&c->dev, dev));
    CK(cu.PrimaryRetain(&c->ctx, c->dev)); CK(cu.CtxPush(c->ctx));
    CK(cu.ModuleLoadData(&c->mod, img));
    CK(cu.ModuleGetFunction(&c->fn, c->mod, "_test1"));
    cu.FuncSetCacheConfig(c->fn, 2);                          /* prefer L1, 1_9_7File.pb:2203 */
    /* memory map, 1_9_7File.pb:2208-2216 */
    uint64_t need = c->N * 96 + ALIGN;
    c->puboffset = need + ALIGN - (need % ALIGN) + PARAMSIZE;
    need += PARAMSIZE;
    uint64_t pubneed = c->H * 4 + w * 4 + 4;
    c->total = need + pubneed + PARAMSIZE + ALIGN;
    if (c->puboffset + pubneed > c->total) c->total = c->puboffset + pubneed + ALIGN;
    CK(cu.MemAlloc(&c->raw, c->total + ALIGN));
    c->base = c->raw + ALIGN - (c->raw % ALIGN);
    size_t csz = 0; CK(cu.ModuleGetGlobal(&c->constA, &csz, c->mod, "_A"));
    uint8_t b[PARAMSIZE]; memset(b, 0, sizeof b);
    uint32_t u;
    u = (uint32_t)w;          memcpy(b + 4, &u, 4);            /* 1_9_7File.pb:2312-2323 */
    u = p;                    memcpy(b + 8, &u, 4);
    u = (uint32_t)c->N;       memcpy(b + 12, &u, 4);
    memcpy(b + 96, &c->puboffset, 8);
    u = (uint32_t)(c->H + 1); memcpy(b + 104, &u, 4);
    u = (uint32_t)(c->H - 1); memcpy(b + 112, &u, 4);
    CK(cu.MemcpyHtoD(c->constA, b, 120));
    CK(cu.MemcpyHtoD(c->base, b, PARAMSIZE));
    CK(cu.EventCreate(&c->e0, 0)); CK(cu.EventCreate(&c->e1, 0));
out:
    free(img);
    { CUcontext t; cu.CtxPop(&t); }
    if (rc) { free(c); return rc; }
    *out = c; return 0;
}

int ref_upload_giants(ref_ctx *c, const void *g2, size_t bytes) {
    int rc = 0; if (bytes != c->N * 64) return -4;
    cu.CtxPush(c->ctx);
    CK(cu.MemcpyHtoD(c->base + PARAMSIZE, g2, bytes));         /* 1_9_7File.pb:2337 */
out:
    { CUcontext t; cu.CtxPop(&t); } return rc;
}
int ref_upload_ht(ref_ctx *c, const void *htgpu, size_t bytes) {
    int rc = 0; if (bytes != c->H * 4 + 4 + c->w * 4) return -4;
    cu.CtxPush(c->ctx);
    CK(cu.MemcpyHtoD(c->base + c->puboffset, htgpu, bytes));   /* 1_9_7File.pb:2350 */
out:
    { CUcontext t; cu.CtxPop(&t); } return rc;
}

/* same image, already resident in device memory of the same (primary) context: device-to-device copy */
int ref_upload_ht_dev(ref_ctx *c, unsigned long long dev_src, size_t bytes) {
    int rc = 0; if (bytes != c->H * 4 + 4 + c->w * 4) return -4;
    cu.CtxPush(c->ctx);
    CK(cu.MemcpyDtoD(c->base + c->puboffset, (CUdeviceptr)dev_src, bytes));
    CK(cu.CtxSynchronize());
out:
    { CUcontext t; cu.CtxPop(&t); } return rc;
}

/* timing only: `iters` back-to-back launches on the same base point, CUDA events around the train;
 * the hit counter is cleared afterwards (records beyond 240 would run into the giants: callers use
 * tables that yield few hits). */
int ref_time_launches(ref_ctx *c, const uint8_t *px_le, const uint8_t *py_le, uint32_t iters, float *ms) {
    int rc = 0; uint32_t words[16];
    for (int i = 0; i < 8; i++) { memcpy(&words[i], px_le + 4 * (7 - i), 4); memcpy(&words[8 + i], py_le + 4 * (7 - i), 4); }
    cu.CtxPush(c->ctx);
    CK(cu.MemcpyHtoD(c->constA + 32, words, 64));
    void *params[1] = {&c->base};
    CK(cu.EventRecord(c->e0, 0));
    for (uint32_t i = 0; i < iters; i++) CK(cu.LaunchKernel(c->fn, c->B, 1, 1, c->T, 1, 1, 0, 0, params, NULL));
    CK(cu.EventRecord(c->e1, 0));
    CK(cu.CtxSynchronize());
    if (ms) cu.EventElapsedTime(ms, c->e0, c->e1);
    { uint32_t z = 0; CK(cu.MemcpyHtoD(c->base, &z, 4)); }
out:
    { CUcontext t; cu.CtxPop(&t); } return rc;
}

/* Px,Py: 32-byte little-endian; recs: cap pairs of {code,idx}; returns raw counter in *nhits */
int ref_launch(ref_ctx *c, const uint8_t *px_le, const uint8_t *py_le, uint32_t *recs, uint32_t cap, uint32_t *nhits, float *ms) {
    int rc = 0; uint32_t words[16];
    for (int i = 0; i < 8; i++) {                              /* swap32: word 0 = most significant, 1_9_7File.pb:2438-2439 */
        memcpy(&words[i], px_le + 4 * (7 - i), 4);
        memcpy(&words[8 + i], py_le + 4 * (7 - i), 4);
    }
    cu.CtxPush(c->ctx);
    CK(cu.MemcpyHtoD(c->constA + 32, words, 64));
    void *params[1] = {&c->base};
    CK(cu.EventRecord(c->e0, 0));
    CK(cu.LaunchKernel(c->fn, c->B, 1, 1, c->T, 1, 1, 0, 0, params, NULL));
    CK(cu.EventRecord(c->e1, 0));
    CK(cu.CtxSynchronize());
    if (ms) cu.EventElapsedTime(ms, c->e0, c->e1);
    uint32_t cnt = 0; CK(cu.MemcpyDtoH(&cnt, c->base, 4));
    if (nhits) *nhits = cnt;
    if (cnt) {
        uint32_t n = cnt > 240 ? 240 : cnt;                    /* record area ends where the giants begin */
        if (n > cap) n = cap;
        if (n) CK(cu.MemcpyDtoH(recs, c->base + 128, (size_t)n * 8));
        uint32_t z = 0; CK(cu.MemcpyHtoD(c->base, &z, 4));      /* 1_9_7File.pb:2500-2504 */
    }
out:
    { CUcontext t; cu.CtxPop(&t); } return rc;
}

void ref_destroy(ref_ctx *c) {
    if (!c) return;
    cu.CtxPush(c->ctx);
    if (c->e0) cu.EventDestroy(c->e0);
    if (c->e1) cu.EventDestroy(c->e1);
    if (c->raw) cu.MemFree(c->raw);
    if (c->mod) cu.ModuleUnload(c->mod);
    { CUcontext t; cu.CtxPop(&t); }
    cu.PrimaryRelease(c->dev);
    free(c);
}
