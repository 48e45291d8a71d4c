/*
 * bsgs_b200.h -- C ABI of libbsgs_b200.so, the B200-native drop-in for the GPU giant-step path of
 * Etayson/BSGS-cuda.
 *
 * What this boundary replaces in the reference (1_9_7File.pb, "F197"):
 *   - the CUDA *driver API* import table and the embedded PTX module  F197:55-106, 5181-23979
 *   - the per-GPU worker's device set-up                              F197:2181-2353  (cuda())
 *   - the per-launch sequence: 64 B H2D of the base point, launch of `_test1`, sync, read-back of
 *     the hit counter and of the {code, idx} records                   F197:2435-2509
 * The reference host keeps: CLI parsing, table files, GetJob (F197:2077-2092), checkerThread
 * (F197:3933-4296).  A PureBasic host binds these symbols with `ImportC "libbsgs_b200.so"` instead of
 * `Import "lib\cuda.lib"` -- see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; all big integers are 32-byte LITTLE-ENDIAN, exactly the
 *     in-memory form of the reference host (lib/Curve64.pb:450-473, m_sethex32), i.e. the bytes the
 *     reference hands to swap32() before its H2D copy (F197:2435-2439);
 *   - every function returns 0 on success or a negative bsgs_status; nothing aborts, prompts or
 *     prints (the reference's exit() blocks on Input(), F197:754-764);
 *   - a context is bound to one device and must be driven by one thread at a time; distinct
 *     contexts may be driven concurrently from distinct threads (one per GPU, like F197:4776);
 *   - the library copies on upload: the caller may free host buffers immediately
 *     (the reference frees them from main after all GPU threads copied, F197:4822-4843).
 */
#ifndef BSGS_B200_H
#define BSGS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsgs_ctx bsgs_ctx;

/* One hit record.  `code` and `idx` have the reference's meaning (F197:2475-2490):
 *   1: X(P + Q_idx) probes true      2: X(P - Q_idx) probes true
 *   4: P.x == Q_idx.x and X(2P) probes true      5: P.x itself probes true (idx = 0)
 * `window` is the index of the base point inside a multi-window call (0 for bsgs_step). */
typedef struct { uint32_t code, idx, window, reserved; } bsgs_hit;
/* `reserved`: 0xFFFFFFFF unless the "verify" option is on.  With it, every record is re-examined on
 * the device (full 256-bit X of P +- Q_idx against the recomputed baby point (j+1)G -- the work of
 * checkerThread, F197:3933-4296): false positives of the 32-bit-tag table are dropped, surviving
 * records carry the baby-step payload j (code 4: 0xFFFFFFFE, no table entry involved). */

typedef enum {
    BSGS_OK = 0,
    BSGS_E_ARG = -1,          /* bad argument / size mismatch */
    BSGS_E_CUDA = -2,         /* CUDA runtime error, see bsgs_last_cuda_error() */
    BSGS_E_NOMEM = -3,        /* device or host allocation failed */
    BSGS_E_STATE = -4,        /* tables not uploaded / committed yet */
    BSGS_E_LIMIT = -5,        /* parameter outside structural limits (N >= 2^32, htsz > 31, w >= 2^32) */
    BSGS_E_OVERFLOW = -6      /* more hits than `cap`; *nhits holds the true count, `cap` were stored */
} bsgs_status;

/* -- lifetime ------------------------------------------------------------------------------ */
/* T,B,p,w,htsz are the reference's -t -b -p -w -htsz (w already parsed to an item count).
 * Replaces cuInit/cuCtxCreate/cuModuleLoadData/cuMemAlloc (F197:2181-2255). */
int  bsgs_create(bsgs_ctx **out, int device, uint32_t T, uint32_t B, uint32_t p, uint64_t w, uint32_t htsz);
void bsgs_destroy(bsgs_ctx *ctx);
const char *bsgs_strerror(int status);
int  bsgs_last_cuda_error(const bsgs_ctx *ctx);            /* cudaError_t of the last failure */

/* -- tables -------------------------------------------------------------------------------- */
/* `g2` = the bytes of <T>_<B>_<p>_<w>_g2.BIN as they are (N*64 bytes; F197:1905-2058).
 * Replaces cuMemcpyHtoD of *GiantArrPacked (F197:2337). */
int  bsgs_upload_giants(bsgs_ctx *ctx, const void *g2, size_t bytes);
/* Device-side generation of the same array, Q_k = (k+1)*(-(2w)G) (replaces giant(0), F197:1418-1488,
 * which is single-threaded on the CPU), and its export as the G2 file image (N*64 bytes). */
int  bsgs_giants_generate(bsgs_ctx *ctx);
int  bsgs_giants_export(bsgs_ctx *ctx, void *g2, size_t bytes);
/* `htgpu` = the bytes of <Gx>_<w>_<H>_htGPUv0.BIN as they are (4H + 4 + 4w bytes; F197:3337-3444).
 * Replaces cuMemcpyHtoD of *GpuHT (F197:2350). */
int  bsgs_upload_ht(bsgs_ctx *ctx, const void *htgpu, size_t bytes);
/* `htcpu` = the bytes of <Gx>_<w>_<H>_htCPUv0.BIN (4H + 4 + 8w bytes; F197:3232-3335): the same table
 * together with its payloads j, which device-side verification ("verify" option) needs. */
int  bsgs_upload_htcpu(bsgs_ctx *ctx, const void *htcpu, size_t bytes);
/* Two-phase variants for multi-GPU start-up: obtain the device staging buffer (so that an NCCL
 * broadcast can land the image in place), then commit (build the in-HBM layouts). */
int  bsgs_giants_staging(bsgs_ctx *ctx, void **dev_ptr, size_t *bytes);
int  bsgs_giants_commit(bsgs_ctx *ctx);
int  bsgs_ht_staging(bsgs_ctx *ctx, void **dev_ptr, size_t *bytes);
int  bsgs_ht_commit(bsgs_ctx *ctx);
/* Single-process form of the same start-up step (the reference runs one host thread per GPU in one
 * process and uploads the table to every GPU from host memory, 1_9_7File.pb:2337-2350, 4776):
 * copy the table of ctxs[0] -- the htGPU image and, when present, the payloads used by device-side
 * verification -- to ctxs[1..n-1] device to device (NVLink peer copies, all in flight at once) and
 * commit it there.  All contexts must have the same w and htsz; ctxs[0] must hold a table. */
int  bsgs_bcast_ht(bsgs_ctx *const *ctxs, int n);

/* Device-side construction of the same table image (replaces HashTableInsert + sortHashTable32bit +
 * packHTFile/packHTGPUFile, F197:2555-2622, 2771-2820, 3232-3444): counting sort of `nkeys` (= w)
 * 64-bit keys, key j being the low 8 bytes of X((j+1)G) (F197:1127, 1221).  The resulting htGPU /
 * htCPU images are byte-identical to the files the reference writes and can be exported. */
int  bsgs_ht_build_from_keys(bsgs_ctx *ctx, const uint64_t *keys, uint64_t nkeys);
/* The whole baby-step table on the device: the w keys X((j+1)G) mod 2^64 are produced by the
 * giant-step kernel itself (replaces GenBabys/baby(), F197:1162-1328: CountCPUs host threads, hours
 * at w >= 2^30), then sorted into the CSR image as above. */
int  bsgs_ht_generate(bsgs_ctx *ctx);
/* measurement only: w uniformly random keys (splitmix64 of seed + j) instead of real baby steps */
int  bsgs_ht_build_synthetic(bsgs_ctx *ctx, uint64_t seed);
/* kind 0: htGPU file image (4H + 4 + 4w bytes); kind 1: htCPU file image (4H + 4 + 8w bytes; only
 * after a device-side build, which knows the payloads j) */
int  bsgs_ht_export(bsgs_ctx *ctx, int kind, void *out, size_t bytes);
/* device address of the resident htGPU image (CSR form), e.g. as the source of a broadcast */
int  bsgs_ht_device_image(bsgs_ctx *ctx, void **dev_ptr, size_t *bytes);

/* -- the hot path -------------------------------------------------------------------------- */
/* One reference launch: same base point, same (code, idx) multiset as `_test1` (order is not part
 * of the contract).  Replaces F197:2435-2509.  `cap` bounds the record buffer explicitly (the
 * reference overflows into the giant array at record 240, F197:2337 / PTX197:L104-109). */
int  bsgs_step(bsgs_ctx *ctx, const uint8_t px_le[32], const uint8_t py_le[32],
               bsgs_hit *out, uint32_t cap, uint32_t *nhits);
/* nwin (1..BSGS_MAX_WINDOWS) GetJob() windows in ONE launch: pxy_le holds nwin records of 64 bytes
 * (x then y).  Each giant point is loaded once for all windows and one field inversion per thread
 * is shared by all of them.  hits carry the window index. */
#define BSGS_MAX_WINDOWS 16
int  bsgs_step_multi(bsgs_ctx *ctx, uint32_t nwin, const uint8_t *pxy_le,
                     bsgs_hit *out, uint32_t cap, uint32_t *nhits);
/* Whole search loop on the library side: windows j = 0..nwindows-1 with base points
 * P_j = P0 + j*D (D = PUBADDBIG, F197:4759-4765), `batch` windows per launch, stops early after the
 * launch in which `stop_after_hits` records have been collected (0 = never).  hit.window = j (the low
 * 32 bits: callers pass at most 2^32 windows per call).  When nwindows > batch, TWO launches are kept in
 * flight on two streams with their own chain and record buffers (the warps of a launch finish unevenly;
 * the CTAs of the next launch take the freed slots): on an early stop the windows of the launch that was
 * already running are searched, reported and counted in *windows_done as well.  Option "overlap" = 0
 * restores one launch at a time.  The second chain buffer (as large as the first) is allocated on first use. */
int  bsgs_run(bsgs_ctx *ctx, const uint8_t p0xy_le[64], const uint8_t dxy_le[64], uint64_t nwindows,
              uint32_t batch, uint32_t stop_after_hits, bsgs_hit *out, uint32_t cap, uint32_t *nhits,
              uint64_t *windows_done);

/* -- measurement helpers (used by bench.py; device-resident timing with CUDA events) -------- */
/* `iters` launches of nwin windows each (pxy_le holds iters*nwin base points, 64 bytes each), alternating
 * between the context's two streams like bsgs_run; returns the elapsed device time of the whole train in *ms
 * (CUDA events: the first before anything starts on either stream, the last after both streams are done).
 * Hit records are discarded. */
int  bsgs_time_launches(bsgs_ctx *ctx, uint32_t nwin, const uint8_t *pxy_le, uint32_t iters, float *ms);
/* probe-only micro-kernel: nkeys pseudo-random 64-bit keys (splitmix64 of seed)
 * probed against the uploaded table; returns device ms and the number of keys that probed true. */
int  bsgs_probe_bench(bsgs_ctx *ctx, uint64_t nkeys, uint64_t seed, uint32_t iters, float *ms, uint64_t *nfound);
/* exact probe of caller-supplied keys (for parity tests of the table layouts) */
int  bsgs_probe_keys(bsgs_ctx *ctx, const uint64_t *keys, uint64_t nkeys, uint8_t *found);
/* geometry actually used by the kernel (threads, points per thread, bytes resident) */
typedef struct { uint32_t grid, block, points_per_thread, sm_count; uint64_t n_giants, bytes_giants, bytes_chain, bytes_table; } bsgs_info;
int  bsgs_get_info(const bsgs_ctx *ctx, bsgs_info *info);
/* device listing used by the parameter suggestion (`Tune`, 1_9_7File.pb:324-431): what the reference
 * reads with cuDeviceGetName / cuDeviceComputeCapability / cuDeviceGetAttribute(16) / cuMemGetInfo_v2
 * before it prints "Cuda device: ... Device have: MP:.. Cores+.." (1_9_7File.pb:778-860).  Needs no
 * context; returns BSGS_E_ARG for a device index that does not exist. */
typedef struct { char name[128]; int32_t cc_major, cc_minor, sm_count, reserved; uint64_t free_bytes, total_bytes; } bsgs_devinfo;
int  bsgs_device_count(int *count);
int  bsgs_device_query(int device, bsgs_devinfo *out);
/* any time: "verify" (0|1; needs payloads: a device-side build or bsgs_upload_htcpu), "overlap" (1|0: two launches
 * in flight in bsgs_run / bsgs_time_launches, default 1), "force_exact" (0|1, tests: every candidate key through
 * the kernel's out-of-line exact path), "trace" (0|1, diagnostics: per-warp timestamps, bsgs_debug_trace).
 * tuning knobs, call before uploads: "blocks_per_sm" (1..3 resident 256-thread CTAs per SM; the
 * grid is SMs x this), "block" (256, or 320 = two 320-thread CTAs per SM), "max_windows" (sizes the chain buffer), "l2_fetch" (32|64|128: device-wide
 * DRAM->L2 fetch granularity) */
int  bsgs_set_option(bsgs_ctx *ctx, const char *name, int64_t value);

/* -- diagnostics ---------------------------------------------------------------------------- */
/* element-wise field operation on n pairs of 8 x u32 little-endian limbs, executed on the device
 * (op: 0 mul, 1 sqr, 2 inv, 3 sub(canonical), 4 add(lazy), 5 canonicalise, 6 low64 of canonical
 * (a+b), 7 low64 of X(2P) given (Px, 1/(2Py)), 8 reduction of the 512-bit value (b << 256) | a).  Used by the
 * parity tests of the limb arithmetic. */
int  bsgs_debug_field(int device, int op, const uint32_t *a, const uint32_t *b, uint32_t *r, uint32_t n);
/* the giant-step kernel's own key extraction (the square of lambda taken without reduction, the key read
 * off its low 64 bits and its limbs of weight >= 2^224, plus the out-of-line exact path it falls back to for
 * ~2^-29 of the candidates) run on n inputs: keys[i] = low 64 bits of canonical (in[i]^2 + A[i] - Q[i]) mod p for
 * mode 0 (in[i] = lambda, 8 limbs, any 256-bit value) or of (in[i] + A[i] - Q[i]) mod p for mode 1 (in[i] = an
 * arbitrary 512-bit value in place of the square, 16 limbs); A, Q < p.  exact[i] = 1 when the classification
 * asked for the exact path; force != 0 sends every input there.  Parity test of that rare path. */
int  bsgs_debug_key(int device, int mode, const uint32_t *in, const uint32_t *A, const uint32_t *Q, uint64_t *keys, uint8_t *exact, uint32_t n, int force);

/* per-warp {SM id, %globaltimer (ns) at kernel start, after PASS 1, at the end} of the last giant-step launch, 4 words
 * per warp; enabled with bsgs_set_option(ctx, "trace", 1).  Used to measure how evenly the warps of an SM progress. */
int  bsgs_debug_trace(bsgs_ctx *ctx, uint64_t *out, uint32_t cap_warps, uint32_t *nwarps);

#ifdef __cplusplus
}
#endif
#endif
