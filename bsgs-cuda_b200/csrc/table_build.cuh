// table_build.cuh -- device-side construction of the baby-step hash table in the reference's own
// CSR form (the htGPU / htCPU file images, 1_9_7File.pb:3232-3444) from 64-bit keys.
//
// The reference builds the table on the CPU: HashTableInsert into per-bucket arrays
// (1_9_7File.pb:2555-2622), a per-bucket insertion sort by tag (sortHashTable32bit, 2771-2820) and a
// pack into "H exclusive prefix counts | total | items" (packHTFile / packHTGPUFile).  Here the same
// image is produced in HBM by a counting sort: histogram of buckets -> exclusive scan -> scatter ->
// per-bucket sort by (tag ascending, j descending: a later insert lands before an equal tag,
// 1_9_7File.pb:2797-2805).  The result is byte-identical to the oracle's orc_build_ht.
//
// At w = 2^30 the bucket counters (1 GB) and the scatter targets are far larger than the 126 MB L2, and the
// histogram and the scatter ran at the speed of random DRAM atomics (50 + 155 ms of a 380 ms build, ncu).
// So the keys are first split into 256 partitions by the top 8 bits of the bucket index (one pass with
// shared-memory counting sort per 2048-key tile: one global reservation per tile and partition, writes in runs),
// and the fine histogram and scatter then walk the partitioned copy in order: at any moment the grid works
// inside one or two partitions, whose counters (H/256 x 4 B) and output range stay resident in L2.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

namespace tbl {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// measurement tables: w uniformly random 64-bit keys (statistically the low 64 bits of X of random points)
__global__ void k_fill_synth(uint64_t* __restrict__ keys, uint64_t w, uint64_t seed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < w; i += stride) keys[i] = mix64(seed + i);
}

// bucket = low 32 bits & (H-1) (1_9_7File.pb:2561)
__global__ void k_hist(const uint64_t* __restrict__ keys, uint64_t w, uint32_t mask, uint32_t* __restrict__ cnt) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < w; i += stride) atomicAdd(cnt + ((uint32_t)keys[i] & mask), 1u);
}

// tag = bits 32..63 (1_9_7File.pb:2583), payload = j (1_9_7File.pb:2584)
__global__ void k_scatter(const uint64_t* __restrict__ keys, uint64_t w, uint32_t mask, uint32_t* __restrict__ cursor,
                          uint32_t* __restrict__ tags, uint32_t* __restrict__ js) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < w; i += stride) {
        uint64_t k = keys[i];
        uint32_t pos = atomicAdd(cursor + ((uint32_t)k & mask), 1u);
        tags[pos] = (uint32_t)(k >> 32);
        js[pos] = (uint32_t)i;
    }
}

// ---- 256-way split by the top bits of the bucket index -------------------------------------------------------
constexpr int kParts = 256, kSplitThreads = 256, kSplitPer = 8, kSplitTile = kSplitThreads * kSplitPer;

__global__ void __launch_bounds__(256) k_hist_coarse(const uint64_t* __restrict__ keys, uint64_t w, uint32_t mask, int shift, unsigned long long* __restrict__ gcount) {
    __shared__ uint32_t h[kParts];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < w; i += stride) atomicAdd(&h[((uint32_t)keys[i] & mask) >> shift], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(gcount + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}
__global__ void k_scan_parts(unsigned long long* gcount) {       // exclusive prefix of 256 counts, in place (one thread: 256 additions)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long run = 0;
        for (int p = 0; p < kParts; p++) { unsigned long long c = gcount[p]; gcount[p] = run; run += c; }
    }
}
// tile of 2048 keys: counted per partition in shared memory (the atomic returns the key's rank inside the tile), sorted
// by partition INSIDE shared memory, the tile's share of every partition reserved with one global atomic, and written
// out element by element of the sorted tile -- consecutive threads write consecutive addresses of a partition (runs of
// whole sectors instead of 2048 scattered 8-byte writes)
__global__ void __launch_bounds__(kSplitThreads) k_split(const uint64_t* __restrict__ keys, uint64_t w, uint32_t mask, int shift,
                                                         unsigned long long* __restrict__ gcursor, uint64_t* __restrict__ tkeys, uint32_t* __restrict__ tjs) {
    __shared__ uint32_t h[kParts], toff[kParts];                  // count / exclusive offset of a partition inside the tile
    __shared__ unsigned long long base[kParts];                   // global position of the tile's first element of a partition
    __shared__ uint64_t skey[kSplitTile];
    __shared__ uint32_t sj[kSplitTile];
    __shared__ uint8_t spart[kSplitTile];
    const uint64_t ntiles = (w + kSplitTile - 1) / kSplitTile;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        h[threadIdx.x] = 0;
        __syncthreads();
        uint64_t k[kSplitPer]; uint32_t rank[kSplitPer];
        const uint64_t i0 = tile * kSplitTile + threadIdx.x;
        const uint32_t n = (uint32_t)(w - tile * kSplitTile < (uint64_t)kSplitTile ? w - tile * kSplitTile : kSplitTile);
#pragma unroll
        for (int q = 0; q < kSplitPer; q++) {
            const uint64_t i = i0 + (uint64_t)q * kSplitThreads;
            if (i < w) { k[q] = keys[i]; rank[q] = atomicAdd(&h[((uint32_t)k[q] & mask) >> shift], 1u); }
        }
        __syncthreads();
        // exclusive scan of the 256 counts (Hillis-Steele on a copy), and the global reservation
        const uint32_t mine = h[threadIdx.x];
        toff[threadIdx.x] = mine;
        __syncthreads();
#pragma unroll
        for (int d = 1; d < kParts; d <<= 1) {
            const uint32_t v = threadIdx.x >= d ? toff[threadIdx.x - d] : 0u;
            __syncthreads();
            toff[threadIdx.x] += v;
            __syncthreads();
        }
        const uint32_t excl = toff[threadIdx.x] - mine;
        __syncthreads();
        toff[threadIdx.x] = excl;
        if (mine) base[threadIdx.x] = atomicAdd(gcursor + threadIdx.x, (unsigned long long)mine);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kSplitPer; q++) {
            const uint64_t i = i0 + (uint64_t)q * kSplitThreads;
            if (i < w) {
                const uint32_t part = ((uint32_t)k[q] & mask) >> shift, e = toff[part] + rank[q];
                skey[e] = k[q]; sj[e] = (uint32_t)i; spart[e] = (uint8_t)part;
            }
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < n; e += kSplitThreads) {
            const uint32_t part = spart[e];
            const unsigned long long pos = base[part] + (e - toff[part]);
            tkeys[pos] = skey[e]; tjs[pos] = sj[e];
        }
        __syncthreads();
    }
}
// scatter of the partitioned copy: as k_scatter, the payload is the original index carried along
__global__ void k_scatter_split(const uint64_t* __restrict__ tkeys, const uint32_t* __restrict__ tjs, uint64_t w, uint32_t mask, uint32_t* __restrict__ cursor,
                                uint32_t* __restrict__ tags, uint32_t* __restrict__ js) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < w; i += stride) {
        uint64_t k = tkeys[i];
        uint32_t pos = atomicAdd(cursor + ((uint32_t)k & mask), 1u);
        tags[pos] = (uint32_t)(k >> 32);
        js[pos] = tjs[i];
    }
}

// one thread per bucket: insertion sort by (tag asc, j desc); makes the scatter order irrelevant
__global__ void k_sort_buckets(const uint32_t* __restrict__ off, uint32_t* __restrict__ tags, uint32_t* __restrict__ js, uint64_t H) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    uint32_t lo = off[h], hi = off[h + 1];
    for (uint32_t i = lo + 1; i < hi; i++) {
        uint32_t t = tags[i], j = js[i], k = i;
        while (k > lo && (tags[k - 1] > t || (tags[k - 1] == t && js[k - 1] < j))) { tags[k] = tags[k - 1]; js[k] = js[k - 1]; k--; }
        tags[k] = t; js[k] = j;
    }
}

// htCPU item stream {tag, j} (packHTFile, 1_9_7File.pb:3232-3335)
__global__ void k_interleave(const uint32_t* __restrict__ tags, const uint32_t* __restrict__ js, uint2* __restrict__ items, uint64_t w) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < w; i += stride) items[i] = make_uint2(tags[i], js[i]);
}

// keys[w] -> off[H+1] (exclusive prefix, off[H] = w) | tags[w]  (+ js[w]); `cursor` is H u32 of scratch
static inline cudaError_t build_csr(const uint64_t* keys, uint64_t w, uint64_t H, uint32_t* off, uint32_t* tags, uint32_t* js,
                                    uint32_t* cursor, int sm_count, cudaStream_t st) {
    cudaError_t e;
    uint32_t mask = (uint32_t)(H - 1);
    int grid = sm_count * 8;
    // partitioned copy (tables with at least 256 buckets; when the 12 bytes per key of scratch do not fit, the direct path)
    uint64_t* tkeys = nullptr; uint32_t* tjs = nullptr; unsigned long long* gpart = nullptr;
    int shift = 0;
    bool split = H >= (uint64_t)kParts && w > 0;
    if (split) {
        while (((uint64_t)kParts << shift) < H) shift++;              // partition = bucket >> shift
        if (cudaMalloc(&tkeys, w * 8) != cudaSuccess || cudaMalloc(&tjs, w * 4) != cudaSuccess || cudaMalloc(&gpart, kParts * 8) != cudaSuccess) {
            cudaGetLastError(); cudaFree(tkeys); cudaFree(tjs); cudaFree(gpart); tkeys = nullptr; tjs = nullptr; gpart = nullptr; split = false;
        }
    }
    auto done = [&](cudaError_t r) { cudaFree(tkeys); cudaFree(tjs); cudaFree(gpart); return r; };
    if (split) {
        if ((e = cudaMemsetAsync(gpart, 0, kParts * 8, st)) != cudaSuccess) return done(e);
        k_hist_coarse<<<grid, 256, 0, st>>>(keys, w, mask, shift, gpart);
        k_scan_parts<<<1, 32, 0, st>>>(gpart);
        k_split<<<grid, kSplitThreads, 0, st>>>(keys, w, mask, shift, gpart, tkeys, tjs);
        if ((e = cudaGetLastError()) != cudaSuccess) return done(e);
    }
    const uint64_t* hkeys = split ? tkeys : keys;
    if ((e = cudaMemsetAsync(off, 0, (H + 1) * 4, st)) != cudaSuccess) return done(e);
    if (w) k_hist<<<grid, 256, 0, st>>>(hkeys, w, mask, off);
    void* tmp = nullptr; size_t tmp_bytes = 0;
    if ((e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, off, off, (int64_t)(H + 1), st)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) return done(e);
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, off, off, (int64_t)(H + 1), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cursor, off, H * 4, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && w) {
        if (split) k_scatter_split<<<grid, 256, 0, st>>>(tkeys, tjs, w, mask, cursor, tags, js);
        else k_scatter<<<grid, 256, 0, st>>>(keys, w, mask, cursor, tags, js);
        k_sort_buckets<<<(unsigned)((H + 255) / 256), 256, 0, st>>>(off, tags, js, H);
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(tmp);
    return done(e != cudaSuccess ? e : e2);
}

}  // namespace tbl
