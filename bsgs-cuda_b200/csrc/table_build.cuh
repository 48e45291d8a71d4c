// table_build.cuh -- device-side construction of the baby-step hash table in the reference's own
// CSR form (the htGPU / htCPU file images, 1_9_7File.pb:3232-3444) from 64-bit keys.
//
// The reference builds the table on the CPU: HashTableInsert into per-bucket arrays
// (1_9_7File.pb:2555-2622), a per-bucket insertion sort by tag (sortHashTable32bit, 2771-2820) and a
// pack into "H exclusive prefix counts | total | items" (packHTFile / packHTGPUFile).  Here the same
// image is produced in HBM by a counting sort: histogram of buckets -> exclusive scan -> scatter ->
// per-bucket sort by (tag ascending, j descending: a later insert lands before an equal tag,
// 1_9_7File.pb:2797-2805).  The result is byte-identical to the oracle's orc_build_ht.
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
    if ((e = cudaMemsetAsync(off, 0, (H + 1) * 4, st)) != cudaSuccess) return e;
    if (w) k_hist<<<grid, 256, 0, st>>>(keys, w, mask, off);
    void* tmp = nullptr; size_t tmp_bytes = 0;
    if ((e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, off, off, (int64_t)(H + 1), st)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, off, off, (int64_t)(H + 1), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cursor, off, H * 4, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && w) {
        k_scatter<<<grid, 256, 0, st>>>(keys, w, mask, cursor, tags, js);
        k_sort_buckets<<<(unsigned)((H + 255) / 256), 256, 0, st>>>(off, tags, js, H);
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(tmp);
    return e != cudaSuccess ? e : e2;
}

}  // namespace tbl
