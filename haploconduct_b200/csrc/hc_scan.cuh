// hc_scan.cuh -- exclusive scan of uint32 counts into 64-bit offsets over many blocks (reduce -> scan of the block
// sums -> scan within the blocks); the ordered compactions of hc_fno.cu rest on it.
#ifndef HC_SCAN_CUH_
#define HC_SCAN_CUH_
#include <cstdint>
#include <cuda_runtime.h>

namespace hc_scan {
namespace {     // internal linkage: the header is included by several translation units

typedef unsigned long long u64;
constexpr int THREADS = 1024;
constexpr int ITEMS = 4;                       // per thread
constexpr int PER_BLOCK = THREADS * ITEMS;

__device__ __forceinline__ u64 block_exclusive(u64 v, u64* wsum, u64* total) {   // exclusive prefix of v over the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const u64 w = wsum[lane];
        u64 s = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        wsum[lane] = s - w;
        if (lane == 31) *total = s;
    }
    __syncthreads();
    return wsum[warp] + inc - v;
}

__global__ void __launch_bounds__(THREADS) reduce_blocks(const uint32_t* __restrict__ in, u64 n, u64* bsum) {
    __shared__ u64 wsum[32];
    __shared__ u64 total;
    const u64 base = (u64)blockIdx.x * PER_BLOCK + (u64)threadIdx.x * ITEMS;
    u64 v = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) if (base + j < n) v += in[base + j];
    block_exclusive(v, wsum, &total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(THREADS) scan_block_sums(u64* bsum, u64 nb, u64* total_out) {   // in place, one block
    __shared__ u64 wsum[32];
    __shared__ u64 total;
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u64 b0 = 0; b0 < nb; b0 += THREADS) {
        const u64 i = b0 + threadIdx.x;
        const u64 v = i < nb ? bsum[i] : 0;
        const u64 ex = block_exclusive(v, wsum, &total);
        const u64 c = carry;
        if (i < nb) bsum[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(THREADS) scan_within(const uint32_t* __restrict__ in, u64 n, const u64* __restrict__ boff, u64* out) {
    __shared__ u64 wsum[32];
    __shared__ u64 total;
    const u64 base = (u64)blockIdx.x * PER_BLOCK + (u64)threadIdx.x * ITEMS;
    uint32_t x[ITEMS];
    u64 v = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) { x[j] = base + j < n ? in[base + j] : 0u; v += x[j]; }
    u64 run = boff[blockIdx.x] + block_exclusive(v, wsum, &total);
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        if (base + j < n) out[base + j] = run;
        run += x[j];
    }
}

// out[i] = sum of in[0..i), *d_total = sum of all; d_bsum needs (n + PER_BLOCK - 1) / PER_BLOCK entries
inline void exclusive_u32(const uint32_t* d_in, u64 n, u64* d_out, u64* d_total, u64* d_bsum, cudaStream_t st) {
    const u64 nb = (n + PER_BLOCK - 1) / PER_BLOCK;
    if (nb == 0) { cudaMemsetAsync(d_total, 0, sizeof(u64), st); return; }
    reduce_blocks<<<(unsigned)nb, THREADS, 0, st>>>(d_in, n, d_bsum);
    scan_block_sums<<<1, THREADS, 0, st>>>(d_bsum, nb, d_total);
    scan_within<<<(unsigned)nb, THREADS, 0, st>>>(d_in, n, d_bsum, d_out);
}

inline u64 blocks_for(u64 n) { return (n + PER_BLOCK - 1) / PER_BLOCK + 1; }

}  // namespace
}  // namespace hc_scan
#endif
