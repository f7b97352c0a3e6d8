// Microbenchmark: L1 data-pipe wavefronts per warp-wide global load, by access width and address pattern.
// nvcc -gencode arch=compute_100a,code=sm_100a -o l1wf l1_wavefronts.cu ; ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,smsp__inst_executed_op_global_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ./l1wf
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int W, int PAT>   // W = bytes per lane (4, 8, 16); PAT: 0 coalesced, 1 lane stride 32 B, 2 lane stride 32 B + 20 B misalignment base, 3 one 128-B line per lane
__global__ void k(const uint8_t* __restrict__ buf, size_t span, uint32_t* out, int iters) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t acc = 0;
    size_t base = (warp * 8192) % span;
    for (int it = 0; it < iters; it++) {
        size_t off;
        if (PAT == 0) off = (size_t)lane * W;
        else if (PAT == 1) off = (size_t)lane * 32;
        else if (PAT == 2) off = (size_t)lane * 32 + 16;
        else off = (size_t)lane * 128;
        const uint8_t* p = buf + base + off;
        if (W == 4) acc += __ldg(reinterpret_cast<const uint32_t*>(p));
        else if (W == 8) { uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); acc += v.x ^ v.y; }
        else { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); acc += v.x ^ v.y ^ v.z ^ v.w; }
        base = (base + 4096 * 37) % span;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int W, int PAT>
void run(const uint8_t* buf, size_t span, uint32_t* out, const char* name) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<W, PAT><<<148 * 4, 256>>>(buf, span, out, 64);
    cudaEventRecord(a);
    k<W, PAT><<<148 * 4, 256>>>(buf, span, out, 2048);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double instr = 148.0 * 4 * 8 * 2048;
    printf("%-28s W=%2d  %.3f ms  %.2f ns per warp-load per SM-slot  (%.1f G warp-loads/s)\n", name, W, ms, ms * 1e6 / (instr / 148), instr / ms / 1e6);
}

int main() {
    const size_t span = 64ull << 20;   // 64 MB: L2 resident, far beyond L1
    uint8_t* buf; uint32_t* out;
    cudaMalloc(&buf, span + (1 << 20));
    cudaMemset(buf, 1, span + (1 << 20));
    cudaMalloc(&out, 148 * 4 * 256 * 4);
    run<4, 0>(buf, span, out, "coalesced");
    run<8, 0>(buf, span, out, "coalesced");
    run<16, 0>(buf, span, out, "coalesced");
    run<4, 1>(buf, span, out, "lane stride 32 B");
    run<8, 1>(buf, span, out, "lane stride 32 B");
    run<16, 1>(buf, span, out, "lane stride 32 B");
    run<16, 2>(buf, span, out, "lane stride 32 B (+16)");
    run<4, 3>(buf, span, out, "one line per lane");
    run<8, 3>(buf, span, out, "one line per lane");
    run<16, 3>(buf, span, out, "one line per lane");
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
