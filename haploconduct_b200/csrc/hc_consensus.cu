// hc_consensus.cu -- the per-column accumulation of SRBuilder::consensus_pos (src/SRBuilder.cpp:297-348) for whole
// pile-ups (SRBuilder::consensus, :406-522) on the device: for every consensus column the four log10 scores
// (A, C, T, G), added in list order with the reference's own addends (host-libm table, one entry per Phred value),
// and the number of sequences covering the column.  What follows per column (:349-401: five pow(10, .), one
// log10, the comparisons) is evaluated by the host part of hc_consensus with the host libm -- the scores here
// are bit-identical to the reference's, so that step is too.
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/hc_b200.h"
#include "hc_layout.h"
#include "hc_consensus.cuh"

namespace {

typedef unsigned long long u64;

__global__ void __launch_bounds__(256) cons_sums(hc_cons_dev D, const hc_cons_problem* __restrict__ prob, u64 n_prob,
                                                 const hc_cons_seq* __restrict__ seqs, const u64* __restrict__ col_off,
                                                 const double* __restrict__ addend /* [94][2]: hit, miss */,
                                                 const int8_t* __restrict__ code_to_q, double* sums, uint16_t* count) {
    // one block per (problem, 256-column tile): blockIdx.x enumerates tiles through col_off (tile -> problem by search)
    __shared__ u64 s_prob;
    if (threadIdx.x == 0) {
        // col_off[p] = first global tile of problem p; find p with col_off[p] <= blockIdx.x < col_off[p + 1]
        u64 lo = 0, hi = n_prob;
        while (hi - lo > 1) {
            const u64 mid = (lo + hi) >> 1;
            if (col_off[mid] <= blockIdx.x) lo = mid; else hi = mid;
        }
        s_prob = lo;
    }
    __syncthreads();
    const hc_cons_problem P = prob[s_prob];
    const int c = (int)((blockIdx.x - col_off[s_prob]) * 256 + threadIdx.x);
    if (c >= P.total_len) return;
    double sA = 0.0, sC = 0.0, sT = 0.0, sG = 0.0;
    uint32_t n_active = 0;
    for (u64 j = P.seq_begin; j < P.seq_end; j++) {
        const hc_cons_seq e = seqs[j];
        if (c < e.pos) break;                                   // pos ascending: nobody further down has started
        const hc_rdesc rd = D.rdesc[e.read];
        const uint32_t len = rd.len[e.mate] & HC_LEN_MASK;
        const uint32_t p = (uint32_t)(c - e.pos);
        if (p >= len) continue;                                  // this sequence has ended (:482-484)
        n_active++;
        const u64 at = 16ull * rd.slot16[e.mate] + (e.rc ? hc_slot_size(len) : 0u) + p;
        uint32_t code, base;
        bool isN;
        if (D.packed) {
            const uint32_t b = D.pk[at];
            code = b & 63u; base = b >> 6; isN = b == 0u;
        } else {
            code = D.qual[at];
            base = (D.base2[at >> 4] >> (2 * (at & 15))) & 3u;
            isN = (D.nmask[at >> 5] >> (at & 31)) & 1u;
        }
        if (isN) continue;                                       // counted, contributes nothing (:343-348)
        const int q = code_to_q[code];
        const double hit = addend[2 * q], miss = addend[2 * q + 1];
        // base codes of the store: 0 A, 1 C, 2 G, 3 T
        sA = __dadd_rn(sA, base == 0u ? hit : miss);
        sC = __dadd_rn(sC, base == 1u ? hit : miss);
        sG = __dadd_rn(sG, base == 2u ? hit : miss);
        sT = __dadd_rn(sT, base == 3u ? hit : miss);
    }
    const u64 g = P.out_offset + (u64)c;
    sums[4 * g + 0] = sA; sums[4 * g + 1] = sC; sums[4 * g + 2] = sT; sums[4 * g + 3] = sG;
    count[g] = (uint16_t)(n_active > 0xffffu ? 0xffffu : n_active);
}

}  // namespace

cudaError_t hc_launch_cons_sums(const hc_cons_dev& D, const hc_cons_problem* d_prob, uint64_t n_prob, const hc_cons_seq* d_seqs,
                                const unsigned long long* d_tile_off, uint64_t n_tiles, const double* d_addend,
                                const int8_t* d_code_to_q, double* d_sums, uint16_t* d_count, cudaStream_t st) {
    if (n_tiles == 0) return cudaSuccess;
    cons_sums<<<(unsigned)n_tiles, 256, 0, st>>>(D, d_prob, n_prob, d_seqs, d_tile_off, d_addend, d_code_to_q, d_sums, d_count);
    return cudaGetLastError();
}
