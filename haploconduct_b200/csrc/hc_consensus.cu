// hc_consensus.cu -- the per-column accumulation of SRBuilder::consensus_pos (src/SRBuilder.cpp:297-348) for whole
// pile-ups (SRBuilder::consensus, :406-522) on the device: for every consensus column the four log10 scores
// (A, C, T, G), added in list order with the reference's own addends (host-libm table, one entry per Phred value),
// and the number of sequences covering the column.  What follows per column (:349-401: five pow(10, .), one
// log10, the comparisons) is evaluated here with the device's pow / log10; the scores are bit-identical to the
// reference's, the transcendental step is not guaranteed to be, so every column whose outcome lies within the
// error bound of a decision (quality rounding, the minQual and 10^-9.3 thresholds, denormal powers) is marked and
// re-evaluated by the host part of hc_consensus with the host libm.
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
                                                 const int8_t* __restrict__ code_to_q, double D_min_qual, double* sums,
                                                 uint16_t* count, char* cbase, char* cqual, u64* marked, u64 marked_cap,
                                                 unsigned long long* n_marked) {
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
    // ---- :349-401 with the device's pow / log10.  Every comparison whose outcome could differ from the host libm's
    // (a few ulp in pow, amplified by 1 - max/total) marks the column; the host redoes marked columns from the scores.
    const double mx = fmax(fmax(sA, sT), fmax(sC, sG));
    const double max_prob = pow(10.0, mx);
    const double total = __dadd_rn(__dadd_rn(__dadd_rn(pow(10.0, sA), pow(10.0, sT)), pow(10.0, sC)), pow(10.0, sG));
    char b = 'N', q = '$';
    bool mark = false;
    if (mx == 0.0) {
        // all scores zero: no base contributed -> 'N' (exact test on exact scores)
    } else if (mx < -290.0) {
        mark = true;                                             // powers near or below the denormal range: leave it to the host
    } else {
        const double r = __ddiv_rn(max_prob, total);
        const double p_inc = 1.0 - r;
        const double keep = 1.0 - p_inc;
        const double eps = 64.0 * 2.220446049250313e-16;         // bound on the relative error of r
        if (n_active > 1 && fabs(keep - D_min_qual) <= eps) mark = true;
        if (n_active > 1 && keep < D_min_qual) {
            // 'N', '$'
        } else {
            const double lim = 5.011872336272715e-10;            // 10^-9.3
            int phred;
            if (fabs(p_inc - lim) <= 1e-3 * lim) mark = true;
            if (p_inc < lim) phred = 93;
            else {
                const double x = -10.0 * log10(p_inc);
                const double fr = x - floor(x);
                // |dx| <= 4.35 * eps / p_inc; p_inc >= 5e-10 here
                if (fabs(fr - 0.5) <= 4.35 * eps / p_inc + 1e-9) mark = true;
                phred = (int)floor(x + 0.5);
            }
            phred = phred < 0 ? 0 : (phred > 93 ? 93 : phred);
            b = mx == sA ? 'A' : (mx == sT ? 'T' : (mx == sC ? 'C' : 'G'));
            q = (char)(phred + 33);
        }
    }
    cbase[g] = b;
    cqual[g] = q;
    if (mark) {
        const unsigned long long k = atomicAdd(n_marked, 1ull);
        if (k < marked_cap) marked[k] = g;
    }
}

__global__ void cons_gather_marked(const u64* __restrict__ marked, u64 n, const double* __restrict__ sums, double* out) {
    const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const u64 g = marked[k];
    out[4 * k + 0] = sums[4 * g + 0]; out[4 * k + 1] = sums[4 * g + 1]; out[4 * k + 2] = sums[4 * g + 2]; out[4 * k + 3] = sums[4 * g + 3];
}

}  // namespace

cudaError_t hc_launch_cons_sums(const hc_cons_dev& D, const hc_cons_problem* d_prob, uint64_t n_prob, const hc_cons_seq* d_seqs,
                                const unsigned long long* d_tile_off, uint64_t n_tiles, const double* d_addend,
                                const int8_t* d_code_to_q, double min_qual, double* d_sums, uint16_t* d_count, char* d_base,
                                char* d_qual, unsigned long long* d_marked, uint64_t marked_cap, unsigned long long* d_n_marked,
                                cudaStream_t st) {
    if (n_tiles == 0) return cudaSuccess;
    cons_sums<<<(unsigned)n_tiles, 256, 0, st>>>(D, d_prob, n_prob, d_seqs, d_tile_off, d_addend, d_code_to_q, min_qual, d_sums, d_count,
                                                 d_base, d_qual, d_marked, marked_cap, d_n_marked);
    return cudaGetLastError();
}

cudaError_t hc_launch_cons_gather(const unsigned long long* d_marked, uint64_t n, const double* d_sums, double* d_out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    cons_gather_marked<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_marked, n, d_sums, d_out);
    return cudaGetLastError();
}
