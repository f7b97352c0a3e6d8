// hc_kernels.cu -- sm_100a kernels of the overlap-edge scoring path.
//
// What the reference does per candidate (src/EdgeCalculator.cpp:143-385 -> :67-139 -> :26-56):
// pick one or two windows from read types / orientations / ord, walk each window base by base,
// add log p(q_A, q_B, match?) in double, count mismatches, skip N, then exp(mean) and three
// threshold tests (:404-413).
//
// How it is done here:
//   * hc_score_kernel     one warp owns a tile of 32 candidates.  Lane = candidate for window
//                         selection and for the final decision; in between, the tile's windows are
//                         cut into 16-position chunks and the chunks are dealt to the 32 lanes round
//                         robin, so short and long windows fill the warp equally well.  A chunk is
//                         1 x 128-bit load (B side, always aligned) + 5 x 32-bit loads and 4 funnel
//                         shifts (A side) of quality codes, 3 loads + 1 funnel shift + XOR/popc of
//                         2-bit bases, and 16 shared-memory lookups of a fixed-point -log p table
//                         indexed by (code_A, code_B, mismatch).  Sums are integers, hence exact and
//                         order independent.  Candidates whose mean lands within HC_FX_MARGIN of a
//                         threshold are queued for
//   * hc_exact_kernel     which re-adds the reference's own double addends in the reference's
//                         order (one thread per queued candidate), so every decision is bit-exact.
//   * hc_compact_*        order-preserving compaction of accepted edges / non-edge overlaps
//                         (count -> scan -> scatter): output order = input order = the reference's
//                         1-thread order.
// No tensor cores: nothing here is a contraction.
#include "hc_kernels.cuh"

namespace {

typedef unsigned long long u64;

struct Win {
    u64 xpos;         // position of A[pos] in the store's position space
    uint32_t ypos16;  // position of B[0] / 16
    uint32_t L;       // window length (src/EdgeCalculator.cpp:88); 0 when not scored
    uint32_t status;  // HC_WIN_*
    uint32_t hasN;
};

struct CandSetup {
    Win w[2];
    int32_t pos3, pos4;
    uint32_t two;   // two windows (any paired read involved)
    uint32_t err;
};

__device__ __forceinline__ uint32_t d_len(const hc_rdesc& d, int m) { return m ? d.len[1] : d.len[0]; }
__device__ __forceinline__ uint32_t d_slot(const hc_rdesc& d, int m) { return m ? d.slot16[1] : d.slot16[0]; }

// overlap_score's guards and window length, src/EdgeCalculator.cpp:74-88
__device__ __forceinline__ void make_window(const hc_kparams& P, const hc_rdesc& dA, int mA, int rcA, const hc_rdesc& dB,
                                            int mB, int rcB, uint32_t pos, Win& w) {
    uint32_t rawA = d_len(dA, mA), rawB = d_len(dB, mB);
    uint32_t lenA = rawA & HC_LEN_MASK, lenB = rawB & HC_LEN_MASK;
    w.L = 0;
    w.xpos = 0;
    w.ypos16 = 0;
    w.hasN = 0;
    if (pos >= lenA) { w.status = HC_WIN_POS_OOR; return; }
    if (lenA < P.min_read_len || lenB < P.min_read_len) { w.status = HC_WIN_SHORT; return; }
    u64 sa = 16ull * d_slot(dA, mA) + (rcA ? hc_slot_size(lenA) : 0u);
    u64 sb = 16ull * d_slot(dB, mB) + (rcB ? hc_slot_size(lenB) : 0u);
    w.xpos = sa + pos;
    w.ypos16 = (uint32_t)(sb >> 4);
    w.L = min(lenA - pos, lenB);
    w.status = HC_WIN_SCORED;
    w.hasN = ((rawA | rawB) & HC_HASN_BIT) ? 1u : 0u;
}

// Window selection of EdgeCalculator::compute_overlap, src/EdgeCalculator.cpp:199-351, and the
// extra positions :222,:262-263,:300-301,:361-372.
__device__ __forceinline__ void setup_candidate(const hc_kparams& P, const hc_candidate& c, CandSetup& s) {
    s.err = 0;
    s.two = 0;
    s.pos3 = s.pos4 = 0;
    s.w[0].L = s.w[1].L = 0;
    s.w[0].status = s.w[1].status = HC_WIN_UNUSED;
    s.w[0].xpos = s.w[1].xpos = 0;
    s.w[0].ypos16 = s.w[1].ypos16 = 0;
    s.w[0].hasN = s.w[1].hasN = 0;
    if (c.idx1 >= P.n_reads || c.idx2 >= P.n_reads || c.idx1 == c.idx2) { s.err = 1; return; }
    const uint4 r1 = __ldg((const uint4*)(P.rdesc + c.idx1));
    const uint4 r2 = __ldg((const uint4*)(P.rdesc + c.idx2));
    hc_rdesc d1, d2;
    d1.slot16[0] = r1.x; d1.slot16[1] = r1.y; d1.len[0] = r1.z; d1.len[1] = r1.w;
    d2.slot16[0] = r2.x; d2.slot16[1] = r2.y; d2.len[0] = r2.z; d2.len[1] = r2.w;
    const int p1 = (d1.len[1] & HC_LEN_MASK) != 0, p2 = (d2.len[1] & HC_LEN_MASK) != 0;  // Read::is_paired()
    const int rc1 = c.ori1 ? 0 : 1, rc2 = c.ori2 ? 0 : 1;
    const int f1 = c.ori1 ? 0 : 1, s1 = 1 - f1, f2 = c.ori2 ? 0 : 1, s2 = 1 - f2;
    const uint32_t l10 = d1.len[0] & HC_LEN_MASK, l11 = d1.len[1] & HC_LEN_MASK;
    const uint32_t l20 = d2.len[0] & HC_LEN_MASK, l21 = d2.len[1] & HC_LEN_MASK;
    if (!p1 && !p2) {                                    // S-S :199-233
        if (P.n_single == 0) { s.err = 1; return; }
        make_window(P, d1, 0, rc1, d2, 0, rc2, c.pos1, s.w[0]);
        s.pos3 = (int32_t)(l10 - c.pos1 - l20);
    } else if (!p1 && p2) {                              // S-P :234-271
        if (P.n_single == 0) { s.err = 1; return; }
        make_window(P, d1, 0, rc1, d2, f2, rc2, c.pos1, s.w[0]);
        make_window(P, d1, 0, rc1, d2, s2, rc2, c.pos2, s.w[1]);
        s.two = 1;
        s.pos3 = (int32_t)(l10 - c.pos2 - l21);
        s.pos4 = (int32_t)(l10 - c.pos1 - l20);
    } else if (p1 && !p2) {                              // P-S :272-309
        if (P.n_single == 0) { s.err = 1; return; }
        make_window(P, d1, f1, rc1, d2, 0, rc2, c.pos1, s.w[0]);
        make_window(P, d2, 0, rc2, d1, s1, rc1, c.pos2, s.w[1]);
        s.two = 1;
        s.pos3 = (int32_t)(l11 + c.pos2 - l20);
        s.pos4 = (int32_t)(l20 + c.pos1 - l10);
    } else {                                             // P-P :312-380
        if (c.ord != '1' && c.ord != '2') { s.err = 1; return; }   // assert :369
        make_window(P, d1, f1, rc1, d2, f2, rc2, c.pos1, s.w[0]);
        if (c.ord == '1') {
            make_window(P, d1, s1, rc1, d2, s2, rc2, c.pos2, s.w[1]);
            s.pos3 = (int32_t)(l11 - c.pos2 - l21);
        } else {
            make_window(P, d2, s2, rc2, d1, s1, rc1, c.pos2, s.w[1]);
            s.pos3 = (int32_t)(l11 + c.pos2 - l21);
        }
        s.two = 1;
        s.pos4 = (int32_t)(l10 - c.pos1 - l20);
    }
}

// One 16-position chunk of one window.  Returns the fixed-point sum of -log p over the chunk,
// the mismatch count, the number of N positions and whether a void (p < ps.mismatch) pair was hit.
template <bool HAS_VOID>
__device__ __forceinline__ void process_chunk(const hc_kparams& P, const uint32_t* __restrict__ T, u64 xpos, uint32_t ypos16,
                                              uint32_t L, uint32_t hasN, uint32_t k, uint32_t& sum, uint32_t& mm,
                                              uint32_t& ncnt, uint32_t& vd) {
    const u64 xp = xpos + 16ull * k;
    const uint32_t rem = L - 16u * k;   // >= 1 positions left in the window
    const uint32_t yq = ypos16 + k;
    // ---- loads (all issued before first use)
    const uint32_t* qa = reinterpret_cast<const uint32_t*>(P.qual + (xp & ~3ull));
    const uint32_t w0 = __ldg(qa), w1 = __ldg(qa + 1), w2 = __ldg(qa + 2), w3 = __ldg(qa + 3), w4 = __ldg(qa + 4);
    const uint4 wb = __ldg(reinterpret_cast<const uint4*>(P.qual) + yq);
    const uint32_t* bap = P.base2 + (xp >> 4);
    const uint32_t b0 = __ldg(bap), b1 = __ldg(bap + 1);
    const uint32_t bb = __ldg(P.base2 + yq);
    // ---- mismatch mask in the 2-bit domain: XOR, fold pairs, popc
    const uint32_t ba = __funnelshift_r(b0, b1, ((uint32_t)xp & 15u) * 2u);
    const uint32_t x2 = ba ^ bb;
    uint32_t m = (x2 | (x2 >> 1)) & 0x55555555u;
    if (rem < 16u) m &= (1u << (2u * rem)) - 1u;
    ncnt = 0;
    if (hasN) {   // rare: either read contains an N (skipped positions, src/EdgeCalculator.cpp:35-39,122-124)
        const uint32_t* nap = P.nmask + (xp >> 5);
        const uint32_t n0 = __ldg(nap), n1 = __ldg(nap + 1);
        const uint32_t nA = __funnelshift_r(n0, n1, (uint32_t)xp & 31u) & 0xffffu;
        const uint32_t nB = (__ldg(P.nmask + (yq >> 1)) >> ((yq & 1u) * 16u)) & 0xffffu;
        uint32_t nn = nA | nB;
        if (rem < 16u) nn &= (1u << rem) - 1u;
        ncnt = __popc(nn);
        uint32_t s = nn;   // spread 16 bits to the even bit positions
        s = (s | (s << 8)) & 0x00ff00ffu;
        s = (s | (s << 4)) & 0x0f0f0f0fu;
        s = (s | (s << 2)) & 0x33333333u;
        s = (s | (s << 1)) & 0x55555555u;
        m &= ~s;
    }
    mm = __popc(m);
    // ---- quality codes: align the A side, build (row, column) byte pairs, look up
    const uint32_t sh = ((uint32_t)xp & 3u) * 8u;
    uint32_t wa[4];
    wa[0] = __funnelshift_r(w0, w1, sh);
    wa[1] = __funnelshift_r(w1, w2, sh);
    wa[2] = __funnelshift_r(w2, w3, sh);
    wa[3] = __funnelshift_r(w3, w4, sh);
    const uint32_t wbv[4] = {wb.x, wb.y, wb.z, wb.w};
    uint32_t acc = 0, orv = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        // mismatch bits of positions 4j..4j+3 (bits 0,2,4,6 of byte j of m) -> bit 7 of bytes 0..3
        const uint32_t mj = (m >> (8 * j)) & 0x55u;
        const uint32_t dep = (mj * 0x02082080u) & 0x80808080u;
        const uint32_t col = (wa[j] ^ hc_swz4(wbv[j])) | dep;
        const uint32_t lo = __byte_perm(col, wbv[j], 0x5140);
        const uint32_t hi = __byte_perm(col, wbv[j], 0x7362);
        const uint32_t t0 = T[lo & 0xffffu], t1 = T[lo >> 16], t2 = T[hi & 0xffffu], t3 = T[hi >> 16];
        acc += (t0 + t1) + (t2 + t3);
        if (HAS_VOID) orv |= (t0 | t1) | (t2 | t3);
    }
    sum = acc;
    vd = HAS_VOID ? ((orv & HC_VOID_BIT) ? 1u : 0u) : 0u;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

struct WinAcc {
    u64 S;          // fixed-point sum of -log p
    uint32_t mm;    // mismatches
    uint32_t nn;    // N positions
    uint32_t vd;    // void
};

// The decision of src/EdgeCalculator.cpp:254-261 (two windows) and :404-413, on per-window
// "mean above threshold" flags.  ov[] are the per-window scores (0 for the early outs).
__device__ __forceinline__ void combine(const hc_kparams& P, uint32_t two, const double ov[2], const double mmr[2],
                                        const int ae[2], const int ao[2], double& score, double& mmrate, uint32_t& cls) {
    int edge_by_score, ov_ok;
    if (two) {
        mmrate = fmax(mmr[0], mmr[1]);
        const int both = ae[0] && ae[1];
        score = both ? 0.5 * (ov[0] + ov[1]) : fmin(ov[0], ov[1]);
        edge_by_score = both;
        ov_ok = ao[0] && ao[1];
    } else {
        mmrate = mmr[0];
        score = ov[0];
        edge_by_score = ae[0];
        ov_ok = ao[0];
    }
    if (edge_by_score) cls = HC_CLASS_EDGE;
    else if (mmrate <= P.merge_contigs) cls = HC_CLASS_EDGE;
    else if (ov_ok) cls = HC_CLASS_NONEDGE;
    else cls = HC_CLASS_DISCARD;
}

__device__ __forceinline__ void write_result(const hc_kparams& P, u64 i, const CandSetup& s, double score, double mmrate,
                                             uint32_t cls, const uint32_t mmc[2], const uint32_t cmp[2],
                                             const uint32_t st[2], uint32_t exact) {
    hc_score16 t;
    t.score = score;
    t.mismatch_rate = mmrate;
    P.tmp[i] = t;
    P.cls[i] = (uint8_t)cls;
    if (P.per_cand) {
        hc_result r;
        r.score = score;
        r.mismatch_rate = mmrate;
        r.pos3 = s.pos3;
        r.pos4 = s.pos4;
        r.mismatches[0] = mmc[0];
        r.mismatches[1] = mmc[1];
        r.compared[0] = cmp[0];
        r.compared[1] = cmp[1];
        r.cls = (uint8_t)cls;
        r.status[0] = (uint8_t)st[0];
        r.status[1] = (uint8_t)st[1];
        r.exact = (uint8_t)exact;
        r.reserved = 0;
        P.per_cand[i] = r;
    }
}

template <bool HAS_VOID>
__global__ void __launch_bounds__(HC_WARPS_MAX * 32, 2) hc_score_kernel(const hc_kparams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* T = reinterpret_cast<uint32_t*>(smem);
    const uint32_t tbl_entries = (P.ncodes + 1u) * 256u;
    for (uint32_t i = threadIdx.x; i < tbl_entries; i += blockDim.x) T[i] = P.fx_table[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    unsigned char* scratch = smem + (size_t)tbl_entries * 4u + (size_t)warp * HC_WARP_SCRATCH;
    uint2* part = reinterpret_cast<uint2*>(scratch);                                     // [HC_PARTMAX]
    uint4* wdA = reinterpret_cast<uint4*>(scratch + HC_PARTMAX * 8u);                    // [64] xpos lo/hi, ypos16, L
    uint2* wdB = reinterpret_cast<uint2*>(scratch + HC_PARTMAX * 8u + HC_WINSLOTS * 16u); // [64] start chunk, hasN
    uint32_t* head = reinterpret_cast<uint32_t*>(scratch + HC_PARTMAX * 8u + HC_WINSLOTS * 24u);  // [HC_PARTMAX/32]

    const u64 ntiles = (P.n + 31ull) >> 5;
    const u64 gwarp = (u64)blockIdx.x * nwarps + warp;
    const u64 twarps = (u64)gridDim.x * nwarps;
    const uint32_t lane_le = 0xffffffffu >> (31 - lane);

    u64 st_windows = 0, st_positions = 0, st_bytes = 0;
    uint32_t st_errors = 0;

    for (u64 tile = gwarp; tile < ntiles; tile += twarps) {
        const u64 i = (tile << 5) + lane;
        const bool valid = i < P.n;
        CandSetup s;
        s.err = 0; s.two = 0; s.pos3 = s.pos4 = 0;
        s.w[0].L = s.w[1].L = 0; s.w[0].status = s.w[1].status = HC_WIN_UNUSED;
        s.w[0].xpos = s.w[1].xpos = 0; s.w[0].ypos16 = s.w[1].ypos16 = 0; s.w[0].hasN = s.w[1].hasN = 0;
        if (valid) {
            const uint4* cp = reinterpret_cast<const uint4*>(P.cand + i);
            const uint4 ca = __ldg(cp), cb = __ldg(cp + 1);
            hc_candidate c;
            c.idx1 = ca.x; c.idx2 = ca.y; c.pos1 = ca.z; c.pos2 = ca.w;
            c.len1 = cb.x; c.len2 = cb.y;
            c.perc1 = cb.z & 0xff; c.perc2 = (cb.z >> 8) & 0xff; c.ord = (cb.z >> 16) & 0xff; c.ori1 = (cb.z >> 24) & 0xff;
            c.ori2 = cb.w & 0xff; c.type1 = (cb.w >> 8) & 0xff; c.type2 = (cb.w >> 16) & 0xff; c.reserved = 0;
            setup_candidate(P, c, s);
        }
        const uint32_t c0 = (s.w[0].L + 15u) >> 4, c1 = (s.w[1].L + 15u) >> 4;
        const uint32_t ct = c0 + c1;
        WinAcc acc[2];
        acc[0].S = acc[1].S = 0; acc[0].mm = acc[1].mm = 0; acc[0].nn = acc[1].nn = 0; acc[0].vd = acc[1].vd = 0;

        // ---- big candidates (>= 64 chunks): the whole warp walks one window at a time
        uint32_t bigmask = __ballot_sync(0xffffffffu, ct >= HC_BIG_CHUNKS);
        while (bigmask) {
            const int src = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
#pragma unroll
            for (int w = 0; w < 2; w++) {
                const uint32_t xl = __shfl_sync(0xffffffffu, (uint32_t)s.w[w].xpos, src);
                const uint32_t xh = __shfl_sync(0xffffffffu, (uint32_t)(s.w[w].xpos >> 32), src);
                const uint32_t yp = __shfl_sync(0xffffffffu, s.w[w].ypos16, src);
                const uint32_t Lw = __shfl_sync(0xffffffffu, s.w[w].L, src);
                const uint32_t hn = __shfl_sync(0xffffffffu, s.w[w].hasN, src);
                const uint32_t cw = (Lw + 15u) >> 4;
                if (cw == 0) continue;
                const u64 xpos = ((u64)xh << 32) | xl;
                u64 S = 0;
                uint32_t mm = 0, nn = 0, vd = 0;
                for (uint32_t k = lane; k < cw; k += 32) {
                    uint32_t sum, m1, n1, v1;
                    process_chunk<HAS_VOID>(P, T, xpos, yp, Lw, hn, k, sum, m1, n1, v1);
                    S += sum; mm += m1; nn += n1; vd |= v1;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    S += __shfl_xor_sync(0xffffffffu, S, d);
                    mm += __shfl_xor_sync(0xffffffffu, mm, d);
                    nn += __shfl_xor_sync(0xffffffffu, nn, d);
                    vd |= __shfl_xor_sync(0xffffffffu, vd, d);
                }
                if (lane == src) { acc[w].S = S; acc[w].mm = mm; acc[w].nn = nn; acc[w].vd = vd; }
            }
        }

        // ---- small candidates: rounds of at most HC_PARTMAX chunks dealt round robin to the lanes
        uint32_t pending = __ballot_sync(0xffffffffu, ct > 0 && ct < HC_BIG_CHUNKS);
        while (pending) {
            const bool mine = (pending >> lane) & 1u;
            const uint32_t nwin = (c0 > 0) + (c1 > 0);
            const uint32_t v = mine ? (ct | (nwin << 16)) : 0u;
            const uint32_t incl = warp_incl_scan(v, lane);
            const bool fit = mine && ((incl & 0xffffu) <= HC_PARTMAX);
            const uint32_t take = __ballot_sync(0xffffffffu, fit);   // a prefix of the pending lanes, never empty
            const int last = 31 - __clz(take);
            const uint32_t C = __shfl_sync(0xffffffffu, incl, last) & 0xffffu;
            if (lane < (int)(HC_PARTMAX / 32u)) head[lane] = 0;
            __syncwarp();
            uint32_t start0 = 0, start1 = 0;
            if (fit) {
                start0 = (incl & 0xffffu) - ct;
                start1 = start0 + c0;
                uint32_t slot = (incl >> 16) - nwin;
                if (c0) {
                    wdA[slot] = make_uint4((uint32_t)s.w[0].xpos, (uint32_t)(s.w[0].xpos >> 32), s.w[0].ypos16, s.w[0].L);
                    wdB[slot] = make_uint2(start0, s.w[0].hasN);
                    atomicOr(&head[start0 >> 5], 1u << (start0 & 31u));
                    slot++;
                }
                if (c1) {
                    wdA[slot] = make_uint4((uint32_t)s.w[1].xpos, (uint32_t)(s.w[1].xpos >> 32), s.w[1].ypos16, s.w[1].L);
                    wdB[slot] = make_uint2(start1, s.w[1].hasN);
                    atomicOr(&head[start1 >> 5], 1u << (start1 & 31u));
                }
            }
            __syncwarp();
            uint32_t running = 0;
            for (uint32_t f0 = 0; f0 < C; f0 += 32) {
                const uint32_t hw = head[f0 >> 5];
                const uint32_t f = f0 + lane;
                if (f < C) {
                    const uint32_t slot = running + __popc(hw & lane_le) - 1u;
                    const uint4 a = wdA[slot];
                    const uint2 b = wdB[slot];
                    uint32_t sum, m1, n1, v1;
                    process_chunk<HAS_VOID>(P, T, ((u64)a.y << 32) | a.x, a.z, a.w, b.y, f - b.x, sum, m1, n1, v1);
                    part[f] = make_uint2(sum, m1 | (n1 << 12) | (v1 << 24));
                }
                running += __popc(hw);
            }
            __syncwarp();
            if (fit) {
                u64 S = 0;
                uint32_t pk = 0;
                for (uint32_t k = 0; k < c0; k++) { const uint2 e = part[start0 + k]; S += e.x; pk += e.y; }
                acc[0].S = S; acc[0].mm = pk & 0xfffu; acc[0].nn = (pk >> 12) & 0xfffu; acc[0].vd = pk >> 24;
                S = 0; pk = 0;
                for (uint32_t k = 0; k < c1; k++) { const uint2 e = part[start1 + k]; S += e.x; pk += e.y; }
                acc[1].S = S; acc[1].mm = pk & 0xfffu; acc[1].nn = (pk >> 12) & 0xfffu; acc[1].vd = pk >> 24;
            }
            __syncwarp();
            pending &= ~take;
        }

        // ---- decision (lane = candidate)
        bool flag = false;
        if (valid) {
            if (s.err) {
                st_errors++;
                const uint32_t z[2] = {0, 0};
                const uint32_t stt[2] = {HC_WIN_UNUSED, HC_WIN_UNUSED};
                write_result(P, i, s, 0.0, 1.0, HC_CLASS_DISCARD, z, z, stt, 0);
            } else {
                double ov[2] = {0.0, 0.0}, mmr[2] = {1.0, 1.0};
                int ae[2], ao[2];
                uint32_t mmc[2] = {0, 0}, cmp[2] = {0, 0}, stt[2];
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    uint32_t status = s.w[w].status;
                    ae[w] = P.zero_above_edge;
                    ao[w] = P.zero_above_ov;
                    if (status == HC_WIN_SCORED) {
                        const uint32_t tl = s.w[w].L - acc[w].nn;
                        mmc[w] = acc[w].mm;
                        if (acc[w].vd) {
                            status = HC_WIN_VOID;               // :125-127, mismatch_rate stays 1.0
                            mmc[w] = 0;                         // counts of a void window are not observable
                        } else if (tl == 0) {
                            status = HC_WIN_EMPTY;              // :129-131
                        } else {
                            cmp[w] = tl;
                            const double dl = (double)tl;
                            mmr[w] = (double)(float)(int)acc[w].mm / dl;                       // :132
                            const double mean = -((double)acc[w].S * (1.0 / HC_FX_SCALE)) / dl;  // :137 (fixed point)
                            ov[w] = exp(mean);                                                  // :138
                            const bool never_e = P.t_edge > 0.0, never_o = P.t_ov > 0.0;       // mean <= 0 always
                            const bool up_e = !never_e && (mean - HC_FX_MARGIN >= P.t_edge);
                            const bool dn_e = never_e || (mean + HC_FX_MARGIN < P.t_edge);
                            const bool up_o = !never_o && (mean - HC_FX_MARGIN >= P.t_ov);
                            const bool dn_o = never_o || (mean + HC_FX_MARGIN < P.t_ov);
                            ae[w] = up_e;
                            ao[w] = up_o;
                            if (!(up_e || dn_e) || !(up_o || dn_o)) flag = true;
                        }
                        st_windows++;
                        st_positions += s.w[w].L;
                        st_bytes += 2ull * ((s.w[w].L + 3u) >> 2) + 2ull * ((s.w[w].L + 7u) >> 3) + 2ull * s.w[w].L;
                    }
                    stt[w] = status;
                }
                st_bytes += 48;
                double score, mmrate;
                uint32_t cls;
                combine(P, s.two, ov, mmr, ae, ao, score, mmrate, cls);
                if (P.exact_edges && cls == HC_CLASS_EDGE) flag = true;
                write_result(P, i, s, score, mmrate, cls, mmc, cmp, stt, 0);
            }
        }
        // queue boundary cases for the reference-order pass (warp-aggregated append)
        const uint32_t fm = __ballot_sync(0xffffffffu, flag);
        if (fm) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P.counters[HC_CNT_FLAGGED], (unsigned long long)__popc(fm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (flag) P.flagged[base + __popc(fm & (lane_le >> 1))] = (uint32_t)i;
        }
    }
    // per-warp statistics, one atomic each at the very end
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        st_windows += __shfl_xor_sync(0xffffffffu, st_windows, d);
        st_positions += __shfl_xor_sync(0xffffffffu, st_positions, d);
        st_bytes += __shfl_xor_sync(0xffffffffu, st_bytes, d);
        st_errors += __shfl_xor_sync(0xffffffffu, st_errors, d);
    }
    if (lane == 0) {
        atomicAdd(&P.counters[HC_CNT_WINDOWS], st_windows);
        atomicAdd(&P.counters[HC_CNT_POSITIONS], st_positions);
        atomicAdd(&P.counters[HC_CNT_ALGBYTES], st_bytes);
        if (st_errors) atomicAdd(&P.counters[HC_CNT_ERRORS], (unsigned long long)st_errors);
    }
}

// ---- reference-order pass ---------------------------------------------------------------------------
// One thread re-adds one window exactly like src/EdgeCalculator.cpp:103-138: same double addends
// (host-built with the reference's expressions and libm), same order, IEEE add/mul/div with
// explicit _rn intrinsics so nothing is contracted into an FMA.
__device__ void exact_window(const hc_kparams& P, const Win& w, double& mean, double& mmrate, uint32_t& mmc,
                             uint32_t& cmp, uint32_t& status) {
    mean = 0.0;
    mmrate = 1.0;
    mmc = 0;
    cmp = 0;
    status = w.status;
    if (w.status != HC_WIN_SCORED) return;
    const uint32_t n1 = P.ncodes + 1u;
    double total = 0.0;
    uint32_t tl = 0, mm = 0;
    const u64 yp = 16ull * w.ypos16;
    for (uint32_t i = 0; i < w.L; i++) {
        const u64 xa = w.xpos + i, xb = yp + i;
        const uint32_t nA = (P.nmask[xa >> 5] >> (xa & 31)) & 1u, nB = (P.nmask[xb >> 5] >> (xb & 31)) & 1u;
        if (nA | nB) continue;                                            // :35-39,:122-124
        const uint32_t a = (P.base2[xa >> 4] >> (2 * (xa & 15))) & 3u, b = (P.base2[xb >> 4] >> (2 * (xb & 15))) & 3u;
        const uint32_t qa = P.qual[xa], qb = P.qual[xb];
        const uint32_t mis = a != b;
        mm += mis;
        const double lp = P.dbl_table[hc_dbl_index(qa, qb, mis, n1)];
        if (lp > 0.0) { status = HC_WIN_VOID; mmc = 0; return; }        // :125-127
        total = __dadd_rn(total, lp);                                     // :119
        tl++;
    }
    mmc = mm;
    if (tl == 0) { status = HC_WIN_EMPTY; return; }                      // :129-131
    cmp = tl;
    const double dl = (double)tl;
    mmrate = __ddiv_rn((double)(float)(int)mm, dl);                      // :132
    mean = __dmul_rn(__ddiv_rn(1.0, dl), total);                         // :137
}

__global__ void hc_exact_kernel(const hc_kparams P) {
    const u64 nf = P.counters[HC_CNT_FLAGGED];
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nf; t += (u64)gridDim.x * blockDim.x) {
        const u64 i = P.flagged[t];
        const hc_candidate c = P.cand[i];
        CandSetup s;
        setup_candidate(P, c, s);
        if (s.err) continue;
        double ov[2] = {0.0, 0.0}, mmr[2] = {1.0, 1.0};
        int ae[2], ao[2];
        uint32_t mmc[2], cmp[2], stt[2];
#pragma unroll
        for (int w = 0; w < 2; w++) {
            double mean;
            exact_window(P, s.w[w], mean, mmr[w], mmc[w], cmp[w], stt[w]);
            if (stt[w] == HC_WIN_SCORED) {
                ov[w] = exp(mean);
                ae[w] = mean >= P.t_edge;   // <=> host-libm exp(mean) > edge_threshold
                ao[w] = mean >= P.t_ov;
            } else {
                ae[w] = P.zero_above_edge;
                ao[w] = P.zero_above_ov;
            }
        }
        double score, mmrate;
        uint32_t cls;
        combine(P, s.two, ov, mmr, ae, ao, score, mmrate, cls);
        write_result(P, i, s, score, mmrate, cls, mmc, cmp, stt, 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) P.counters[HC_CNT_EXACT] = nf;
}

// ---- order-preserving compaction ----------------------------------------------------------------------
#define HC_CB_THREADS 256
#define HC_CB_ITEMS 4096   // candidates per block

__global__ void __launch_bounds__(HC_CB_THREADS) hc_compact_count(const uint8_t* __restrict__ cls, u64 n, uint32_t* blockcounts) {
    const u64 base = (u64)blockIdx.x * HC_CB_ITEMS;
    uint32_t e = 0, o = 0;
    for (uint32_t k = threadIdx.x; k < HC_CB_ITEMS; k += HC_CB_THREADS) {
        const u64 i = base + k;
        if (i < n) {
            const uint32_t c = cls[i];
            e += c == HC_CLASS_EDGE;
            o += c == HC_CLASS_NONEDGE;
        }
    }
    __shared__ uint32_t se[HC_CB_THREADS / 32], so[HC_CB_THREADS / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, d);
        o += __shfl_xor_sync(0xffffffffu, o, d);
    }
    if ((threadIdx.x & 31) == 0) { se[threadIdx.x >> 5] = e; so[threadIdx.x >> 5] = o; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t te = 0, to = 0;
        for (int w = 0; w < HC_CB_THREADS / 32; w++) { te += se[w]; to += so[w]; }
        blockcounts[2 * blockIdx.x] = te;
        blockcounts[2 * blockIdx.x + 1] = to;
    }
}

// single block: exclusive scan of the per-block counts (64-bit running totals kept in blockoffs)
__global__ void __launch_bounds__(1024) hc_compact_scan(const uint32_t* blockcounts, uint32_t nblocks, u64* blockoffs,
                                                        unsigned long long* counters) {
    __shared__ u64 wsum_e[32], wsum_o[32];
    __shared__ u64 carry_e, carry_o;
    if (threadIdx.x == 0) { carry_e = 0; carry_o = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024) {
        const uint32_t b = b0 + threadIdx.x;
        u64 e = b < nblocks ? blockcounts[2 * b] : 0, o = b < nblocks ? blockcounts[2 * b + 1] : 0;
        u64 ie = e, io = o;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 te = __shfl_up_sync(0xffffffffu, ie, d), to = __shfl_up_sync(0xffffffffu, io, d);
            if (lane >= d) { ie += te; io += to; }
        }
        if (lane == 31) { wsum_e[warp] = ie; wsum_o[warp] = io; }
        __syncthreads();
        if (warp == 0) {
            u64 ve = wsum_e[lane], vo = wsum_o[lane];
            u64 se = ve, so = vo;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 te = __shfl_up_sync(0xffffffffu, se, d), to = __shfl_up_sync(0xffffffffu, so, d);
                if (lane >= d) { se += te; so += to; }
            }
            wsum_e[lane] = se - ve;
            wsum_o[lane] = so - vo;
        }
        __syncthreads();
        const u64 ce = carry_e, co = carry_o;
        if (b < nblocks) {
            blockoffs[2 * b] = ce + wsum_e[warp] + ie - e;
            blockoffs[2 * b + 1] = co + wsum_o[warp] + io - o;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_e = ce + wsum_e[31] + ie; carry_o = co + wsum_o[31] + io; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { counters[HC_CNT_EDGES] = carry_e; counters[HC_CNT_NONEDGES] = carry_o; }
}

__global__ void __launch_bounds__(HC_CB_THREADS) hc_compact_scatter(const hc_kparams P, const u64* __restrict__ blockoffs,
                                                                   hc_edge* edges, u64 edges_cap, uint64_t* nonedge,
                                                                   u64 nonedge_cap, u64 cand_offset) {
    __shared__ uint32_t we[HC_CB_THREADS / 32], wo[HC_CB_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 base = (u64)blockIdx.x * HC_CB_ITEMS;
    u64 eoff = blockoffs[2 * blockIdx.x], ooff = blockoffs[2 * blockIdx.x + 1];
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t k0 = 0; k0 < HC_CB_ITEMS; k0 += HC_CB_THREADS) {
        const u64 i = base + k0 + threadIdx.x;
        const uint32_t c = i < P.n ? P.cls[i] : HC_CLASS_DISCARD;
        const uint32_t be = __ballot_sync(0xffffffffu, c == HC_CLASS_EDGE);
        const uint32_t bo = __ballot_sync(0xffffffffu, c == HC_CLASS_NONEDGE);
        if (lane == 0) { we[warp] = __popc(be); wo[warp] = __popc(bo); }
        __syncthreads();
        uint32_t pe = 0, po = 0, te = 0, to = 0;
#pragma unroll
        for (int w = 0; w < HC_CB_THREADS / 32; w++) {
            if (w < warp) { pe += we[w]; po += wo[w]; }
            te += we[w];
            to += wo[w];
        }
        if (c == HC_CLASS_EDGE) {
            const u64 dst = eoff + pe + __popc(be & lt);
            if (dst < edges_cap) {
                const hc_candidate cd = P.cand[i];
                CandSetup s;
                setup_candidate(P, cd, s);
                const hc_score16 t = P.tmp[i];
                hc_edge e;
                e.cand = i + cand_offset;
                e.score = t.score;
                e.mismatch_rate = t.mismatch_rate;
                e.pos3 = s.pos3;
                e.pos4 = s.pos4;
                edges[dst] = e;
            }
        } else if (c == HC_CLASS_NONEDGE) {
            const u64 dst = ooff + po + __popc(bo & lt);
            if (dst < nonedge_cap) nonedge[dst] = i + cand_offset;
        }
        eoff += te;
        ooff += to;
        __syncthreads();
    }
}

}  // namespace

// ---- launchers ----------------------------------------------------------------------------------------
cudaError_t hc_score_occupancy(uint32_t ncodes, int sm_count, size_t smem_per_sm, hc_launch_cfg* cfg) {
    const size_t table = (size_t)(ncodes + 1u) * 1024u;
    int best_warps = 0, best_nw = 0, best_ctas = 0;
    const int options[2] = {HC_WARPS_MAX, 8};
    for (int o = 0; o < 2; o++) {
        const int nw = options[o];
        const size_t per_cta = table + (size_t)nw * HC_WARP_SCRATCH + 1024u;   // +1 KB the driver reserves per CTA
        if (per_cta > 227u * 1024u) continue;
        int ctas = (int)(smem_per_sm / per_cta);
        if (ctas * nw * 32 > 2048) ctas = 2048 / (nw * 32);
        if (ctas < 1) continue;
        if (ctas * nw > best_warps) { best_warps = ctas * nw; best_nw = nw; best_ctas = ctas; }
    }
    if (best_nw == 0) return cudaErrorInvalidConfiguration;
    cfg->threads = best_nw * 32;
    cfg->smem = table + (size_t)best_nw * HC_WARP_SCRATCH;
    cfg->blocks = sm_count * best_ctas;
    return cudaSuccess;
}

cudaError_t hc_launch_score(const hc_kparams& P, const hc_launch_cfg& cfg, cudaStream_t st) {
    cudaError_t e;
    if (P.has_void) {
        e = cudaFuncSetAttribute(hc_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
        if (e != cudaSuccess) return e;
        hc_score_kernel<true><<<cfg.blocks, cfg.threads, cfg.smem, st>>>(P);
    } else {
        e = cudaFuncSetAttribute(hc_score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
        if (e != cudaSuccess) return e;
        hc_score_kernel<false><<<cfg.blocks, cfg.threads, cfg.smem, st>>>(P);
    }
    return cudaGetLastError();
}

cudaError_t hc_launch_exact(const hc_kparams& P, cudaStream_t st) {
    hc_exact_kernel<<<296, 128, 0, st>>>(P);
    return cudaGetLastError();
}

uint32_t hc_compact_blocks(uint64_t n) { return (uint32_t)((n + HC_CB_ITEMS - 1) / HC_CB_ITEMS); }

// d_blockcounts: uint32[2*nblocks] followed (8-byte aligned) by uint64[2*nblocks] block offsets
cudaError_t hc_launch_compact(const hc_kparams& P, hc_edge* d_edges, uint64_t edges_cap, uint64_t* d_nonedge,
                              uint64_t nonedge_cap, uint32_t* d_blockcounts, uint64_t cand_offset, cudaStream_t st) {
    const uint32_t nb = hc_compact_blocks(P.n);
    u64* offs = reinterpret_cast<u64*>(d_blockcounts + 2ull * nb + (2ull * nb & 1ull));
    if (nb > 0) {
        hc_compact_count<<<nb, HC_CB_THREADS, 0, st>>>(P.cls, P.n, d_blockcounts);
    }
    hc_compact_scan<<<1, 1024, 0, st>>>(d_blockcounts, nb, offs, P.counters);
    if (nb > 0) {
        hc_compact_scatter<<<nb, HC_CB_THREADS, 0, st>>>(P, offs, d_edges, edges_cap, d_nonedge, nonedge_cap, cand_offset);
    }
    return cudaGetLastError();
}
