// hc_fno.cu -- FindNextOverlaps (FNO1) on the device: SRBuilder::findNextOverlaps,
// src/FindNextOverlaps.cpp:890-958 (updateOverlap :25-327, findCliqueIndex :331-347,
// computeOverlapData :351-565).
//
// The reference walks the edge stream sequentially; for every edge (u, v) it tries every pair
// (new read of u) x (new read of v) and keeps, per unordered pair of new reads, the FIRST attempt in
// processing order -- whether or not that attempt then succeeds (:84-97 precede :115-118).  Here:
//   1. fno_count      attempts per edge                       -> exclusive scan = sequence numbers
//   2. fno_claim      every keyed attempt does atomicMin(sequence number) on its pair's hash slot
//   3. fno_resolve    an attempt survives if it is a plain copy (:46-72) or holds its pair's minimum,
//                     and computeOverlapData succeeds; its record is written at its sequence number
//                                                                -> exclusive scan of the flags = output positions
//   4. fno_compact    survivors are moved to their positions: processing order
// All integer / float32 arithmetic, bit-identical to the reference (perc uses IEEE float division,
// max, multiplication and floor, :375,:429,:487,:549).
#include <cstdio>
#include <cstdlib>
#include <time.h>
#include <cstring>
#include <string>
#include <cuda_runtime.h>
#include "../../include/hc_b200.h"
#include "hc_scan.cuh"
#include "hc_stage.h"

#ifndef HC_FNO_CELL
#define HC_FNO_CELL 1   // 1: keys and minima in two arrays; 2: {key, min} cells of 16 bytes -- measured slower (claim 19.8 vs 17.1 ms, resolve 15.9 vs 12.3 ms)
#endif

namespace {

typedef unsigned long long u64;

struct FnoDev {
    u64 n_vertices;
    const uint8_t* visited;
    const uint8_t* label;
    const hc_fno_read* vertex_read;
    const u64* sr_off;
    const uint32_t* sr_idx;
    const hc_fno_subread* sr_sub;
    const hc_fno_read* superread;
    uint32_t resolve_orientations, no_inclusions;
};

__device__ __forceinline__ int perc_of(int t, int la, int lb) {
    const float a = __fdiv_rn((float)t, (float)la), b = __fdiv_rn((float)t, (float)lb);
    return (int)floorf(__fmul_rn(fmaxf(a, b), 100.0f));
}

struct Derived {
    int pos1, pos2, perc, ol1, ol2;
    char ord1, ord2, t1, t2;
};

// src/FindNextOverlaps.cpp:351-565
__device__ bool compute_overlap_data(const hc_fno_read& r1, const hc_fno_read& r2, int idx1l, int idx1r, int idx2l, int idx2r,
                                     const hc_fno_edge& e, Derived& d) {
    const int pos1 = e.pos1, pos2 = e.pos2;
    const bool p1 = r1.len2 > 0, p2 = r2.len2 > 0;
    const int l11 = (int)r1.len1, l12 = (int)r1.len2, l21 = (int)r2.len1, l22 = (int)r2.len2;
    d.pos2 = 0;
    d.pos1 = (pos1 + idx1l) - idx2l;
    if (!p1 && !p2) {                                              // S-S :357-385
        d.t1 = 's'; d.t2 = 's';
        int len;
        if (d.pos1 < 0) { d.ord1 = '2'; d.pos1 = -d.pos1; len = l21; }
        else { d.ord1 = '1'; len = l11; }
        d.ol1 = min(min(len - d.pos1, l11), l21);
        d.ol2 = 0;
        d.perc = perc_of(d.ol1, l11, l21);
        d.ord2 = '-';
        return d.pos1 < len;
    }
    if (p1 && !p2) {                                               // P-S :387-443
        d.t1 = 'p'; d.t2 = 's';
        if (d.pos1 < 0) {
            d.ord1 = '2'; d.pos1 = -d.pos1;
            if (d.pos1 >= l21) return false;
            d.ol1 = l11;
        } else {
            d.ord1 = '1';
            if (d.pos1 >= l11) return false;
            d.ol1 = l11 - d.pos1;
        }
        d.pos2 = e.ord == '1' ? idx2r - (idx1r + pos2) : (pos2 + idx2r) - idx1r;
        if (d.pos2 >= l21 || d.pos2 < 0) return false;
        d.ord2 = '-';
        d.ol2 = min(l21 - d.pos2, l12);
        d.perc = min(perc_of(d.ol1 + d.ol2, l11 + l12, l21), 100);
        return true;
    }
    if (!p1 && p2) {                                               // S-P :445-489
        d.t1 = 's'; d.t2 = 'p';
        if (d.pos1 < 0) {
            d.ord1 = '2'; d.pos1 = -d.pos1;
            if (d.pos1 >= l21) return false;
            d.ol1 = l21 - d.pos1;
        } else {
            d.ord1 = '1';
            if (d.pos1 >= l11) return false;
            d.ol1 = l21;
        }
        d.pos2 = e.ord == '2' ? idx1r - (pos2 + idx2r) : idx1r + pos2 - idx2r;
        if (d.pos2 >= l11 || d.pos2 < 0) return false;
        d.ord2 = '-';
        d.ol2 = min(l11 - d.pos2, l22);
        d.perc = min(perc_of(d.ol1 + d.ol2, l11, l21 + l22), 100);
        return true;
    }
    d.t1 = 'p'; d.t2 = 'p';                                        // P-P :491-551
    if (d.pos1 < 0) {
        d.ord1 = '2'; d.pos1 = -d.pos1;
        if (d.pos1 >= l21) return false;
        d.ol1 = min(l11, l21 - d.pos1);
    } else {
        d.ord1 = '1';
        if (d.pos1 >= l11) return false;
        d.ol1 = min(l11 - d.pos1, l21);
    }
    d.pos2 = e.ord == '1' ? (pos2 + idx1r) - idx2r : idx1r - (pos2 + idx2r);
    if (d.pos2 < 0) {
        d.ord2 = d.ord1 == '1' ? '2' : '1';
        d.pos2 = -d.pos2;
        if (d.pos2 >= l22) return false;
        d.ol2 = min(l12, l22 - d.pos2);
    } else {
        d.ord2 = d.ord1 == '1' ? '1' : '2';
        if (d.pos2 >= l12) return false;
        d.ol2 = min(l12 - d.pos2, l22);
    }
    d.perc = min(perc_of(d.ol1 + d.ol2, l11 + l12, l21 + l22), 100);
    return true;
}

__device__ __forceinline__ u64 hash_slot(u64 key, u64 mask) { return (key * 0x9E3779B97F4A7C15ull) & mask; }

// Iterates the attempts of one edge in the reference's order and calls f(seq, s1, s2, a, b, keyed).
template <class F>
__device__ __forceinline__ void for_each_attempt(const FnoDev& D, const hc_fno_edge& e, u64 seq0, F f) {
    const bool vu = D.visited[e.u], vv = D.visited[e.v];
    const u64 a0 = vu ? D.sr_off[e.u] : 0, a1 = vu ? D.sr_off[e.u + 1] : 1;
    const u64 b0 = vv ? D.sr_off[e.v] : 0, b1 = vv ? D.sr_off[e.v + 1] : 1;
    u64 seq = seq0;
    for (u64 a = a0; a < a1; a++)
        for (u64 b = b0; b < b1; b++, seq++) f(seq, a, b, vu, vv);
}

__global__ void fno_count(FnoDev D, const hc_fno_edge* edges, u64 n, uint32_t* cnt) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const hc_fno_edge e = edges[i];
        const u64 a = D.visited[e.u] ? D.sr_off[e.u + 1] - D.sr_off[e.u] : 1;
        const u64 b = D.visited[e.v] ? D.sr_off[e.v + 1] - D.sr_off[e.v] : 1;
        cnt[i] = (uint32_t)(a * b);
    }
}

__global__ void fno_claim(FnoDev D, const hc_fno_edge* edges, u64 n, const u64* off, u64* keys, u64* mins, u64 mask) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const hc_fno_edge e = edges[i];
        if (!D.visited[e.u] && !D.visited[e.v]) continue;          // plain copy, no first-found bookkeeping (:46-72)
        for_each_attempt(D, e, off[i], [&](u64 seq, u64 a, u64 b, bool vu, bool vv) {
            const u64 id1 = vu ? D.superread[D.sr_idx[a]].id : D.vertex_read[e.u].id;
            const u64 id2 = vv ? D.superread[D.sr_idx[b]].id : D.vertex_read[e.v].id;
            if (id1 == id2) return;                                 // :241-243
            const u64 key = (min(id1, id2) << 32) | max(id1, id2);
            u64 h = hash_slot(key, mask);
            while (true) {
                const u64 prev = atomicCAS(&keys[HC_FNO_CELL * h], ~0ull, key);
                if (prev == ~0ull || prev == key) break;
                h = (h + 1) & mask;
            }
            atomicMin(&mins[HC_FNO_CELL * h], seq);
        });
    }
}

// One pass: flags[seq] = "attempt seq yields an overlap" and, when it does, its record at rec[seq] (attempt order);
// fno_compact then moves the records to their ranks.  (Deriving everything a second time for the emission cost as
// much as the first pass: the gathers through vertex -> super-read lists are what these kernels spend their time on.)
__global__ void fno_resolve(FnoDev D, const hc_fno_edge* edges, u64 n, const u64* off, const u64* keys, const u64* mins, u64 mask,
                            uint32_t* flags, hc_fno_overlap* rec) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const hc_fno_edge e = edges[i];
        char ori1 = '+', ori2 = '+';
        if (D.resolve_orientations && e.nonedge) {                  // :34-37
            ori1 = (e.ori1 == D.label[e.u]) ? '+' : '-';
            ori2 = (e.ori2 == D.label[e.v]) ? '+' : '-';
        }
        const hc_fno_read ru = D.vertex_read[e.u], rv = D.vertex_read[e.v];
        const bool pu = ru.len2 > 0, pv = rv.len2 > 0;
        if (!D.visited[e.u] && !D.visited[e.v]) {                   // :46-72
            const u64 seq = off[i];
            const bool ok = !(D.no_inclusions && e.perc == 100);
            flags[seq] = ok;
            if (ok) {
                hc_fno_overlap o;
                memset(&o, 0, sizeof(o));
                o.id1 = ru.id; o.id2 = rv.id; o.pos1 = e.pos1; o.pos2 = e.pos2; o.ord = e.ord; o.ori1 = ori1; o.ori2 = ori2;
                o.perc = e.perc; o.len1 = e.len1; o.len2 = e.len2; o.type1 = pu ? 'p' : 's'; o.type2 = pv ? 'p' : 's';
                rec[seq] = o;
            }
            continue;
        }
        for_each_attempt(D, e, off[i], [&](u64 seq, u64 a, u64 b, bool vu, bool vv) {
            const hc_fno_read s1 = vu ? D.superread[D.sr_idx[a]] : ru;
            const hc_fno_read s2 = vv ? D.superread[D.sr_idx[b]] : rv;
            bool ok = s1.id != s2.id;
            Derived d;
            if (ok) {
                const u64 key = (min(s1.id, s2.id) << 32) | max(s1.id, s2.id);
                u64 h = hash_slot(key, mask);
                while (keys[HC_FNO_CELL * h] != key) h = (h + 1) & mask;
                ok = mins[HC_FNO_CELL * h] == seq;                                // first found wins, even if it fails below
            }
            if (ok) {
                int i1l = 0, i1r = 0, i2l = 0, i2r = 0;
                if (vu) {                                           // findCliqueIndex :331-347
                    const hc_fno_subread s = D.sr_sub[a];
                    i1l = s.index1 - s.startpos1;
                    i1r = (s1.len2 > 0 || pu) ? s.index2 - s.startpos2 : i1l;
                }
                if (vv) {
                    const hc_fno_subread s = D.sr_sub[b];
                    i2l = s.index1 - s.startpos1;
                    i2r = (s2.len2 > 0 || pv) ? s.index2 - s.startpos2 : i2l;
                }
                ok = compute_overlap_data(s1, s2, i1l, i1r, i2l, i2r, e, d);
                if (ok && D.no_inclusions && d.perc == 100) ok = false;
            }
            flags[seq] = ok;
            if (ok) {
                hc_fno_overlap o;
                memset(&o, 0, sizeof(o));
                if (d.ord1 == '1') { o.id1 = s1.id; o.id2 = s2.id; o.type1 = d.t1; o.type2 = d.t2; }
                else { o.id1 = s2.id; o.id2 = s1.id; o.type1 = d.t2; o.type2 = d.t1; }
                o.pos1 = d.pos1; o.pos2 = d.pos2; o.ord = d.ord2; o.ori1 = ori1; o.ori2 = ori2;
                o.perc = d.perc; o.len1 = d.ol1; o.len2 = d.ol2;
                rec[seq] = o;
            }
        });
    }
}

// records of the successful attempts, attempt order -> output order (ranks from the scan of the flags)
__global__ void fno_compact(const uint32_t* __restrict__ flags, const u64* __restrict__ outpos, const hc_fno_overlap* __restrict__ rec,
                            u64 attempts, hc_fno_overlap* __restrict__ out, u64 out_cap) {
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < attempts; t += (u64)gridDim.x * blockDim.x)
        if (flags[t] && outpos[t] < out_cap) out[outpos[t]] = rec[t];
}


// ---- FindNextOverlaps3 (src/FindNextOverlaps3.cpp:90-406) ------------------------------------------------
__device__ __forceinline__ int perc_one(int l, int a) { return (int)floorf(__fmul_rn(__fdiv_rn((float)l, (float)a), 100.0f)); }

// deduceOverlap :176-406; false = "this overlap will be ignored"
__device__ bool deduce_overlap(const hc_fno_read& A, const hc_fno_read& B, const hc_fno3_pos& pa, const hc_fno3_pos& pb,
                               hc_fno_overlap& o) {
    const bool pA = A.len2 > 0, pB = B.len2 > 0;
    memset(&o, 0, sizeof(o));
    o.ori1 = '+'; o.ori2 = '+'; o.ord = '-';
    const int i1l = pa.index1, i1r = pa.index2, i2l = pb.index1, i2r = pb.index2;
    const bool a_first = i1l - i2l >= 0;
    o.id1 = a_first ? A.id : B.id;
    o.id2 = a_first ? B.id : A.id;
    o.pos1 = a_first ? i1l - i2l : i2l - i1l;
    if (!pA && !pB) {                                                       // S-S :202-243
        const int lenA = (int)A.len1, lenB = (int)B.len1;
        if (o.pos1 > (a_first ? lenA : lenB)) return false;
        o.len1 = a_first ? min(lenA - o.pos1, lenB) : min(lenA, lenB - o.pos1);
        o.perc = perc_of(o.len1, lenA, lenB);
        o.type1 = 's'; o.type2 = 's';
        return true;
    }
    if (pA && !pB) {                                                        // P-S :244-287
        const int lenA1 = (int)A.len1, lenA2 = (int)A.len2, lenB = (int)B.len1;
        o.len1 = a_first ? lenA1 - o.pos1 : min(lenA1, lenB - o.pos1);
        if (o.len1 <= 0) return false;
        o.type1 = a_first ? 'p' : 's'; o.type2 = a_first ? 's' : 'p';
        o.perc = perc_one(o.len1, lenA1);
        o.pos2 = i2r - i1r;
        o.len2 = min(lenA2, lenB - o.pos2);
        if (o.len2 <= 0 || o.pos2 < 0) return false;
        o.perc2 = perc_one(o.len2, lenA2);
        return true;
    }
    if (!pA && pB) {                                                        // S-P :288-331
        const int lenA = (int)A.len1, lenB1 = (int)B.len1, lenB2 = (int)B.len2;
        o.len1 = a_first ? min(lenB1, lenA - o.pos1) : lenB1 - o.pos1;
        if (o.len1 <= 0) return false;
        o.type1 = a_first ? 's' : 'p'; o.type2 = a_first ? 'p' : 's';
        o.perc = perc_one(o.len1, lenB1);
        o.pos2 = i1r - i2r;
        o.len2 = min(lenB2, lenA - o.pos2);
        if (o.len2 <= 0 || o.pos2 < 0) return false;
        o.perc2 = perc_one(o.len2, lenB2);
        return true;
    }
    const int lenA = (int)A.len1, lenB = (int)B.len1, lenC = (int)A.len2, lenD = (int)B.len2;   // P-P :332-401
    o.len1 = a_first ? min(lenA - o.pos1, lenB) : min(lenA, lenB - o.pos1);
    const bool back = i1r - i2r >= 0;
    o.pos2 = back ? i1r - i2r : i2r - i1r;
    o.len2 = back ? min(lenC - o.pos2, lenD) : min(lenC, lenD - o.pos2);
    if (o.len1 <= 0 || o.len2 <= 0) return false;
    o.perc = perc_of(o.len1, lenA, lenB);
    o.perc2 = perc_of(o.len2, lenC, lenD);
    o.ord = (a_first == back) ? '1' : '2';
    o.type1 = 'p'; o.type2 = 'p';
    return true;
}

__global__ void fno3_count(const u64* off, u64 n, uint32_t* cnt) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) {
        const u64 c = off[k + 1] - off[k];
        cnt[k] = (uint32_t)(c * (c - 1) / 2);
    }
}

// MODE 0: claim (atomicMin of the sequence number per pair of new reads); 1: flag survivors; 2: emit
template <int MODE>
__global__ void fno3_pass(const u64* off, u64 n, const uint32_t* sr_idx, const hc_fno3_pos* sr_pos, const hc_fno_read* reads,
                          uint32_t no_inclusions, const u64* seq0, u64* keys, u64* mins, u64 mask, uint32_t* flags,
                          const u64* outpos, hc_fno_overlap* out, u64 out_cap) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) {
        u64 seq = seq0[k];
        for (u64 i = off[k]; i < off[k + 1]; i++) {
            for (u64 j = i + 1; j < off[k + 1]; j++, seq++) {
                const hc_fno_read A = reads[sr_idx[i]], B = reads[sr_idx[j]];
                const u64 key = (min(A.id, B.id) << 32) | max(A.id, B.id);
                u64 h = hash_slot(key, mask);
                if (MODE == 0) {
                    while (true) {
                        const u64 prev = atomicCAS(&keys[HC_FNO_CELL * h], ~0ull, key);
                        if (prev == ~0ull || prev == key) break;
                        h = (h + 1) & mask;
                    }
                    atomicMin(&mins[HC_FNO_CELL * h], seq);
                    continue;
                }
                if (MODE != 0) while (keys[HC_FNO_CELL * h] != key) h = (h + 1) & mask;
                bool ok = mins[HC_FNO_CELL * h] == seq;                                   // first original wins, :116-121
                hc_fno_overlap o;
                if (ok) ok = deduce_overlap(A, B, sr_pos[i], sr_pos[j], o);
                if (ok) {
                    const unsigned perc = o.perc2 > 0 ? (unsigned)(0.5 * (o.perc + o.perc2)) : (unsigned)o.perc;   // Overlap::get_perc
                    ok = !(no_inclusions && perc == 100) && o.len1 > 0;      // :157-165
                }
                // MODE 1: flag + record at the attempt's index; fno_compact moves the records to their ranks
                flags[seq] = ok;
                if (ok) out[seq] = o;
            }
        }
    }
}

thread_local std::string g_fno_err;

}  // namespace

extern "C" const char* hc_last_error(void);
void hc_set_last_error(const char* msg);   // hc_api.cu

#define FCU(call)                                                                         \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            hc_set_last_error((std::string(#call) + ": " + cudaGetErrorString(_e)).c_str()); \
            rc = HC_ERR_CUDA;                                                             \
            goto done;                                                                    \
        }                                                                                 \
    } while (0)

// HC_FNO_TIMING=1: phase times of hc_fno1 on stderr (synchronises the device at every mark)
struct FnoPhase {
    bool on;
    double t0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    FnoPhase() : on(getenv("HC_FNO_TIMING") != nullptr), t0(0) { if (on) t0 = now(); }
    void mark(const char* what) {
        if (!on) return;
        cudaDeviceSynchronize();
        const double t = now();
        fprintf(stderr, "[hc_fno1] %-28s %8.2f ms\n", what, t - t0);
        t0 = t;
    }
};

extern "C" int hc_fno1(const hc_fno_input* in, const hc_fno_edge* edges, uint64_t n_edges, hc_fno_overlap* out, uint64_t out_cap,
                       uint64_t* n_out, int device) {
    if (!in || !n_out || (n_edges && !edges) || (out_cap && !out)) { hc_set_last_error("hc_fno1: NULL argument"); return HC_ERR_ARG; }
    *n_out = 0;
    const u64 V = in->n_vertices, NS = in->n_superreads;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long i = 0; i < (long long)n_edges; i++) bad |= (edges[i].u >= V || edges[i].v >= V);
    if (bad) { hc_set_last_error("hc_fno1: edge vertex out of range"); return HC_ERR_ARG; }
    const u64 nsr = V ? in->sr_off[V] : 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long i = 0; i < (long long)nsr; i++) bad |= (in->sr_idx[i] >= NS);
    if (bad) { hc_set_last_error("hc_fno1: super-read index out of range"); return HC_ERR_ARG; }
    // the first-found-wins table keys a pair of new reads as (min id << 32 | max id): every id that can appear in a key
    // -- a super-read's, an unmerged vertex' -- must fit 32 bits (rename_fas.py numbers reads from 0), else distinct
    // pairs would collide and overlaps the reference finds would be dropped silently
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long i = 0; i < (long long)NS; i++) bad |= (in->superread[i].id >> 32) != 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long v = 0; v < (long long)V; v++) bad |= (!in->visited[v] && in->vertex_read[v].id != ~0ull && (in->vertex_read[v].id >> 32) != 0);
    if (bad) { hc_set_last_error("hc_fno1: a new read id does not fit 32 bits (ids of new reads must be below 2^32)"); return HC_ERR_ARG; }
    int rc = HC_OK;
    FnoPhase ph;
    ph.mark("argument checks");
    uint8_t *d_vis = nullptr, *d_lab = nullptr;
    hc_fno_read *d_vr = nullptr, *d_sr = nullptr;
    u64 *d_sroff = nullptr, *d_off = nullptr, *d_total = nullptr, *d_keys = nullptr, *d_mins = nullptr, *d_outpos = nullptr, *d_bsum = nullptr;
    uint32_t *d_sridx = nullptr, *d_cnt = nullptr, *d_flags = nullptr;
    hc_fno_subread* d_sub = nullptr;
    hc_fno_edge* d_edges = nullptr;
    hc_fno_overlap *d_out = nullptr, *d_rec = nullptr;
    u64 attempts = 0, produced = 0, cap = 64, ncopy;
    FnoDev D;
    const int threads = 256;
    int blocks;
    FCU(cudaSetDevice(device));
    if (n_edges == 0) goto done;
    blocks = (int)((n_edges + threads - 1) / threads < 4096 ? (n_edges + threads - 1) / threads : 4096);
    FCU(hc_scratch_alloc((void**)&d_vis, V ? V : 1)); FCU(hc_scratch_alloc((void**)&d_lab, V ? V : 1));
    FCU(hc_scratch_alloc((void**)&d_vr, (V ? V : 1) * sizeof(hc_fno_read))); FCU(hc_scratch_alloc((void**)&d_sr, (NS ? NS : 1) * sizeof(hc_fno_read)));
    FCU(hc_scratch_alloc((void**)&d_sroff, (V + 1) * sizeof(u64))); FCU(hc_scratch_alloc((void**)&d_sridx, (nsr ? nsr : 1) * sizeof(uint32_t)));
    FCU(hc_scratch_alloc((void**)&d_sub, (nsr ? nsr : 1) * sizeof(hc_fno_subread)));
    FCU(hc_scratch_alloc((void**)&d_edges, n_edges * sizeof(hc_fno_edge))); FCU(hc_scratch_alloc((void**)&d_cnt, n_edges * sizeof(uint32_t)));
    FCU(hc_scratch_alloc((void**)&d_off, n_edges * sizeof(u64))); FCU(hc_scratch_alloc((void**)&d_total, sizeof(u64)));
    FCU(hc_copy_h2d(d_vis, in->visited, V)); FCU(hc_copy_h2d(d_lab, in->label, V));
    FCU(hc_copy_h2d(d_vr, in->vertex_read, V * sizeof(hc_fno_read)));
    FCU(hc_copy_h2d(d_sr, in->superread, NS * sizeof(hc_fno_read)));
    FCU(hc_copy_h2d(d_sroff, in->sr_off, (V + 1) * sizeof(u64)));
    FCU(hc_copy_h2d(d_sridx, in->sr_idx, nsr * sizeof(uint32_t)));
    FCU(hc_copy_h2d(d_sub, in->sr_sub, nsr * sizeof(hc_fno_subread)));
    FCU(hc_copy_h2d(d_edges, edges, n_edges * sizeof(hc_fno_edge)));
    ph.mark("allocations + copies in");
    D.n_vertices = V; D.visited = d_vis; D.label = d_lab; D.vertex_read = d_vr; D.sr_off = d_sroff; D.sr_idx = d_sridx;
    D.sr_sub = d_sub; D.superread = d_sr; D.resolve_orientations = in->resolve_orientations; D.no_inclusions = in->no_inclusions;
    fno_count<<<blocks, threads>>>(D, d_edges, n_edges, d_cnt);
    FCU(hc_scratch_alloc((void**)&d_bsum, hc_scan::blocks_for(n_edges) * sizeof(u64)));
    hc_scan::exclusive_u32(d_cnt, n_edges, d_off, d_total, d_bsum, 0);
    FCU(cudaMemcpy(&attempts, d_total, sizeof(u64), cudaMemcpyDeviceToHost));
    ph.mark("count + scan");
    if (attempts == 0) goto done;
    while (cap < 2 * attempts + 2) cap <<= 1;
    // open-addressing table: pair key -> smallest sequence number (both arrays indexed [HC_FNO_CELL * slot])
    FCU(hc_scratch_alloc((void**)&d_keys, 2 * cap * sizeof(u64)));
    d_mins = d_keys + (HC_FNO_CELL == 2 ? 1 : cap);
    FCU(cudaMemset(d_keys, 0xff, 2 * cap * sizeof(u64)));
    FCU(hc_scratch_alloc((void**)&d_flags, attempts * sizeof(uint32_t))); FCU(hc_scratch_alloc((void**)&d_outpos, attempts * sizeof(u64)));
    FCU(cudaMemset(d_flags, 0, attempts * sizeof(uint32_t)));
    ph.mark("tables: alloc + memset");
    fno_claim<<<blocks, threads>>>(D, d_edges, n_edges, d_off, d_keys, d_mins, cap - 1);
    ph.mark("fno_claim");
    FCU(hc_scratch_alloc((void**)&d_rec, attempts * sizeof(hc_fno_overlap)));
    fno_resolve<<<blocks, threads>>>(D, d_edges, n_edges, d_off, d_keys, d_mins, cap - 1, d_flags, d_rec);
    ph.mark("fno_resolve");
    hc_scratch_free(d_bsum); d_bsum = nullptr;
    FCU(hc_scratch_alloc((void**)&d_bsum, hc_scan::blocks_for(attempts) * sizeof(u64)));
    hc_scan::exclusive_u32(d_flags, attempts, d_outpos, d_total, d_bsum, 0);
    FCU(cudaMemcpy(&produced, d_total, sizeof(u64), cudaMemcpyDeviceToHost));
    *n_out = produced;
    if (produced > out_cap) { hc_set_last_error("hc_fno1: output buffer too small (required size returned in n_out)"); rc = HC_ERR_CAPACITY; goto done; }
    if (produced == 0) goto done;
    ncopy = produced;
    FCU(hc_scratch_alloc((void**)&d_out, ncopy * sizeof(hc_fno_overlap)));
    ph.mark("scan of flags");
    fno_compact<<<148 * 16, 256>>>(d_flags, d_outpos, d_rec, attempts, d_out, ncopy);
    FCU(cudaGetLastError());
    ph.mark("fno_compact");
    FCU(hc_copy_d2h(out, d_out, ncopy * sizeof(hc_fno_overlap)));
    ph.mark("copy out");
done:
    hc_scratch_free(d_vis); hc_scratch_free(d_lab); hc_scratch_free(d_vr); hc_scratch_free(d_sr); hc_scratch_free(d_sroff); hc_scratch_free(d_sridx); hc_scratch_free(d_sub);
    hc_scratch_free(d_edges); hc_scratch_free(d_cnt); hc_scratch_free(d_off); hc_scratch_free(d_total); hc_scratch_free(d_keys);
    hc_scratch_free(d_flags); hc_scratch_free(d_outpos); hc_scratch_free(d_out); hc_scratch_free(d_bsum); hc_scratch_free(d_rec);
    return rc;
}


extern "C" int hc_fno3(uint64_t n_originals, const uint64_t* off, const uint32_t* sr_idx, const hc_fno3_pos* sr_pos, uint64_t n_reads,
                       const hc_fno_read* reads, int no_inclusions, hc_fno_overlap* out, uint64_t out_cap, uint64_t* n_out,
                       int device) {
    if (!n_out || (n_originals && (!off || !sr_idx || !sr_pos || !reads)) || (out_cap && !out)) {
        hc_set_last_error("hc_fno3: NULL argument");
        return HC_ERR_ARG;
    }
    *n_out = 0;
    if (n_originals == 0) return HC_OK;
    const u64 nent = off[n_originals];
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long i = 0; i < (long long)nent; i++) bad |= (sr_idx[i] >= n_reads);
    if (bad) { hc_set_last_error("hc_fno3: read index out of range"); return HC_ERR_ARG; }
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long i = 0; i < (long long)n_reads; i++) bad |= (reads[i].id >> 32) != 0;      // pair keys are (min id << 32 | max id)
    if (bad) { hc_set_last_error("hc_fno3: a new read id does not fit 32 bits (ids of new reads must be below 2^32)"); return HC_ERR_ARG; }
    int rc = HC_OK;
    u64 *d_off = nullptr, *d_seq = nullptr, *d_total = nullptr, *d_keys = nullptr, *d_mins = nullptr, *d_outpos = nullptr, *d_bsum = nullptr;
    uint32_t *d_idx = nullptr, *d_cnt = nullptr, *d_flags = nullptr;
    hc_fno3_pos* d_pos = nullptr;
    hc_fno_read* d_reads = nullptr;
    hc_fno_overlap *d_out = nullptr, *d_rec = nullptr;
    u64 attempts = 0, produced = 0, cap = 64;
    const int threads = 128;
    const int blocks = (int)((n_originals + threads - 1) / threads < 8192 ? (n_originals + threads - 1) / threads : 8192);
    FCU(cudaSetDevice(device));
    FCU(hc_scratch_alloc((void**)&d_off, (n_originals + 1) * sizeof(u64))); FCU(hc_scratch_alloc((void**)&d_idx, (nent ? nent : 1) * sizeof(uint32_t)));
    FCU(hc_scratch_alloc((void**)&d_pos, (nent ? nent : 1) * sizeof(hc_fno3_pos))); FCU(hc_scratch_alloc((void**)&d_reads, (n_reads ? n_reads : 1) * sizeof(hc_fno_read)));
    FCU(hc_scratch_alloc((void**)&d_cnt, n_originals * sizeof(uint32_t))); FCU(hc_scratch_alloc((void**)&d_seq, n_originals * sizeof(u64)));
    FCU(hc_scratch_alloc((void**)&d_total, sizeof(u64)));
    FCU(hc_copy_h2d(d_off, off, (n_originals + 1) * sizeof(u64)));
    FCU(hc_copy_h2d(d_idx, sr_idx, nent * sizeof(uint32_t)));
    FCU(hc_copy_h2d(d_pos, sr_pos, nent * sizeof(hc_fno3_pos)));
    FCU(hc_copy_h2d(d_reads, reads, n_reads * sizeof(hc_fno_read)));
    fno3_count<<<blocks, threads>>>(d_off, n_originals, d_cnt);
    FCU(hc_scratch_alloc((void**)&d_bsum, hc_scan::blocks_for(n_originals) * sizeof(u64)));
    hc_scan::exclusive_u32(d_cnt, n_originals, d_seq, d_total, d_bsum, 0);
    FCU(cudaMemcpy(&attempts, d_total, sizeof(u64), cudaMemcpyDeviceToHost));
    if (attempts == 0) goto done;
    while (cap < 2 * attempts + 2) cap <<= 1;
    // open-addressing table: pair key -> smallest sequence number (both arrays indexed [HC_FNO_CELL * slot])
    FCU(hc_scratch_alloc((void**)&d_keys, 2 * cap * sizeof(u64)));
    d_mins = d_keys + (HC_FNO_CELL == 2 ? 1 : cap);
    FCU(cudaMemset(d_keys, 0xff, 2 * cap * sizeof(u64)));
    FCU(hc_scratch_alloc((void**)&d_flags, attempts * sizeof(uint32_t))); FCU(hc_scratch_alloc((void**)&d_outpos, attempts * sizeof(u64)));
    fno3_pass<0><<<blocks, threads>>>(d_off, n_originals, d_idx, d_pos, d_reads, no_inclusions, d_seq, d_keys, d_mins, cap - 1, nullptr, nullptr, nullptr, 0);
    FCU(hc_scratch_alloc((void**)&d_rec, attempts * sizeof(hc_fno_overlap)));
    fno3_pass<1><<<blocks, threads>>>(d_off, n_originals, d_idx, d_pos, d_reads, no_inclusions, d_seq, d_keys, d_mins, cap - 1, d_flags, nullptr, d_rec, attempts);
    hc_scratch_free(d_bsum); d_bsum = nullptr;
    FCU(hc_scratch_alloc((void**)&d_bsum, hc_scan::blocks_for(attempts) * sizeof(u64)));
    hc_scan::exclusive_u32(d_flags, attempts, d_outpos, d_total, d_bsum, 0);
    FCU(cudaMemcpy(&produced, d_total, sizeof(u64), cudaMemcpyDeviceToHost));
    *n_out = produced;
    if (produced > out_cap) { hc_set_last_error("hc_fno3: output buffer too small (required size returned in n_out)"); rc = HC_ERR_CAPACITY; goto done; }
    if (produced == 0) goto done;
    FCU(hc_scratch_alloc((void**)&d_out, produced * sizeof(hc_fno_overlap)));
    fno_compact<<<148 * 16, 256>>>(d_flags, d_outpos, d_rec, attempts, d_out, produced);
    FCU(cudaGetLastError());
    FCU(hc_copy_d2h(out, d_out, produced * sizeof(hc_fno_overlap)));
done:
    hc_scratch_free(d_off); hc_scratch_free(d_idx); hc_scratch_free(d_pos); hc_scratch_free(d_reads); hc_scratch_free(d_cnt); hc_scratch_free(d_seq); hc_scratch_free(d_total);
    hc_scratch_free(d_keys); hc_scratch_free(d_flags); hc_scratch_free(d_outpos); hc_scratch_free(d_out); hc_scratch_free(d_bsum);
    hc_scratch_free(d_rec);
    return rc;
}
