// hc_adjacency.cu -- the adjacency lists of the overlap graph on the device: OverlapGraph::addEdge in insertion order
// (src/OverlapGraph.cpp:94-101) and OverlapGraph::sortEdges (:722-764: every list ordered by non-overlap length, then
// vertex2; adj_in rebuilt by walking the sorted lists).  SURVEY 8f rank 1, second half (the first half -- which edges
// survive the insert -- is hc_dedup_edges).
//
// One primitive, used twice: "group items by bucket and order every group by a 64-bit key, ties by item index":
//   count per bucket (atomicAdd) -> exclusive scan -> scatter with a per-bucket cursor (arbitrary order inside a group)
//   -> one thread per bucket sorts its group in place (insertion sort up to 32 items, heap sort beyond; the order is total,
//   so any sorting algorithm gives the same list).
// adj_out: bucket = vertex1, key = item index (insertion order) or non-overlap length << 32 | vertex2 (sortEdges);
// adj_in:  bucket = vertex2, key = position of the edge in the adj_out order.
#include <cstring>
#include <string>
#include <cuda_runtime.h>
#include "../../include/hc_b200.h"
#include "hc_scan.cuh"
#include "hc_stage.h"

namespace {

typedef unsigned long long u64;

struct Item { u64 key; uint32_t idx; uint32_t pad; };     // 16 bytes

__device__ __forceinline__ bool item_less(const Item& a, const Item& b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; }

__global__ void adj_count(const hc_adj_edge* __restrict__ e, const uint8_t* __restrict__ keep, u64 n, int by_v2, uint32_t* cnt) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        if (!keep || keep[i]) atomicAdd(&cnt[by_v2 ? e[i].vertex2 : e[i].vertex1], 1u);
}

// adj_out: the kept edges, bucket = vertex1
__global__ void adj_scatter_out(const hc_adj_edge* __restrict__ e, const uint8_t* __restrict__ keep, u64 n, int sort,
                                const u64* __restrict__ off, uint32_t* cursor, Item* items) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (keep && !keep[i]) continue;
        const hc_adj_edge x = e[i];
        const u64 p = off[x.vertex1] + atomicAdd(&cursor[x.vertex1], 1u);
        Item it;
        it.key = sort ? ((u64)x.nonoverlap_len << 32) | x.vertex2 : 0ull;
        it.idx = (uint32_t)i;
        it.pad = 0;
        items[p] = it;
    }
}

// adj_in: the edges in adj_out order (perm), bucket = vertex2, key = position in that order
__global__ void adj_scatter_in(const hc_adj_edge* __restrict__ e, const uint32_t* __restrict__ perm, u64 kept,
                               const u64* __restrict__ off, uint32_t* cursor, Item* items) {
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < kept; p += (u64)gridDim.x * blockDim.x) {
        const hc_adj_edge x = e[perm[p]];
        const u64 q = off[x.vertex2] + atomicAdd(&cursor[x.vertex2], 1u);
        Item it;
        it.key = p;
        it.idx = x.vertex1;          // what the in-list holds: the source vertex
        it.pad = 0;
        items[q] = it;
    }
}

__device__ void sift_down(Item* a, uint32_t start, uint32_t end) {
    uint32_t root = start;
    while (2 * root + 1 <= end) {
        uint32_t child = 2 * root + 1, sw = root;
        if (item_less(a[sw], a[child])) sw = child;
        if (child + 1 <= end && item_less(a[sw], a[child + 1])) sw = child + 1;
        if (sw == root) return;
        const Item t = a[root]; a[root] = a[sw]; a[sw] = t;
        root = sw;
    }
}

// one thread per bucket; out[p] = idx of the p-th item; tie[v] = 1 if two items of a group of more than 16 share a key
// (std::sort leaves their order to its implementation: the host part settles those lists with std::sort itself)
__global__ void adj_sort_groups(u64 V, const u64* __restrict__ off, Item* items, uint32_t* __restrict__ out, uint8_t* __restrict__ tie) {
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (u64)gridDim.x * blockDim.x) {
        const u64 a = off[v], b = off[v + 1];
        const uint32_t m = (uint32_t)(b - a);
        Item* g = items + a;
        if (m > 1 && m <= 32) {
            for (uint32_t i = 1; i < m; i++) {
                const Item x = g[i];
                uint32_t j = i;
                while (j > 0 && item_less(x, g[j - 1])) { g[j] = g[j - 1]; j--; }
                g[j] = x;
            }
        } else if (m > 32) {
            for (int s = (int)(m - 2) / 2; s >= 0; s--) sift_down(g, (uint32_t)s, m - 1);
            for (uint32_t end = m - 1; end > 0; end--) {
                const Item t = g[end]; g[end] = g[0]; g[0] = t;
                sift_down(g, 0, end - 1);
            }
        }
        bool t = false;
        for (uint32_t i = 0; i < m; i++) {
            out[a + i] = g[i].idx;
            if (i && g[i].key == g[i - 1].key) t = true;
        }
        if (tie) tie[v] = (t && m > 16) ? 1 : 0;
    }
}

}  // namespace

void hc_set_last_error(const char* msg);   // hc_api.cu

#define ACU(call)                                                                            \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            hc_set_last_error((std::string(#call) + ": " + cudaGetErrorString(_e)).c_str()); \
            rc = HC_ERR_CUDA;                                                                \
            goto done;                                                                       \
        }                                                                                    \
    } while (0)

extern "C" int hc_build_adjacency(const hc_adj_edge* edges, uint64_t n, const uint8_t* keep, uint64_t n_vertices, int sort,
                                  uint64_t* out_off, uint32_t* out_perm, uint64_t* in_off, uint32_t* in_src, uint8_t* ties,
                                  uint64_t* n_kept, int device) {
    if (!out_off || !n_kept || (n && (!edges || !out_perm)) || (in_off && n && !in_src)) { hc_set_last_error("hc_build_adjacency: NULL argument"); return HC_ERR_ARG; }
    if (n >> 32) { hc_set_last_error("hc_build_adjacency: more than 2^32 edges"); return HC_ERR_ARG; }
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long i = 0; i < (long long)n; i++) bad |= (edges[i].vertex1 >= n_vertices || edges[i].vertex2 >= n_vertices);
    if (bad) { hc_set_last_error("hc_build_adjacency: vertex out of range"); return HC_ERR_ARG; }
    *n_kept = 0;
    const u64 V = n_vertices;
    int rc = HC_OK;
    hc_adj_edge* d_e = nullptr;
    uint8_t *d_keep = nullptr, *d_tie = nullptr;
    uint32_t *d_cnt = nullptr, *d_perm = nullptr, *d_src = nullptr;
    u64 *d_off = nullptr, *d_total = nullptr, *d_bsum = nullptr, kept = 0;
    Item* d_items = nullptr;
    const int threads = 256;
    const int eb = (int)((n + threads - 1) / threads < 148 * 16 ? (n + threads - 1) / threads : 148 * 16);
    const int vb = (int)((V + threads - 1) / threads < 148 * 16 ? (V + threads - 1) / threads : 148 * 16);
    ACU(cudaSetDevice(device));
    if (V == 0) goto done;
    ACU(hc_scratch_alloc((void**)&d_cnt, (V + 1) * sizeof(uint32_t))); ACU(hc_scratch_alloc((void**)&d_off, (V + 1) * sizeof(u64)));
    ACU(hc_scratch_alloc((void**)&d_total, sizeof(u64))); ACU(hc_scratch_alloc((void**)&d_bsum, hc_scan::blocks_for(V + 1) * sizeof(u64)));
    ACU(cudaMemsetAsync(d_cnt, 0, (V + 1) * sizeof(uint32_t)));
    if (n) {
        ACU(hc_scratch_alloc((void**)&d_e, n * sizeof(hc_adj_edge)));
        ACU(hc_copy_h2d(d_e, edges, n * sizeof(hc_adj_edge)));
        if (keep) { ACU(hc_scratch_alloc((void**)&d_keep, n)); ACU(hc_copy_h2d(d_keep, keep, n)); }
        adj_count<<<eb ? eb : 1, threads>>>(d_e, d_keep, n, 0, d_cnt);
    }
    hc_scan::exclusive_u32(d_cnt, V + 1, d_off, d_total, d_bsum, 0);
    ACU(cudaMemcpy(&kept, d_total, sizeof(u64), cudaMemcpyDeviceToHost));
    *n_kept = kept;
    ACU(hc_copy_d2h(out_off, d_off, (V + 1) * sizeof(u64)));
    if (kept) {
        ACU(hc_scratch_alloc((void**)&d_items, kept * sizeof(Item))); ACU(hc_scratch_alloc((void**)&d_perm, kept * sizeof(uint32_t)));
        if (ties) ACU(hc_scratch_alloc((void**)&d_tie, V));
        ACU(cudaMemsetAsync(d_cnt, 0, (V + 1) * sizeof(uint32_t)));
        adj_scatter_out<<<eb, threads>>>(d_e, d_keep, n, sort, d_off, d_cnt, d_items);
        adj_sort_groups<<<vb, threads>>>(V, d_off, d_items, d_perm, sort ? d_tie : nullptr);
        ACU(cudaGetLastError());
        ACU(hc_copy_d2h(out_perm, d_perm, kept * sizeof(uint32_t)));
        if (ties) { if (sort) ACU(hc_copy_d2h(ties, d_tie, V)); else memset(ties, 0, V); }
    } else if (ties) {
        memset(ties, 0, V);
    }
    if (in_off) {
        ACU(cudaMemsetAsync(d_cnt, 0, (V + 1) * sizeof(uint32_t)));
        if (n) adj_count<<<eb ? eb : 1, threads>>>(d_e, d_keep, n, 1, d_cnt);
        hc_scan::exclusive_u32(d_cnt, V + 1, d_off, d_total, d_bsum, 0);
        ACU(hc_copy_d2h(in_off, d_off, (V + 1) * sizeof(u64)));
        if (kept) {
            ACU(hc_scratch_alloc((void**)&d_src, kept * sizeof(uint32_t)));
            ACU(cudaMemsetAsync(d_cnt, 0, (V + 1) * sizeof(uint32_t)));
            adj_scatter_in<<<(int)((kept + threads - 1) / threads < 148 * 16 ? (kept + threads - 1) / threads : 148 * 16), threads>>>(d_e, d_perm, kept, d_off, d_cnt, d_items);
            adj_sort_groups<<<vb, threads>>>(V, d_off, d_items, d_src, nullptr);
            ACU(cudaGetLastError());
            ACU(hc_copy_d2h(in_src, d_src, kept * sizeof(uint32_t)));
        }
    }
done:
    hc_scratch_free(d_e); hc_scratch_free(d_keep); hc_scratch_free(d_tie); hc_scratch_free(d_cnt); hc_scratch_free(d_perm); hc_scratch_free(d_src);
    hc_scratch_free(d_off); hc_scratch_free(d_total); hc_scratch_free(d_bsum); hc_scratch_free(d_items);
    return rc;
}
