/* fno_oracle.c -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of SRBuilder::findNextOverlaps
 * (FNO1) on array inputs: src/FindNextOverlaps.cpp updateOverlap :25-327, findCliqueIndex :331-347,
 * computeOverlapData :351-565, first-found-wins bookkeeping :84-97/:162-175/:261-273.
 * Parity status: PINNED against the unmodified reference (oracle/_ref/ref_driver --merge-fno1 runs
 * the reference's own findNextOverlaps() and dumps its inputs; tests/test_fno.py compares the
 * resulting overlaps.txt byte for byte, and tests/golden/fno1_*.npz keep reference outputs).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/hc_b200.h"

static int imin(int a, int b) { return a < b ? a : b; }

/* src/FindNextOverlaps.cpp:351-565.  r1/r2: the two new reads (a super-read, or the unmerged original read). */
static int compute_overlap_data(const hc_fno_read* r1, const hc_fno_read* r2, int idx1l, int idx1r, int idx2l, int idx2r,
                                const hc_fno_edge* e, int* new_pos1, int* new_pos2, char* ord1, char* ord2, char* type1,
                                char* type2, int* perc, int* ol1, int* ol2) {
    const int pos1 = e->pos1, pos2 = e->pos2;
    const int p1 = r1->len2 > 0, p2 = r2->len2 > 0;
    const int l11 = (int)r1->len1, l12 = (int)r1->len2, l21 = (int)r2->len1, l22 = (int)r2->len2;
    int len;
    *new_pos2 = 0;
    if (!p1 && !p2) {                                              /* S-S :357-385 */
        *type1 = 's'; *type2 = 's';
        *new_pos1 = (pos1 + idx1l) - idx2l;
        if (*new_pos1 < 0) { *ord1 = '2'; *new_pos1 = -*new_pos1; len = l21; }
        else { *ord1 = '1'; len = l11; }
        *ol1 = imin(imin(len - *new_pos1, l11), l21);
        *ol2 = 0;
        { float a = *ol1 / (float)l11, b = *ol1 / (float)l21; float m = a > b ? a : b; *perc = (int)floor(m * 100); }
        *ord2 = '-';
        if (*new_pos1 >= len) return 0;
    } else if (p1 && !p2) {                                        /* P-S :387-443 */
        *type1 = 'p'; *type2 = 's';
        const int len1 = l11 + l12, len2 = l21;
        *new_pos1 = (pos1 + idx1l) - idx2l;
        if (*new_pos1 < 0) {
            *ord1 = '2'; *new_pos1 = -*new_pos1;
            if (*new_pos1 >= l21) return 0;
            *ol1 = l11;
        } else {
            *ord1 = '1';
            if (*new_pos1 >= l11) return 0;
            *ol1 = l11 - *new_pos1;
        }
        if (e->ord == '1') *new_pos2 = idx2r - (idx1r + pos2);
        else *new_pos2 = (pos2 + idx2r) - idx1r;
        if (*new_pos2 >= l21) return 0;
        else if (*new_pos2 < 0) return 0;
        *ord2 = '-';
        *ol2 = imin(l21 - *new_pos2, l12);
        { int t = *ol1 + *ol2; float a = t / (float)len1, b = t / (float)len2; float m = a > b ? a : b; *perc = (int)floor(m * 100); }
        *perc = imin(*perc, 100);
    } else if (!p1 && p2) {                                        /* S-P :445-489 */
        *type1 = 's'; *type2 = 'p';
        const int len1 = l11, len2 = l21 + l22;
        *new_pos1 = pos1 + idx1l - idx2l;
        if (*new_pos1 < 0) {
            *ord1 = '2'; *new_pos1 = -*new_pos1;
            if (*new_pos1 >= l21) return 0;
            *ol1 = l21 - *new_pos1;
        } else {
            *ord1 = '1';
            if (*new_pos1 >= l11) return 0;
            *ol1 = l21;
        }
        if (e->ord == '2') *new_pos2 = idx1r - (pos2 + idx2r);
        else *new_pos2 = idx1r + pos2 - idx2r;
        if (*new_pos2 >= l11) return 0;
        else if (*new_pos2 < 0) return 0;
        *ord2 = '-';
        *ol2 = imin(l11 - *new_pos2, l22);
        { int t = *ol1 + *ol2; float a = t / (float)len1, b = t / (float)len2; float m = a > b ? a : b; *perc = (int)floor(m * 100); }
        *perc = imin(*perc, 100);
    } else {                                                       /* P-P :491-551 */
        *type1 = 'p'; *type2 = 'p';
        *new_pos1 = (pos1 + idx1l) - idx2l;
        if (*new_pos1 < 0) {
            *ord1 = '2'; *new_pos1 = -*new_pos1;
            if (*new_pos1 >= l21) return 0;
            *ol1 = imin(l11, l21 - *new_pos1);
        } else {
            *ord1 = '1';
            if (*new_pos1 >= l11) return 0;
            *ol1 = imin(l11 - *new_pos1, l21);
        }
        if (e->ord == '1') *new_pos2 = (pos2 + idx1r) - idx2r;
        else *new_pos2 = idx1r - (pos2 + idx2r);
        if (*new_pos2 < 0) {
            *ord2 = (*ord1 == '1') ? '2' : '1';
            *new_pos2 = -*new_pos2;
            if (*new_pos2 >= l22) return 0;
            *ol2 = imin(l12, l22 - *new_pos2);
        } else {
            *ord2 = (*ord1 == '1') ? '1' : '2';
            if (*new_pos2 >= l12) return 0;
            *ol2 = imin(l12 - *new_pos2, l22);
        }
        { int t = *ol1 + *ol2; float a = t / (float)(l11 + l12), b = t / (float)(l21 + l22); float m = a > b ? a : b; *perc = (int)floor(m * 100); }
        *perc = imin(*perc, 100);
    }
    return 1;
}

/* findCliqueIndex :331-347 for the three call patterns of updateOverlap */
static void clique_indices(const hc_fno_subread* s, int sr_paired, int read_paired, int* idxl, int* idxr) {
    *idxl = s->index1 - s->startpos1;
    *idxr = (sr_paired || read_paired) ? s->index2 - s->startpos2 : *idxl;
}

/* open-addressing set of 64-bit keys standing in for overlaps_found (vector<set<read_id_t>>) */
typedef struct { uint64_t* k; uint64_t cap; } keyset;
static int keyset_test_and_set(keyset* s, uint64_t key) {   /* returns 1 if the key was already present */
    uint64_t h = (key * 0x9E3779B97F4A7C15ull) & (s->cap - 1);
    while (s->k[h] != ~0ull) {
        if (s->k[h] == key) return 1;
        h = (h + 1) & (s->cap - 1);
    }
    s->k[h] = key;
    return 0;
}

int hco_fno1(const hc_fno_input* in, const hc_fno_edge* edges, uint64_t n_edges, hc_fno_overlap* out, uint64_t out_cap,
             uint64_t* n_out) {
    uint64_t attempts = 0;
    for (uint64_t i = 0; i < n_edges; i++) {
        const hc_fno_edge* e = &edges[i];
        uint64_t a = in->visited[e->u] ? in->sr_off[e->u + 1] - in->sr_off[e->u] : 1;
        uint64_t b = in->visited[e->v] ? in->sr_off[e->v + 1] - in->sr_off[e->v] : 1;
        attempts += a * b;
    }
    keyset ks;
    ks.cap = 64;
    while (ks.cap < 2 * attempts + 2) ks.cap <<= 1;
    ks.k = (uint64_t*)malloc(ks.cap * sizeof(uint64_t));
    if (!ks.k) return HC_ERR_NOMEM;
    memset(ks.k, 0xff, ks.cap * sizeof(uint64_t));
    uint64_t n = 0;
    int rc = 0;
    for (uint64_t i = 0; i < n_edges; i++) {
        const hc_fno_edge* e = &edges[i];
        const uint32_t u = e->u, v = e->v;
        char ori1 = '+', ori2 = '+';
        if (in->resolve_orientations && e->nonedge) {                               /* :34-37 */
            ori1 = (e->ori1 == in->label[u]) ? '+' : '-';
            ori2 = (e->ori2 == in->label[v]) ? '+' : '-';
        }
        const hc_fno_read* ru = &in->vertex_read[u];
        const hc_fno_read* rv = &in->vertex_read[v];
        const int pu = ru->len2 > 0, pv = rv->len2 > 0;
        if (!in->visited[u] && !in->visited[v]) {                                   /* :46-72 */
            if (!(in->no_inclusions && e->perc == 100)) {
                if (n < out_cap) {
                    hc_fno_overlap* o = &out[n];
                    memset(o, 0, sizeof(*o));
                    o->id1 = ru->id; o->id2 = rv->id; o->pos1 = e->pos1; o->pos2 = e->pos2; o->ord = e->ord;
                    o->ori1 = ori1; o->ori2 = ori2; o->perc = e->perc; o->len1 = e->len1; o->len2 = e->len2;
                    o->type1 = pu ? 'p' : 's'; o->type2 = pv ? 'p' : 's';
                }
                n++;
            }
            continue;
        }
        const uint64_t a0 = in->visited[u] ? in->sr_off[u] : 0, a1 = in->visited[u] ? in->sr_off[u + 1] : 1;
        const uint64_t b0 = in->visited[v] ? in->sr_off[v] : 0, b1 = in->visited[v] ? in->sr_off[v + 1] : 1;
        for (uint64_t a = a0; a < a1; a++) {
            const hc_fno_read* s1 = in->visited[u] ? &in->superread[in->sr_idx[a]] : ru;
            for (uint64_t b = b0; b < b1; b++) {
                const hc_fno_read* s2 = in->visited[v] ? &in->superread[in->sr_idx[b]] : rv;
                if (s1->id == s2->id) continue;                                     /* :241-243 (asserted != in the other cases) */
                const uint64_t lo = s1->id < s2->id ? s1->id : s2->id, hi = s1->id < s2->id ? s2->id : s1->id;
                if (keyset_test_and_set(&ks, (lo << 32) | hi)) continue;            /* first found wins, even if it fails below */
                int idx1l = 0, idx1r = 0, idx2l = 0, idx2r = 0;
                if (in->visited[u]) clique_indices(&in->sr_sub[a], s1->len2 > 0, pu, &idx1l, &idx1r);
                if (in->visited[v]) clique_indices(&in->sr_sub[b], s2->len2 > 0, pv, &idx2l, &idx2r);
                int np1, np2, perc, ol1, ol2;
                char ord1, ord2, t1, t2;
                if (!compute_overlap_data(s1, s2, idx1l, idx1r, idx2l, idx2r, e, &np1, &np2, &ord1, &ord2, &t1, &t2, &perc, &ol1, &ol2))
                    continue;
                if (in->no_inclusions && perc == 100) continue;
                if (n < out_cap) {
                    hc_fno_overlap* o = &out[n];
                    memset(o, 0, sizeof(*o));
                    if (ord1 == '1') { o->id1 = s1->id; o->id2 = s2->id; o->type1 = t1; o->type2 = t2; }
                    else { o->id1 = s2->id; o->id2 = s1->id; o->type1 = t2; o->type2 = t1; }
                    o->pos1 = np1; o->pos2 = np2; o->ord = ord2; o->ori1 = ori1; o->ori2 = ori2;
                    o->perc = perc; o->len1 = ol1; o->len2 = ol2;
                }
                n++;
            }
        }
    }
    free(ks.k);
    *n_out = n;
    if (n > out_cap) rc = HC_ERR_CAPACITY;
    return rc;
}

/* ---- FindNextOverlaps3 ------------------------------------------------------------------------------
 * src/FindNextOverlaps3.cpp: nodeDictApproach :90-173 (first original wins per pair of new reads,
 * :116-121; drop rules :157-165) and deduceOverlap :176-406.  Originals arrive in the iteration
 * order of the reference's unordered_map (dumped by ref_driver). */
static int perc_max(int l, int a, int b) {
    float x = l / (float)a, y = l / (float)b;
    float m = x > y ? x : y;
    return (int)floor(m * 100);
}

static int deduce_overlap(const hc_fno_read* A, const hc_fno_read* B, const hc_fno3_pos* pa, const hc_fno3_pos* pb,
                          hc_fno_overlap* o) {   /* returns 0 for the reference's "this overlap will be ignored" */
    const int pA = A->len2 > 0, pB = B->len2 > 0;
    memset(o, 0, sizeof(*o));
    o->ori1 = '+'; o->ori2 = '+'; o->ord = '-';
    if (!pA && !pB) {                                                       /* S-S :202-243 */
        const int idx1 = pa->index1, idx2 = pb->index1, lenA = (int)A->len1, lenB = (int)B->len1;
        if (idx1 - idx2 >= 0) {
            o->id1 = A->id; o->id2 = B->id; o->pos1 = idx1 - idx2;
            if (o->pos1 > lenA) return 0;
            o->len1 = imin(lenA - o->pos1, lenB);
        } else {
            o->id1 = B->id; o->id2 = A->id; o->pos1 = idx2 - idx1;
            if (o->pos1 > lenB) return 0;
            o->len1 = imin(lenA, lenB - o->pos1);
        }
        o->perc = perc_max(o->len1, lenA, lenB);
        o->type1 = 's'; o->type2 = 's';
        return 1;
    }
    const int i1l = pa->index1, i1r = pa->index2, i2l = pb->index1, i2r = pb->index2;
    if (pA && !pB) {                                                        /* P-S :244-287 */
        const int lenA1 = (int)A->len1, lenA2 = (int)A->len2, lenB = (int)B->len1;
        if (i1l - i2l >= 0) {
            o->id1 = A->id; o->id2 = B->id; o->pos1 = i1l - i2l; o->len1 = lenA1 - o->pos1;
            if (o->len1 <= 0) return 0;
            o->type1 = 'p'; o->type2 = 's';
        } else {
            o->id1 = B->id; o->id2 = A->id; o->pos1 = i2l - i1l; o->len1 = imin(lenA1, lenB - o->pos1);
            if (o->len1 <= 0) return 0;
            o->type1 = 's'; o->type2 = 'p';
        }
        o->perc = (int)floor(o->len1 / (float)lenA1 * 100);
        o->pos2 = i2r - i1r;
        o->len2 = imin(lenA2, lenB - o->pos2);
        if (o->len2 <= 0 || o->pos2 < 0) return 0;
        o->perc2 = (int)floor(o->len2 / (float)lenA2 * 100);
        return 1;
    }
    if (!pA && pB) {                                                        /* S-P :288-331 */
        const int lenA = (int)A->len1, lenB1 = (int)B->len1, lenB2 = (int)B->len2;
        if (i1l - i2l >= 0) {
            o->id1 = A->id; o->id2 = B->id; o->pos1 = i1l - i2l; o->len1 = imin(lenB1, lenA - o->pos1);
            if (o->len1 <= 0) return 0;
            o->type1 = 's'; o->type2 = 'p';
        } else {
            o->id1 = B->id; o->id2 = A->id; o->pos1 = i2l - i1l; o->len1 = lenB1 - o->pos1;
            if (o->len1 <= 0) return 0;
            o->type1 = 'p'; o->type2 = 's';
        }
        o->perc = (int)floor(o->len1 / (float)lenB1 * 100);
        o->pos2 = i1r - i2r;
        o->len2 = imin(lenB2, lenA - o->pos2);
        if (o->len2 <= 0 || o->pos2 < 0) return 0;
        o->perc2 = (int)floor(o->len2 / (float)lenB2 * 100);
        return 1;
    }
    {                                                                       /* P-P :332-401 */
        const int lenA = (int)A->len1, lenB = (int)B->len1, lenC = (int)A->len2, lenD = (int)B->len2;
        int front, back;
        if (i1l - i2l >= 0) { o->id1 = A->id; o->id2 = B->id; o->pos1 = i1l - i2l; o->len1 = imin(lenA - o->pos1, lenB); front = 1; }
        else { o->id1 = B->id; o->id2 = A->id; o->pos1 = i2l - i1l; o->len1 = imin(lenA, lenB - o->pos1); front = 0; }
        if (i1r - i2r >= 0) { o->pos2 = i1r - i2r; o->len2 = imin(lenC - o->pos2, lenD); back = 1; }
        else { o->pos2 = i2r - i1r; o->len2 = imin(lenC, lenD - o->pos2); back = 0; }
        if (o->len1 <= 0 || o->len2 <= 0) return 0;
        o->perc = perc_max(o->len1, lenA, lenB);
        o->perc2 = perc_max(o->len2, lenC, lenD);
        o->ord = (front == back) ? '1' : '2';
        o->type1 = 'p'; o->type2 = 'p';
        return 1;
    }
}

int hco_fno3(uint64_t n_originals, const uint64_t* off, const uint32_t* sr_idx, const hc_fno3_pos* sr_pos, uint64_t n_reads,
             const hc_fno_read* reads, int no_inclusions, hc_fno_overlap* out, uint64_t out_cap, uint64_t* n_out) {
    uint64_t attempts = 0;
    for (uint64_t k = 0; k < n_originals; k++) { uint64_t c = off[k + 1] - off[k]; attempts += c * (c - 1) / 2; }
    keyset ks;
    ks.cap = 64;
    while (ks.cap < 2 * attempts + 2) ks.cap <<= 1;
    ks.k = (uint64_t*)malloc(ks.cap * sizeof(uint64_t));
    if (!ks.k) return HC_ERR_NOMEM;
    memset(ks.k, 0xff, ks.cap * sizeof(uint64_t));
    uint64_t n = 0;
    (void)n_reads;
    for (uint64_t k = 0; k < n_originals; k++) {
        for (uint64_t i = off[k]; i < off[k + 1]; i++) {
            for (uint64_t j = i + 1; j < off[k + 1]; j++) {
                const hc_fno_read* A = &reads[sr_idx[i]];
                const hc_fno_read* B = &reads[sr_idx[j]];
                const uint64_t lo = A->id < B->id ? A->id : B->id, hi = A->id < B->id ? B->id : A->id;
                if (keyset_test_and_set(&ks, (lo << 32) | hi)) continue;                      /* :116-121 */
                hc_fno_overlap o;
                if (!deduce_overlap(A, B, &sr_pos[i], &sr_pos[j], &o)) continue;             /* len1 == 0 -> not written, :160 */
                const unsigned perc = o.perc2 > 0 ? (unsigned)(0.5 * (o.perc + o.perc2)) : (unsigned)o.perc;   /* Overlap::get_perc */
                if (no_inclusions && perc == 100) continue;                                  /* :157-159 */
                if (!(o.len1 > 0)) continue;
                if (n < out_cap) out[n] = o;
                n++;
            }
        }
    }
    free(ks.k);
    *n_out = n;
    return n > out_cap ? HC_ERR_CAPACITY : 0;
}
