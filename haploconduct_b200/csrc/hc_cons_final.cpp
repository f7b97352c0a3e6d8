// hc_cons_final.cpp -- host side of hc_consensus: the addend table and, per consensus column, the part of
// SRBuilder::consensus_pos that goes through pow / log10 (src/SRBuilder.cpp:349-401), plus the column walk of
// SRBuilder::consensus (:406-522).  Compiled by the host compiler with -ffp-contract=off and linked against the
// same libm the reference uses, so that identical scores give identical characters.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "hc_cons_final.h"

void hc_cons_addends(double* out /* [94][2] */) {
    for (int q = 0; q <= 93; q++) {
        const double p = pow(10, -q / 10.0);          // SRBuilder::phred_to_prob, :289-293
        out[2 * q] = log10(1 - p);                     // the base that was read, :318
        out[2 * q + 1] = log10(p / 3.0);               // each of the three others, :319-321
    }
}

// :349-401 from the four scores; returns 0 when the reference returns 0 (p_incorrect is NaN)
int hc_cons_final_pos(double sA, double sC, double sT, double sG, unsigned n_active, double min_qual, char* base, char* qual) {
    const double max_score = std::max({sA, sT, sC, sG});
    const double max_prob = std::pow(10.0, max_score);
    const double total_prob = std::pow(10.0, sA) + std::pow(10.0, sT) + std::pow(10.0, sC) + std::pow(10.0, sG);
    if (max_score == 0 || total_prob == 0.0) { *base = 'N'; *qual = '$'; return 1; }
    const double p_incorrect = 1 - (max_prob / total_prob);
    if (n_active > 1 && (1 - p_incorrect) < min_qual) { *base = 'N'; *qual = '$'; return 1; }
    if (p_incorrect != p_incorrect) return 0;
    int phred;
    if (p_incorrect < std::pow(10.0, -9.3)) phred = 93;
    else phred = (int)round(-10 * log10(p_incorrect));
    if (phred < 0) phred = 0;
    else if (phred > 93) phred = 93;
    char nuc;
    if (max_score == sA) nuc = 'A';
    else if (max_score == sT) nuc = 'T';
    else if (max_score == sC) nuc = 'C';
    else nuc = 'G';
    *base = nuc;
    *qual = (char)(phred + 33);
    return 1;
}

void hc_cons_walk(const hc_cons_problem* P, const hc_cons_seq* seqs, const uint32_t* seq_len, const uint16_t* count,
                  uint32_t min_clique_size, char* cons_seq, char* cons_qual, hc_cons_result* res) {
    const uint64_t n = P->seq_end - P->seq_begin;
    const hc_cons_seq* e = seqs + P->seq_begin;
    const uint32_t* len = seq_len + P->seq_begin;
    const unsigned min_support = P->subreads_needed ? 2u : min_clique_size;      // :412-417
    int trim = 0;
    res->ret = 0;
    res->length = 0;
    if (P->error_correction) {                                                   // :420-433
        uint64_t k = 0;
        unsigned support = 1;
        while (support < min_support && k < n) { support++; k++; }
        if (k == n) { res->ret = -1; return; }                                   // "Not enough support for super-read."
        trim = e[k].pos;
    }
    // the sequences that started before trim_pos resume at trim_pos - pos (:438-445); one that is already over
    // at that offset makes the reference give up at the first column it processes (:468-472)
    if (trim < P->total_len)
        for (uint64_t j = 0; j < n && e[j].pos <= trim; j++)
            if ((uint32_t)(trim - e[j].pos) >= len[j]) return;                   // ret 0, empty strings
    const int last_start = n ? e[n - 1].pos : 0;
    char* cs = cons_seq + P->out_offset;      // per-column characters on entry, the consensus strings on return
    char* cq = cons_qual + P->out_offset;
    int out = 0;
    for (int c = trim; c < P->total_len; c++) {
        const unsigned n_active = count[P->out_offset + (uint64_t)c];
        if (P->error_correction && n_active < min_support && c >= last_start) break;   // suffix without support, :459-462
        if (n_active == 0) { res->ret = 0; res->length = 0; return; }                  // nobody covers the column, :488-491
        if (cq[c] == 0) {                                                               // consensus_pos returned 0 (:366-369)
            res->ret = trim;                                                            // :507-511: strings cleared, trim_pos returned
            res->length = 0;
            return;
        }
        cs[out] = cs[c];                     // out <= c: the strings start at the problem's offset
        cq[out] = cq[c];
        out++;
    }
    res->ret = trim;
    res->length = out;
}
