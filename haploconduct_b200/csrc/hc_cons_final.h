// hc_cons_final.h -- see hc_cons_final.cpp
#ifndef HC_CONS_FINAL_H_
#define HC_CONS_FINAL_H_
#include <cstdint>
#include "../../include/hc_b200.h"

void hc_cons_addends(double* out /* [94][2]: log10(1 - p_q), log10(p_q / 3) */);
// :349-401 for one column from its four scores with the host libm; 0 <=> the reference's consensus_pos returns 0
int hc_cons_final_pos(double sA, double sC, double sT, double sG, unsigned n_active, double min_qual, char* base, char* qual);
// one problem: cons_seq / cons_qual hold the per-column characters (quality 0 = consensus_pos failed) from
// out_offset on; the walk of :447-513 compacts them into the consensus strings in place
void hc_cons_walk(const hc_cons_problem* P, const hc_cons_seq* seqs, const uint32_t* seq_len, const uint16_t* count,
                  uint32_t min_clique_size, char* cons_seq, char* cons_qual, hc_cons_result* res);
#endif
