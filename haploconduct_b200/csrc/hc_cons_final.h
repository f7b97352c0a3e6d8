// hc_cons_final.h -- see hc_cons_final.cpp
#ifndef HC_CONS_FINAL_H_
#define HC_CONS_FINAL_H_
#include <cstdint>
#include "../../include/hc_b200.h"

void hc_cons_addends(double* out /* [94][2]: log10(1 - p_q), log10(p_q / 3) */);
// one problem: columns trim_pos.. from the device scores; characters go to cons_seq/cons_qual + out_offset
void hc_cons_walk(const hc_cons_problem* P, const hc_cons_seq* seqs, const uint32_t* seq_len, const double* sums,
                  const uint16_t* count, uint32_t min_clique_size, double min_qual, char* cons_seq, char* cons_qual,
                  hc_cons_result* res);
#endif
