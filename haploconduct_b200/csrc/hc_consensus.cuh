// hc_consensus.cuh -- launcher of the consensus score kernel (hc_consensus.cu), called from hc_api.cu.
#ifndef HC_CONSENSUS_CUH_
#define HC_CONSENSUS_CUH_
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/hc_b200.h"
#include "hc_layout.h"

struct hc_cons_dev {          // the store planes of one device
    const uint8_t* pk;
    const uint8_t* qual;
    const uint32_t* base2;
    const uint32_t* nmask;
    const hc_rdesc* rdesc;
    int packed;
};

// d_tile_off[p] = first 256-column tile of problem p (n_prob + 1 entries); per-column outputs are indexed by out_offset + column.
// d_base / d_qual: the consensus character of every column by the device's libm; d_marked: columns the host must redo.
cudaError_t hc_launch_cons_sums(const hc_cons_dev& D, const hc_cons_problem* d_prob, uint64_t n_prob, const hc_cons_seq* d_seqs,
                                const unsigned long long* d_tile_off, uint64_t n_tiles, const double* d_addend,
                                const int8_t* d_code_to_q, double min_qual, double* d_sums, uint16_t* d_count, char* d_base,
                                char* d_qual, unsigned long long* d_marked, uint64_t marked_cap, unsigned long long* d_n_marked,
                                cudaStream_t st);
cudaError_t hc_launch_cons_gather(const unsigned long long* d_marked, uint64_t n, const double* d_sums, double* d_out, cudaStream_t st);
#endif
