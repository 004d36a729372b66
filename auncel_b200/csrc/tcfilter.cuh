#pragma once
#include "scan.cuh"

namespace auncel {

struct TcArgs {
    const float* vnorm;   // ||v||^2 per arena row
    const float* list_nmax;  // max ||v||^2 per inverted list (the error margin of a tile uses it for every row)
    const float* qnorm;   // ||q||^2 per query of the batch
    float c1, c2, c3;     // error-bound constants, see tcfilter.cu
    unsigned long long* cand;  // survivors: (pair position << 32) | offset in list
    int cand_cap;
    int N;                // queries per tile (multiple of 32, <= 256)
    int stream_b;         // tcfilter.cu: 1 = the query tile is too large to stay resident (d > 256): its k-chunks
                          // travel through the stage ring next to the list's, 256 queries per tile
    int dry;              // experiments (tcfilter3.cu): 1 = accumulators released unread, 2 = read and tested, nothing appended
};

bool tc_stream_queries(int dpad);  // tcfilter.cu may stream the query tile for this dimension (tile = 256 queries)
int tc_tile_queries(int dpad, bool streamed = false);
void launch_row_norms(const float* x, long long n, int dpad, float* out, cudaStream_t s);
void launch_list_norm_max(const float* vnorm, const long long* list_off, long nlist, float* out, cudaStream_t s);
void launch_tc_filter(const RoundParams& rp, const TcArgs& ta, const void* codes_map, const void* queries_map,
                      int num_sms, cudaStream_t s);
// TMEM-resident queries (tcfilter2.cu): queries per tile for this dimension, 0 = not supported (d > 256);
// both tensor maps with 64-row boxes
int tc2_tile_queries(int dpad);
void launch_tc_filter2(const RoundParams& rp, const TcArgs& ta, const void* codes_map64, const void* queries_map64,
                       int num_sms, cudaStream_t s);
// CTA pairs (tcfilter3.cu, cta_group::2): each CTA keeps half of the tile's queries; queries map with an N/2-row box
int tc3_tile_queries(int dpad);
void launch_tc_filter3(const RoundParams& rp, const TcArgs& ta, const void* codes_map, const void* queries_map_half,
                       int num_sms, cudaStream_t s);
void launch_rerank(const RoundParams& rp, const TcArgs& ta, int num_sms, cudaStream_t s);
void launch_tc_audit(const RoundParams& tc, const float* ex_d, const unsigned* ex_off, const int* ex_cnt,
                     unsigned long long* ctr, cudaStream_t s);
// tensor map over xq_sorted with an N-row, 128B-swizzled box (the MMA's B operand)
void make_queries_tensor_map_tc(void* out_map, const float* xq_sorted, long long nrows, int dpad, int N);

}  // namespace auncel
