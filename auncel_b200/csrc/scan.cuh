// Shared declarations of the round pipeline (plan -> scan -> merge/check).
#pragma once
#include "engine.h"

namespace auncel {

// control block slots (device ints, mirrored to pinned host memory once per round)
enum { CTL_N_ACTIVE = 0, CTL_TOTAL_TILES = 1, CTL_TILE_COUNTER = 2, CTL_TOTAL_PAIRS = 3,
       CTL_ERR = 4, CTL_NFIX = 5, CTL_NCAND = 6, CTL_OVERFLOW = 7, CTL_MIN_RCNT = 8, CTL_NOT_FULL = 9, CTL_MERGE_NEXT = 10,
       CTL_REDO_N = 11,
       CTL_REM_SUM = 12,  // sum over the still-active queries of min(stages left up to their bound, 4096)
       CTL_SIZE = 16 };

// per-query running state, SoA, carved from IvfIndex::state
struct QState {
    int* limit;      // hard stage limit (nprobe / max_codes)
    int* cut;        // 1: limit comes from max_codes (break before the tune block)
    int* bound;      // current upper bound on stages to scan (== stop stage once decided)
    int* decided;    // tune mode: my_nprobe fixed
    int* rcnt;       // valid entries in R
    float* tau;      // scan filter threshold (K-th best so far, or +-FLT_MAX)
    int* stoped;     // plateau counter (IndexIVF.cpp:588-598)
    float* pre_val;
    unsigned long long* mynp;
    float* Rd;                  // n x K   best-first distances
    unsigned long long* Rcode;  // n x K   (probe rank << 32) | offset in list
    static size_t bytes(long n, int K) {
        return (size_t)n * (7 * 4 + 8 + 8 /*align slack*/) + (size_t)n * K * 12 + 256;
    }
    void carve(unsigned char* base, long n, int K) {
        unsigned char* p = base;
        auto take = [&](size_t b) { unsigned char* r = p; p += (b + 15) / 16 * 16; return r; };
        Rcode = (unsigned long long*)take((size_t)n * K * 8);
        mynp = (unsigned long long*)take((size_t)n * 8);
        Rd = (float*)take((size_t)n * K * 4);
        limit = (int*)take((size_t)n * 4);
        cut = (int*)take((size_t)n * 4);
        bound = (int*)take((size_t)n * 4);
        decided = (int*)take((size_t)n * 4);
        rcnt = (int*)take((size_t)n * 4);
        tau = (float*)take((size_t)n * 4);
        stoped = (int*)take((size_t)n * 4);
        pre_val = (float*)take((size_t)n * 4);
    }
};

// slot_cnt[slot] = number of candidates | SLOT_SORTED when they are ordered by (distance, offset);
// unordered slots (tensor-core rerank, short exact-scan results) are ordered by merge_check
constexpr int SLOT_SORTED = 1 << 30;
// sharded rounds (shard_rounds.cu): a candidate's 32-bit code is (shard << SHARD_CODE_SHIFT) | offset in the shard's list
constexpr int SHARD_CODE_SHIFT = 26;

struct RoundParams {
    // index
    const float* codes;
    const long long* list_off;
    const long long* ids;
    int dpad;
    long nlist;
    int metric;
    // queries
    const float* xq;      // n x dpad
    const int* ckeys;     // n x nlist ranked centroid ids
    const float* cdis;    // n x nlist ranked centroid distances
    long n;
    int K;
    int cap;              // entries per candidate slot (>= K; the first tensor-core round uses wider slots)
    // round
    const int* active;    // n_active -> query
    int n_active;
    int r0, w, S;
    int qt;               // queries per scan tile this round: 32 / nsub
    int nsub;             // sub-slots per (query, rank, segment) = row subsets of the scan tile (1, 2 or 4)
    int nc;               // consumer warps of the scan CTA: 8, or 4 (two CTAs per SM; narrow tiles only)
    int defer_sort;       // 1: the exact scan may hand over <= K candidates unsorted (many queries: merge_check has the warps)
    int merged;           // 1: stage_merge_kernel reduced every pair's S * nsub sub-slots to its first one
    int unsorted;         // (logging)  1: slots were filled by rerank_kernel in arrival order (tensor-core rounds)
    int* pair_flag;       // tensor-core rounds: per slot, 1 = overflowed -> redo this pair with the exact scan
    int filtered;         // plan only the flagged pairs
    // exact redo of the flagged pairs of a tensor-core round: they are few and scattered, so they are
    // scanned with the narrow (row-split) tiles into a compact pool of their own -- flagged pair ->
    // redo_ord[pair] -> 4 sub-slots of K entries.  The scan launch of the redo has cand_d/cand_off/slot_cnt
    // pointing at that pool; merge_check reads it through the redo_* members.
    int* redo_ord;
    float* redo_d;
    unsigned* redo_off;
    int* redo_cnt;
    // plan
    int* list_cnt;
    int* list_pair_off;   // nlist + 1
    int* list_tile_off;   // nlist + 1
    int* list_cursor;
    unsigned long long* pairs;
    float* xq_sorted;     // total_pairs x dpad: query rows in pair order
    int* ctl;
    unsigned long long* round_work;  // [0] sum of |list| over the round's pairs (algorithmic distance evaluations),
                                     // [1] vectors of the distinct lists touched, [2] vectors staged (one pass per query tile)
    // pool: slot = ((a * w + p_rel) * S + seg) * nsub + sub, `cap` entries each
    float* cand_d;
    unsigned* cand_off;
    int* slot_cnt;
    QState st;
    int shard_rank;       // >= 0: candidates carry their shard in the code; labels of other shards' vectors are -1 here
};

void launch_plan(const RoundParams& rp, cudaStream_t s);  // rp.filtered: only flagged pairs, slot counts kept
void launch_scan(const RoundParams& rp, const void* codes_map, const void* queries_map, int num_sms, cudaStream_t s);
void make_queries_tensor_map(void* out_map, const float* xq_sorted, long long nrows, int dpad);
// CUtensorMap (128 B, 64 B aligned) over the list arena [nrows x dpad] f32, box 128 rows x 32 floats
void make_codes_tensor_map(void* out_map, const float* codes, long long nrows, int dpad, int box_rows = 128);

}  // namespace auncel
