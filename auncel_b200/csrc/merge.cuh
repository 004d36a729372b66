#pragma once
#include "scan.cuh"

namespace auncel {

struct TuneParams {
    int mode;              // 0 fixed, 1 tune, 2 training
    int query_topk;
    int profile, overhead_profile;
    int nprobe;
    long max_codes;
    int time_tune;                     // latency-budget cut (IndexIVF.cpp:545-549) on the modelled clock
    long long us_per_list, ns_per_code;
    ErrModelView model;
    const float* require_acc;          // n (device)
    const float* gt_kth;               // n or null
    float* t_recalls;                  // n or null
    const float* dtb;                  // n x max_num
    int max_num;
    float* snapshots;                  // training: n x n_traces x K
    int n_traces;
};

void launch_init_state(const RoundParams& rp, const TuneParams& tp, const unsigned long long* mynp_in,
                       int* active_out, cudaStream_t s);
void launch_set_online(int metric, long nlist, long n, const float* cdis, const int* ckeys,
                       const float* interdis, const float* arcos, int arcos_size, float* dtb,
                       int max_num, int* ctl, cudaStream_t s);
// merges the S * nsub partial results of every (query, rank) pair into the pair's first sub-slot
void launch_stage_merge(const RoundParams& rp, int num_sms, cudaStream_t s);
// orders the unsorted slots of a tensor-core round by (distance, offset), keeps the K best of each
void launch_slot_sort(const RoundParams& rp, int num_sms, cudaStream_t s);
void launch_merge_check(const RoundParams& rp, const TuneParams& tp, cudaStream_t s);
void launch_compact_active(const RoundParams& rp, int r1, int* active_out, int* h_ctl_pinned,
                           cudaStream_t s);
void launch_finalize(const RoundParams& rp, const TuneParams& tp, float* D, long long* I,
                     unsigned long long* mynp_out, unsigned long long* stats_dev, cudaStream_t s);

}  // namespace auncel
