// merge_tables (/root/reference/Auncel/IndexShards.cpp:44-105) on the device: per query, a
// k-way merge over the heads of nshard sorted result rows.  The reference keeps a size-nshard
// heap; nshard is the number of GPUs (<= 16), so a linear scan over the heads selects the
// same element (lowest shard on equal distances).  Labels < 0 end a shard's row; exhausted
// output slots get label -1 and C::neutral() of the *merge* heap -- CMin for L2, i.e.
// -FLT_MAX (:58-60, :303-305), CMax for IP, i.e. +FLT_MAX -- exactly as the reference does.
#include "engine.h"

namespace auncel {

constexpr int MAX_SHARDS = 64;

__global__ void merge_tables_kernel(int metric, long n, long k, int nshard, const float* __restrict__ all_D,
                                    const long long* __restrict__ all_I, const long long* __restrict__ tr,
                                    float* __restrict__ D, long long* __restrict__ I) {
    long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const long stride = n * k;
    int ptr[MAX_SHARDS];
    for (int s = 0; s < nshard; s++) ptr[s] = 0;
    for (long j = 0; j < k; j++) {
        int best = -1;
        float bv = 0.f;
        for (int s = 0; s < nshard; s++) {
            int p = ptr[s];
            if (p >= k) continue;
            if (all_I[stride * s + q * k + p] < 0) continue;
            float v = all_D[stride * s + q * k + p];
            if (best < 0 || (metric == METRIC_L2 ? v < bv : v > bv)) {
                best = s;
                bv = v;
            }
        }
        if (best < 0) {
            I[q * k + j] = -1;
            D[q * k + j] = metric == METRIC_L2 ? -FLT_MAX : FLT_MAX;
        } else {
            D[q * k + j] = bv;
            I[q * k + j] = all_I[stride * best + q * k + ptr[best]] + (tr ? tr[best] : 0);
            ptr[best]++;
        }
    }
}

void launch_merge_tables(int metric, long n, long k, long nshard, const float* all_D, const long long* all_I,
                         const long long* translations, float* D, long long* I, cudaStream_t s) {
    AUNCEL_CHECK(nshard >= 1 && nshard <= MAX_SHARDS, "nshard must be in [1, 64]");
    if (n == 0 || k == 0) return;
    merge_tables_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(metric, n, k, (int)nshard, all_D, all_I,
                                                                  translations, D, I);
    CUDA_CHECK(cudaGetLastError());
    if (s == nullptr) CUDA_CHECK(cudaStreamSynchronize(s));
}

}  // namespace auncel
