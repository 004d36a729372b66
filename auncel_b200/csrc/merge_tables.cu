// merge_tables (/root/reference/Auncel/IndexShards.cpp:44-105) on the device: per query, a
// k-way merge over the heads of nshard sorted result rows.
//
// One warp per query.  The warp stages the query's nshard x k candidate distances in shared
// memory with coalesced loads and finds where every shard's row ends (first label < 0, :72,:96);
// lane 0 then replays the reference's size-nshard heap literally -- heap_push / heap_pop of
// Heap.h:88-142 on (value, shard) with the strict comparison of the *merge* heap (CMin for L2,
// CMax for IP, IndexShards.cpp:303-309) -- so that equal distances coming from different shards
// leave in the reference's order, not merely in some sorted order.  The walk records
// (shard, position) per output slot; all lanes then fetch the labels (+ translations, :93).
// Exhausted output slots get label -1 and C::neutral() of the merge heap: -FLT_MAX for L2,
// +FLT_MAX for IP (:88-90), exactly as the reference does.
#include "engine.h"

namespace auncel {

constexpr int MAX_SHARDS = 64;

template <int METRIC>
__device__ __forceinline__ bool mt_cmp(float a, float b) {  // C::cmp of the merge heap
    return METRIC == METRIC_L2 ? a < b : a > b;
}

template <int METRIC>
__global__ void merge_tables_kernel(long n, int k, int nshard, const float* __restrict__ all_D,
                                    const long long* __restrict__ all_I, long stride_D, long stride_I,
                                    const long long* __restrict__ tr, float* __restrict__ D,
                                    long long* __restrict__ I) {
    extern __shared__ __align__(8) unsigned char mt_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const size_t per_warp = (size_t)nshard * k + k + MAX_SHARDS;  // 4-byte words
    float* sD = reinterpret_cast<float*>(mt_smem) + warp * per_warp;
    unsigned* sOut = reinterpret_cast<unsigned*>(sD + (size_t)nshard * k);  // (shard << 24) | position, or ~0u
    int* sLen = reinterpret_cast<int*>(sOut + k);
    for (long q = blockIdx.x * (long)wpb + warp; q < n; q += (long)gridDim.x * wpb) {
        for (int s = 0; s < nshard; s++) {
            int first_neg = k;
            for (int p = lane; p < k; p += 32) {
                sD[s * k + p] = all_D[stride_D * s + q * k + p];
                if (all_I[stride_I * s + q * k + p] < 0) first_neg = min(first_neg, p);
            }
            for (int o = 16; o > 0; o >>= 1) first_neg = min(first_neg, __shfl_xor_sync(0xffffffffu, first_neg, o));
            if (lane == 0) sLen[s] = first_neg;
        }
        __syncwarp();
        if (lane == 0) {
            float hv[MAX_SHARDS + 1];   // 1-based heap of (value, shard)
            int hs[MAX_SHARDS + 1];
            int ptr[MAX_SHARDS];
            int heap_size = 0;
            auto push = [&](float val, int sh) {  // Heap.h:124-142
                int i = ++heap_size;
                while (i > 1) {
                    const int f = i >> 1;
                    if (!mt_cmp<METRIC>(val, hv[f])) break;
                    hv[i] = hv[f];
                    hs[i] = hs[f];
                    i = f;
                }
                hv[i] = val;
                hs[i] = sh;
            };
            auto pop = [&]() {  // Heap.h:88-117
                const int kk = heap_size--;
                const float val = hv[kk];
                int i = 1;
                while (true) {
                    const int i1 = i << 1, i2 = i1 + 1;
                    if (i1 > kk) break;
                    if (i2 == kk + 1 || mt_cmp<METRIC>(hv[i1], hv[i2])) {
                        if (mt_cmp<METRIC>(val, hv[i1])) break;
                        hv[i] = hv[i1];
                        hs[i] = hs[i1];
                        i = i1;
                    } else {
                        if (mt_cmp<METRIC>(val, hv[i2])) break;
                        hv[i] = hv[i2];
                        hs[i] = hs[i2];
                        i = i2;
                    }
                }
                hv[i] = hv[kk];
                hs[i] = hs[kk];
            };
            for (int s = 0; s < nshard; s++) {
                ptr[s] = 0;
                if (sLen[s] > 0) push(sD[s * k], s);
            }
            for (int j = 0; j < k; j++) {
                if (heap_size == 0) {
                    sOut[j] = ~0u;
                } else {
                    const int s = hs[1];
                    const int p = ptr[s]++;
                    sOut[j] = ((unsigned)s << 24) | (unsigned)p;
                    pop();
                    if (p + 1 < sLen[s]) push(sD[s * k + p + 1], s);
                }
            }
        }
        __syncwarp();
        for (int j = lane; j < k; j += 32) {
            const unsigned o = sOut[j];
            if (o == ~0u) {
                I[q * k + j] = -1;
                D[q * k + j] = METRIC == METRIC_L2 ? -FLT_MAX : FLT_MAX;
            } else {
                const int s = (int)(o >> 24), p = (int)(o & 0xffffffu);
                D[q * k + j] = sD[s * k + p];
                I[q * k + j] = all_I[stride_I * s + q * k + p] + (tr ? tr[s] : 0);
            }
        }
        __syncwarp();
    }
}

void launch_merge_tables(int metric, long n, long k, long nshard, const float* all_D, const long long* all_I,
                         const long long* translations, float* D, long long* I, cudaStream_t s) {
    launch_merge_tables_strided(metric, n, k, nshard, all_D, all_I, n * k, n * k, translations, D, I, s);
    if (s == nullptr) CUDA_CHECK(cudaStreamSynchronize(s));
}

// shard s's tables start at all_D + s * stride_D / all_I + s * stride_I (elements)
void launch_merge_tables_strided(int metric, long n, long k, long nshard, const float* all_D, const long long* all_I,
                                 long stride_D, long stride_I, const long long* translations, float* D, long long* I,
                                 cudaStream_t s) {
    AUNCEL_CHECK(nshard >= 1 && nshard <= MAX_SHARDS, "nshard must be in [1, 64]");
    AUNCEL_CHECK(k < (1 << 24), "k too large");
    if (n == 0 || k == 0) return;
    const size_t per_warp = ((size_t)nshard * k + k + MAX_SHARDS) * 4;
    AUNCEL_CHECK(per_warp <= 200 * 1024, "nshard * k too large for the on-chip merge");
    int wpb = (int)std::max<size_t>(1, std::min<size_t>(4, (96 * 1024) / per_warp));
    const size_t smem = per_warp * wpb;
    auto kern = metric == METRIC_L2 ? merge_tables_kernel<METRIC_L2> : merge_tables_kernel<METRIC_IP>;
    if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)std::min<long>((n + wpb - 1) / wpb, 148L * 16);
    kern<<<blocks, wpb * 32, smem, s>>>(n, (int)k, (int)nshard, all_D, all_I, stride_D, stride_I, translations, D, I);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
