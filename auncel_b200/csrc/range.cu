// IndexIVF::range_search (/root/reference/Auncel/IndexIVF.cpp:741-860) with
// IVFFlatScanner::scan_codes_range (IndexIVFFlat.cpp:139-155): every vector of the nprobe nearest
// lists whose distance passes C::cmp(radius, dis) -- L2: dis < radius, inner product: dis > radius.
//
// The result container of the reference (RangeSearchResult, AuxIndexStructures.h:31-50) is filled in
// two steps -- per-query counts -> lims -> do_allocation -> copy -- and so is this: pass 1 counts the
// hits of every (query, probe rank) pair, an exclusive scan turns the counts into offsets (lims =
// the offsets of rank 0), pass 2 recomputes the distances and writes the hits at their final place.
// Within a query the hits therefore appear in the reference's scan order (probe rank, then in-list
// order).  Distances use the reference's exact arithmetic (exact.cuh); one warp per pair, one list
// row per lane.
#include "exact.cuh"
#include "merge.cuh"

namespace auncel {

namespace {

template <int METRIC>
__device__ __forceinline__ float row_distance(const float* __restrict__ xq, const float* __restrict__ y, int dpad) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* a = reinterpret_cast<const float4*>(xq);
    const float4* b = reinterpret_cast<const float4*>(y);
    for (int k = 0; k < dpad / 4; k++) exact_step<METRIC>(s, a[k], b[k]);
    return exact_finish(s);
}

// WRITE == false: cnt[pair] = hits; WRITE == true: hits go to out_D/out_I at off[pair] + rank in list
template <int METRIC, bool WRITE>
__global__ void __launch_bounds__(128)
range_scan_kernel(const float* __restrict__ codes, const long long* __restrict__ list_off,
                  const long long* __restrict__ ids, int dpad, long nlist, const float* __restrict__ xq, long n,
                  int nprobe, const int* __restrict__ ckeys, float radius, unsigned long long* __restrict__ cnt_or_off,
                  float* __restrict__ out_D, long long* __restrict__ out_I, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) float s_q[];  // one query row per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long pair = blockIdx.x * (long)(blockDim.x >> 5) + warp;
    if (pair >= n * nprobe) return;
    const long q = pair / nprobe;
    const int p = (int)(pair - q * nprobe);
    float* sq = s_q + (size_t)warp * dpad;
    for (int c = lane; c < dpad; c += 32) sq[c] = xq[q * dpad + c];
    __syncwarp();
    const int l = ckeys[q * nlist + p];
    const long long L0 = list_off[l];
    const int L = (int)(list_off[l + 1] - L0);
    unsigned long long pos = WRITE ? cnt_or_off[pair] : 0ull;
    for (int v0 = 0; v0 < L; v0 += 32) {
        const int v = v0 + lane;
        bool hit = false;
        float dis = 0.f;
        if (v < L) {
            dis = row_distance<METRIC>(sq, codes + (L0 + v) * (long long)dpad, dpad);
            hit = METRIC == METRIC_L2 ? radius > dis : radius < dis;  // C::cmp(radius, dis), IndexIVFFlat.cpp:150
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (WRITE && hit) {
            const unsigned long long o = pos + __popc(m & ((1u << lane) - 1));
            out_D[o] = dis;
            out_I[o] = ids[L0 + v];
        }
        pos += __popc(m);
    }
    if (!WRITE && lane == 0) {
        cnt_or_off[pair] = pos;
        if (L > 0) {  // IndexIVFStats: nlist, ndis (IndexIVF.cpp:799-800)
            atomicAdd(&stats[0], 1ull);
            atomicAdd(&stats[1], (unsigned long long)L);
        }
    }
}

// exclusive scan of m counts in place (single block, chunked with a carry); total -> out_total
__global__ void __launch_bounds__(1024) range_offsets_kernel(unsigned long long* __restrict__ a, long m,
                                                            unsigned long long* __restrict__ out_total) {
    __shared__ unsigned long long s[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long base = 0; base < m; base += 1024) {
        const long i = base + threadIdx.x;
        const unsigned long long c = i < m ? a[i] : 0ull;
        s[threadIdx.x] = c;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const unsigned long long v = threadIdx.x >= off ? s[threadIdx.x - off] : 0ull;
            __syncthreads();
            s[threadIdx.x] += v;
            __syncthreads();
        }
        if (i < m) a[i] = carry + s[threadIdx.x] - c;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_total = carry;
}

__global__ void range_lims_kernel(const unsigned long long* __restrict__ off, long n, int nprobe,
                                  const unsigned long long* __restrict__ total, long long* __restrict__ lims) {
    long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (q < n) lims[q] = (long long)off[q * nprobe];
    if (q == n) lims[n] = (long long)*total;
}

}  // namespace

void IvfIndex::range_search(long n, const float* x_dev, float radius, int nprobe_in, long long* lims_host) {
    CUDA_CHECK(cudaSetDevice(device));
    AUNCEL_CHECK(trained, "index is not trained");
    AUNCEL_CHECK(nprobe_in >= 1, "nprobe must be >= 1");
    const int nprobe = (int)std::min<long>(nprobe_in, nlist);
    range_total = 0;
    stats = SearchStats();
    if (n == 0) {
        lims_host[0] = 0;
        return;
    }
    const float* xs = x_dev;
    if (dpad != d) {
        q_x.ensure((size_t)n * dpad);
        launch_pad_rows(x_dev, n, d, q_x.p, dpad, stream);
        xs = q_x.p;
    }
    CUDA_CHECK(cudaEventRecord(ev0, stream));
    coarse_rank(n, xs);  // quantizer->search (IndexIVF.cpp:748)
    CUDA_CHECK(cudaMemsetAsync(ctl.p, 0, (CTL_SIZE + 8) * sizeof(int), stream));
    if (exact_ties)
        launch_fix_ties(metric, c_raw.p, nlist, nprobe, entry_table(nprobe), nullptr, (int)n, c_tie0.p, nprobe, nullptr,
                        fix_list.p, ctl.p + CTL_NFIX, c_dis.p, c_keys.p, stream);
    const long pairs_n = n * (long)nprobe;
    unsigned long long* off = range_off.ensure(pairs_n + 4);
    unsigned long long* st = off + pairs_n;  // [0] lists visited, [1] codes scanned, [2] total hits
    CUDA_CHECK(cudaMemsetAsync(st, 0, 4 * sizeof(unsigned long long), stream));
    const int wpb = 4;
    const unsigned blocks = (unsigned)((pairs_n + wpb - 1) / wpb);
    const size_t smem = (size_t)wpb * dpad * sizeof(float);
    if (metric == METRIC_L2)
        range_scan_kernel<METRIC_L2, false><<<blocks, wpb * 32, smem, stream>>>(
            codes.p, list_off.p, ids.p, dpad, nlist, xs, n, nprobe, c_keys.p, radius, off, nullptr, nullptr, st);
    else
        range_scan_kernel<METRIC_IP, false><<<blocks, wpb * 32, smem, stream>>>(
            codes.p, list_off.p, ids.p, dpad, nlist, xs, n, nprobe, c_keys.p, radius, off, nullptr, nullptr, st);
    range_offsets_kernel<<<1, 1024, 0, stream>>>(off, pairs_n, st + 2);
    long long* d_lims = reinterpret_cast<long long*>(range_lims.ensure(n + 1));
    range_lims_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, stream>>>(off, n, nprobe, st + 2, d_lims);
    unsigned long long h_st[3];
    CUDA_CHECK(cudaMemcpyAsync(h_st, st, sizeof(h_st), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaMemcpyAsync(lims_host, d_lims, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    // RangeSearchResult::do_allocation, then the second pass
    range_total = (long long)h_st[2];
    range_D.ensure(std::max<size_t>(range_total, 1));
    range_I.ensure(std::max<size_t>(range_total, 1));
    if (range_total > 0) {
        if (metric == METRIC_L2)
            range_scan_kernel<METRIC_L2, true><<<blocks, wpb * 32, smem, stream>>>(
                codes.p, list_off.p, ids.p, dpad, nlist, xs, n, nprobe, c_keys.p, radius, off, range_D.p, range_I.p, st);
        else
            range_scan_kernel<METRIC_IP, true><<<blocks, wpb * 32, smem, stream>>>(
                codes.p, list_off.p, ids.p, dpad, nlist, xs, n, nprobe, c_keys.p, radius, off, range_D.p, range_I.p, st);
    }
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventRecord(ev1, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    stats.nq = n;
    stats.nlist = h_st[0];
    stats.ndis = h_st[1];
    stats.search_ms = ms;
    stats.launches = 8;
}

}  // namespace auncel
