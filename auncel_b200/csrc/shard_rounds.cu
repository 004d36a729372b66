// Error-bounded search over shards with SINGLE-INDEX semantics (the "rounds" form).
//
// The reference's own distributed mode (IndexShards.cpp:261-311 over copy_subset_to shards) lets every
// shard take its stop decision from its own partial top-k; its my_nprobe is therefore per shard.  Here
// the shards instead reproduce what ONE index holding all the vectors would do (IndexIVF.cpp:515-660):
// every rank keeps the full per-query state (top-max_topk, stage, decision) and advances it in the same
// rounds; a rank only scans its part of every probed list.  Between the scan and the stage-order replay
// (merge_check_kernel) of a round the ranks exchange what their scans produced:
//
//   local scan  ->  <= K best per (query, rank-stage) pair, sorted          (stage_merge / slot_sort)
//   collect     ->  compact entries (query, stage, distance, shard << 26 | offset): only non-empty pairs travel
//   ncclAllGather (counts, then entries padded to the largest count)
//   place       ->  one K-wide slot per pair holding the K best of the union
//   merge_check ->  the same replay as on one GPU, on every rank, on identical input
//
// A vector that enters the single index's top-k at a stage is among the K best of its list for that
// threshold in its own shard, so the union's K best per pair are the single index's; the state after
// every round is bit-identical on all ranks and equal to the one-GPU run's (ties between equal distances
// are ordered by (shard, local offset) instead of the list insertion order).  Because the state is
// identical, the host round loops of all ranks take the same decisions (window, filter or exact scan,
// number of rounds) and the collectives line up without any control exchange.
#include "engine.h"
#include "merge.cuh"

namespace auncel {

namespace {

// (query, stage offset in the round) rather than the pair index: the active list is compacted with atomics,
// so its ORDER differs from rank to rank -- only its content is the same
struct __align__(16) XEntry {
    unsigned q;
    unsigned p_rel;
    float d;
    unsigned code;
};
static_assert(sizeof(XEntry) == 16, "packed exchange entry");

__global__ void xc_inverse_kernel(const int* __restrict__ active, int n_active, int* __restrict__ inv) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_active) inv[active[a]] = a;
}

// where a pair's reduced candidates live after the local scan: main pool (first sub-slot) or redo pool
__device__ __forceinline__ int xc_source(const RoundParams& rp, long pidx, int nseg, const float*& cd, const unsigned*& co) {
    if (rp.redo_ord != nullptr && rp.pair_flag != nullptr && rp.pair_flag[pidx] != 0) {
        const long slot = (long)rp.redo_ord[pidx] * 4;
        cd = rp.redo_d + slot * rp.K;
        co = rp.redo_off + slot * rp.K;
        return min(rp.redo_cnt[slot] & ~SLOT_SORTED, rp.K);
    }
    const long slot = pidx * nseg;
    cd = rp.cand_d + slot * rp.cap;
    co = rp.cand_off + slot * rp.cap;
    const int c = rp.slot_cnt[slot] & ~SLOT_SORTED;
    return c > rp.cap ? 0 : min(c, rp.K);  // (an overflowed slot is always flagged, handled above)
}

// EMIT = false: count this rank's entries;  EMIT = true: write them (positions reserved per warp)
template <bool EMIT>
__global__ void xc_collect_kernel(RoundParams rp, long npairs, int nseg, unsigned shard, XEntry* out,
                                  unsigned long long* counter) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long base = warp * 32; base < npairs; base += nwarps * 32) {
        const long pidx = base + lane;
        const float* cd = nullptr;
        const unsigned* co = nullptr;
        const int c = pidx < npairs ? xc_source(rp, pidx, nseg, cd, co) : 0;
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long start = 0;
        if (lane == 0) start = atomicAdd(counter, (unsigned long long)total);
        if (!EMIT) continue;
        start = __shfl_sync(0xffffffffu, start, 0) + (unsigned long long)(incl - c);
        unsigned todo = __ballot_sync(0xffffffffu, c > 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int cs = __shfl_sync(0xffffffffu, c, src);
            const unsigned long long ss = __shfl_sync(0xffffffffu, start, src);
            const float* cds = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)cd, src));
            const unsigned* cos = reinterpret_cast<const unsigned*>(__shfl_sync(0xffffffffu, (unsigned long long)co, src));
            for (int i = lane; i < cs; i += 32) {
                XEntry e;
                e.q = (unsigned)rp.active[(base + src) / rp.w];
                e.p_rel = (unsigned)((base + src) % rp.w);
                e.d = cds[i];
                e.code = (shard << SHARD_CODE_SHIFT) | cos[i];
                out[ss + i] = e;
            }
        }
    }
}

// entries of all ranks (rank r: all[r * stride .. + counts[r])) -> per-pair totals
__global__ void xc_count_kernel(const XEntry* __restrict__ all, const unsigned long long* __restrict__ counts,
                                size_t stride, int world, const int* __restrict__ inv, int w, int* __restrict__ cnt2) {
    for (int r = 0; r < world; r++) {
        const size_t c = counts[r];
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < c; i += (size_t)gridDim.x * blockDim.x) {
            const XEntry e = all[r * stride + i];
            atomicAdd(&cnt2[(size_t)inv[e.q] * w + e.p_rel], 1);
        }
    }
}

// pairs whose union exceeds a slot get an ordinal in the overflow pool
__global__ void xc_assign_kernel(long npairs, int K, const int* __restrict__ cnt2, int* __restrict__ ovf_ord,
                                 int* __restrict__ n_ovf) {
    const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    ovf_ord[p] = cnt2[p] > K ? atomicAdd(n_ovf, 1) : -1;
}

__global__ void xc_place_kernel(const XEntry* __restrict__ all, const unsigned long long* __restrict__ counts,
                                size_t stride, int world, const int* __restrict__ inv, int w, int K, int wide,
                                const int* __restrict__ ovf_ord, int* __restrict__ fill, float* __restrict__ pd, unsigned* __restrict__ po,
                                float* __restrict__ od, unsigned* __restrict__ oo) {
    for (int r = 0; r < world; r++) {
        const size_t c = counts[r];
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < c; i += (size_t)gridDim.x * blockDim.x) {
            const XEntry e = all[r * stride + i];
            const size_t pair = (size_t)inv[e.q] * w + e.p_rel;
            const int pos = atomicAdd(&fill[pair], 1);
            const int ord = ovf_ord[pair];
            if (ord < 0) {
                pd[pair * K + pos] = e.d;
                po[pair * K + pos] = e.code;
            } else {
                od[(size_t)ord * wide + pos] = e.d;
                oo[(size_t)ord * wide + pos] = e.code;
            }
        }
    }
}

// overflow pairs: order the union (up to world * K entries) by (distance, code), keep the K best in the
// pair's regular slot.  One warp per pair, keys in shared memory (P = next power of two of `wide`).
__global__ void xc_reduce_kernel(long npairs, int K, int wide, int P, int metric, const int* __restrict__ ovf_ord,
                                 int* __restrict__ cnt2, const float* __restrict__ od, const unsigned* __restrict__ oo,
                                 float* __restrict__ pd, unsigned* __restrict__ po) {
    extern __shared__ unsigned long long xr_key[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned long long* key = xr_key + (size_t)warp * P;
    for (long base = ((long)blockIdx.x * wpb + warp) * 32; base < npairs; base += (long)gridDim.x * wpb * 32) {
        const long mine = base + lane;
        const int ord_l = mine < npairs ? ovf_ord[mine] : -1;
        unsigned todo = __ballot_sync(0xffffffffu, ord_l >= 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const long pidx = base + src;
            const int ord = __shfl_sync(0xffffffffu, ord_l, src);
            const int rc = min(cnt2[pidx], wide);
            for (int i = lane; i < P; i += 32) {
                unsigned long long kk = ~0ull;
                if (i < rc) {
                    uint32_t o = f2ord(od[(size_t)ord * wide + i]);
                    if (metric == METRIC_IP) o = ~o;
                    kk = ((unsigned long long)o << 32) | oo[(size_t)ord * wide + i];
                }
                key[i] = kk;
            }
            __syncwarp();
            for (int size = 2; size <= P; size <<= 1)
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int t = lane; t < P / 2; t += 32) {
                        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                        const bool up = ((lo & size) == 0);
                        const unsigned long long x = key[lo], y = key[hi];
                        if ((x > y) == up) {
                            key[lo] = y;
                            key[hi] = x;
                        }
                    }
                    __syncwarp();
                }
            const int keep = min(rc, K);
            for (int i = lane; i < keep; i += 32) {
                const unsigned long long kk = key[i];
                uint32_t o = (uint32_t)(kk >> 32);
                if (metric == METRIC_IP) o = ~o;
                pd[(size_t)pidx * K + i] = ord2f(o);
                po[(size_t)pidx * K + i] = (unsigned)(kk & 0xffffffffu);
            }
            __syncwarp();
            if (lane == 0) cnt2[pidx] = keep | SLOT_SORTED;
        }
    }
}

}  // namespace

// Called between the scan phase and merge_check of a round.  On return rp describes the merged pool:
// one slot of K entries per (query, rank-stage) pair, ordered, no redo indirection.
void IvfIndex::exchange_candidates(RoundParams& rp, size_t nredo) {
    ShardExchange& X = *shard_x;
    const long npairs = (long)rp.n_active * rp.w;
    const int K = rp.K, world = X.world;
    if (npairs == 0) return;
    AUNCEL_CHECK((unsigned long long)npairs < (1ull << 32), "round too large for the exchange (pairs)");
    // 1. one sorted sub-slot of <= K entries per pair
    const int nseg = rp.S * rp.nsub;
    if (nseg > 1 && !rp.merged) {
        launch_stage_merge(rp, num_sms, stream);
        rp.merged = 1;
    }
    if (rp.redo_ord != nullptr && nredo > 0) {
        RoundParams rv = rp;
        rv.n_active = (int)nredo;
        rv.w = 1;
        rv.S = 1;
        rv.nsub = 4;
        rv.cap = K;
        rv.cand_d = rp.redo_d;
        rv.cand_off = rp.redo_off;
        rv.slot_cnt = rp.redo_cnt;
        launch_stage_merge(rv, num_sms, stream);
    }
    // 2. how many entries does this rank send?  (count, all-gather the counts, read them)
    unsigned long long* ctr = X.ctr.ensure(2 + (size_t)world);  // [0] local count, [1] emit cursor, [2..] all counts
    CUDA_CHECK(cudaMemsetAsync(ctr, 0, (2 + (size_t)world) * sizeof(unsigned long long), stream));
    const unsigned cblocks = (unsigned)std::min<long>((npairs + 255) / 256, (long)num_sms * 16);
    xc_collect_kernel<false><<<cblocks, 256, 0, stream>>>(rp, npairs, nseg, (unsigned)X.rank, nullptr, ctr);
    CUDA_CHECK(cudaGetLastError());
    X.all_gather(ctr, ctr + 2, sizeof(unsigned long long), stream);
    X.h_counts.resize(world);
    CUDA_CHECK(cudaMemcpyAsync(X.h_counts.data(), ctr + 2, world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    size_t maxc = 0, total = 0;
    for (int r = 0; r < world; r++) {
        maxc = std::max<size_t>(maxc, X.h_counts[r]);
        total += X.h_counts[r];
    }
    X.entries_sent += X.h_counts[X.rank];
    X.entries_recv += total;
    X.exchanges++;
    // 3. entries: every rank sends `maxc` slots (the tail of a shorter list is never read)
    int* cnt2 = X.cnt2.ensure((size_t)npairs);
    CUDA_CHECK(cudaMemsetAsync(cnt2, 0, (size_t)npairs * sizeof(int), stream));
    X.pool.ensure((size_t)npairs * K * 8);
    float* pd = reinterpret_cast<float*>(X.pool.p);
    unsigned* po = reinterpret_cast<unsigned*>(X.pool.p + (size_t)npairs * K * 4);
    if (maxc > 0) {
        XEntry* mine = reinterpret_cast<XEntry*>(X.send.ensure(maxc * sizeof(XEntry)));
        XEntry* all = reinterpret_cast<XEntry*>(X.recv.ensure(maxc * sizeof(XEntry) * world));
        xc_collect_kernel<true><<<cblocks, 256, 0, stream>>>(rp, npairs, nseg, (unsigned)X.rank, mine, ctr + 1);
        CUDA_CHECK(cudaGetLastError());
        X.all_gather(mine, all, maxc * sizeof(XEntry), stream);
        X.bytes_recv += maxc * sizeof(XEntry) * world;
        // 4. place: per-pair totals, overflow ordinals, scatter, reduce
        const unsigned eblocks = (unsigned)std::min<size_t>((maxc + 255) / 256, (size_t)num_sms * 16);
        int* inv = X.inv.ensure((size_t)rp.n);
        xc_inverse_kernel<<<(unsigned)((rp.n_active + 255) / 256), 256, 0, stream>>>(rp.active, rp.n_active, inv);
        xc_count_kernel<<<eblocks, 256, 0, stream>>>(all, ctr + 2, maxc, world, inv, rp.w, cnt2);
        int* ovf_ord = X.ovf_ord.ensure((size_t)npairs);
        int* fill = X.fill.ensure((size_t)npairs + 1);
        CUDA_CHECK(cudaMemsetAsync(fill, 0, ((size_t)npairs + 1) * sizeof(int), stream));
        xc_assign_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, stream>>>(npairs, K, cnt2, ovf_ord, fill + npairs);
        // a pair overflows when its union has more than K entries: at most total / (K + 1) pairs
        const int wide = world * K;
        const size_t max_ovf = total / ((size_t)K + 1) + 1;
        X.ovf_pool.ensure(max_ovf * wide * 8);
        float* od = reinterpret_cast<float*>(X.ovf_pool.p);
        unsigned* oo = reinterpret_cast<unsigned*>(X.ovf_pool.p + max_ovf * wide * 4);
        xc_place_kernel<<<eblocks, 256, 0, stream>>>(all, ctr + 2, maxc, world, inv, rp.w, K, wide, ovf_ord, fill, pd, po, od, oo);
        int P = 64;
        while (P < wide) P <<= 1;
        const int wpb = P <= 1024 ? 4 : 1;
        const size_t smem = (size_t)wpb * P * sizeof(unsigned long long);
        AUNCEL_CHECK(smem <= 200 * 1024, "world * K too large for the exchange");
        if (smem > 48 * 1024)
            CUDA_CHECK(cudaFuncSetAttribute(xc_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned rblocks = (unsigned)std::min<long>((npairs + wpb * 32 - 1) / (wpb * 32), (long)num_sms * 8);
        xc_reduce_kernel<<<rblocks, wpb * 32, smem, stream>>>(npairs, K, wide, P, rp.metric, ovf_ord, cnt2, od, oo, pd, po);
        CUDA_CHECK(cudaGetLastError());
    }
    // 5. the round continues on the merged pool
    rp.cand_d = pd;
    rp.cand_off = po;
    rp.slot_cnt = cnt2;
    rp.cap = K;
    rp.S = 1;
    rp.nsub = 1;
    rp.merged = 0;
    rp.pair_flag = nullptr;
    rp.redo_ord = nullptr;
    rp.redo_d = nullptr;
    rp.redo_off = nullptr;
    rp.redo_cnt = nullptr;
    if (maxc > 0) launch_slot_sort(rp, num_sms, stream);  // pairs within a slot: (distance, code) order
}

}  // namespace auncel
