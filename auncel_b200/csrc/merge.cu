// Subsystems (3) + (4): per-query running top-k and Auncel's termination check, on device.
//
// merge_check_kernel replays, for every active query, the sequential part of
// IndexIVF::search_preassigned (/root/reference/Auncel/IndexIVF.cpp:526-675): after the
// candidates of probe rank `stage` are merged into the running sorted top-k, the tune block
// (:551-638) is evaluated exactly as the reference does -- sorted copy of the heap
// (IP: arccos LUT first), cur_num (IVF_pro.cpp:258-291), the plateau counter, my_nprobe =
// stage * multipler, the break rule -- or, in calibration mode, the top-k snapshot of the
// power-of-two stages (:640-673) is emitted.  The decision is a pure function of the sorted
// top-k after each stage, so merging per-(query, list) partial results in probe order
// reproduces the reference's heap walk (tests/test_lowlevel_ivf.cpp:426-557 invariant).
#include "merge.cuh"

namespace auncel {

__device__ __forceinline__ float neutral(int metric) { return metric == METRIC_L2 ? FLT_MAX : -FLT_MAX; }

// --------------------------------------------------------------------------- init
__global__ void init_state_kernel(RoundParams rp, TuneParams tp, const unsigned long long* mynp_in,
                                  int* active_out) {
    long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (q >= rp.n) return;
    int limit = tp.nprobe, cut = 0;
    if (tp.max_codes > 0 || tp.time_tune) {
        // Both cuts depend only on the sizes of the lists along the probe order, so the stage at which
        // the reference breaks is known before anything is scanned.
        long cum = 0;
        unsigned long long vt_ns = 0;
        const double t0 = 0.0;  // IndexIVF::time() at the start of the query, on the modelled clock
        const double budget = tp.time_tune ? dmul((double)tp.require_acc[q], 0.95) : 0.0;
        for (int p = 0; p < tp.nprobe; p++) {
            int l = rp.ckeys[q * rp.nlist + p];
            const long long sz = rp.list_off[l + 1] - rp.list_off[l];
            cum += sz;
            if (tp.max_codes > 0 && cum >= tp.max_codes) {  // IndexIVF.cpp:541-543
                limit = p + 1;
                cut = 1;
                break;
            }
            if (tp.time_tune) {  // IndexIVF.cpp:545-549; time() = tv_sec + tv_usec * 1e-6 (:329-333)
                vt_ns += (unsigned long long)tp.us_per_list * 1000ull + (unsigned long long)sz * (unsigned long long)tp.ns_per_code;
                const unsigned long long us = vt_ns / 1000ull;
                const double now = dadd((double)(us / 1000000ull), dmul((double)(us % 1000000ull), 1e-6));
                const double el = dmul(dsub(now, t0), 1000.0);
                if (el >= dsub(budget, __ddiv_rn(el, (double)(p + 1)))) {
                    limit = p + 1;
                    cut = 1;
                    break;
                }
            }
        }
    }
    int bound = limit, decided = 0;
    unsigned long long mynp = 0;
    const long n8 = rp.nlist / 8;
    if (tp.mode == 2) {
        bound = (int)min((long)limit, n8 + 1);  // scans one list past nlist/8, then breaks (:643)
    } else if (tp.mode == 1) {
        if (tp.overhead_profile) {
            bound = (int)min((long)limit, max(1L, n8));  // :633-636
            decided = 1;
        } else {
            mynp = mynp_in ? mynp_in[q] : 0ull;
            if (mynp != 0) {  // stale decision from an earlier call is replayed (:627-632)
                unsigned long long b = mynp < (unsigned long long)limit ? mynp : (unsigned long long)limit;
                bound = (int)b;
                decided = 1;
            }
        }
    }
    rp.st.limit[q] = limit;
    rp.st.cut[q] = cut;
    rp.st.bound[q] = bound;
    rp.st.decided[q] = decided;
    rp.st.rcnt[q] = 0;
    rp.st.tau[q] = neutral(rp.metric);
    rp.st.stoped[q] = 0;
    rp.st.pre_val[q] = 0.f;
    rp.st.mynp[q] = mynp;
    active_out[q] = (int)q;
}

void launch_init_state(const RoundParams& rp, const TuneParams& tp, const unsigned long long* mynp_in,
                       int* active_out, cudaStream_t s) {
    if (rp.n == 0) return;
    init_state_kernel<<<(unsigned)((rp.n + 127) / 128), 128, 0, s>>>(rp, tp, mynp_in, active_out);
    CUDA_CHECK(cudaGetLastError());
}

// --------------------------------------------------------------------------- set_online
// error_pro::set_online, IVF_pro.cpp:196-238: one warp per query.
__global__ void set_online_kernel(int metric, long nlist, long n, const float* __restrict__ cdis,
                                  const int* __restrict__ ckeys, const float* __restrict__ interdis,
                                  const float* __restrict__ arcos, int arcos_size, float* __restrict__ dtb,
                                  int max_num, int* ctl) {
    long q = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (q >= n) return;
    const float* cd = cdis + q * nlist;
    const int* ci = ckeys + q * nlist;
    int err = 0;
    const int cur = ci[0];
    float a = cd[0];
    if (metric == METRIC_IP) a = arcos_lookup(arcos, arcos_size, a, &err);
    for (int k = lane; k < max_num; k += 32) {
        float out = 0.f;  // dtb[max_num-1] is never written by the reference (stays 0)
        if (k < max_num - 1 && k + 1 < nlist) {
            float c2c = interdis[tri_index((size_t)nlist, (size_t)cur, (size_t)ci[k + 1])];
            float b = cd[k + 1];
            if (metric == METRIC_IP) b = arcos_lookup(arcos, arcos_size, b, &err);
            out = cosine_theorem(a, b, c2c, &err);
        }
        dtb[q * max_num + k] = out;
    }
    if (err) atomicOr(&ctl[CTL_ERR], err);
}

void launch_set_online(int metric, long nlist, long n, const float* cdis, const int* ckeys,
                       const float* interdis, const float* arcos, int arcos_size, float* dtb,
                       int max_num, int* ctl, cudaStream_t s) {
    if (n == 0) return;
    set_online_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, s>>>(metric, nlist, n, cdis, ckeys, interdis,
                                                                      arcos, arcos_size, dtb, max_num, ctl);
    CUDA_CHECK(cudaGetLastError());
}

// --------------------------------------------------------------------------- slot ordering
// rerank_kernel appends the survivors of a (query, list) pair in arrival order.  merge_check needs them
// ordered by (distance, offset) and only the K best; doing that inside its per-query stage loop would
// serialise up to w sorts behind one warp, so the slots are ordered here, one warp per slot, all in parallel.
constexpr int SS_WARPS = 8;
constexpr int SS_MAX = 512;  // widest slot (entries)

__global__ void __launch_bounds__(SS_WARPS * 32) slot_sort_kernel(RoundParams rp, long nslots) {
    __shared__ unsigned long long s_key[SS_WARPS][SS_MAX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = rp.K, cap = rp.cap, metric = rp.metric;
    unsigned long long* key = s_key[warp];
    const long nwarps = (long)gridDim.x * SS_WARPS;
    for (long base = ((long)blockIdx.x * SS_WARPS + warp) * 32; base < nslots; base += nwarps * 32) {
        // 32 slot counts per coalesced load, then the slots that need work one after the other
        const long mine = base + lane;
        const int cnt_l = mine < nslots ? rp.slot_cnt[mine] : 0;
        const bool todo_l = !(cnt_l & SLOT_SORTED) && (cnt_l & ~SLOT_SORTED) > 1 && (cnt_l & ~SLOT_SORTED) <= cap;
        unsigned todo = __ballot_sync(0xffffffffu, todo_l);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const long slot = base + src;
            const int rc = __shfl_sync(0xffffffffu, cnt_l, src);
            float* cd = rp.cand_d + slot * cap;
            unsigned* co = rp.cand_off + slot * cap;
            if (rc <= 32) {
                unsigned long long kk = ~0ull;
                if (lane < rc) {
                    uint32_t o = f2ord(cd[lane]);
                    if (metric == METRIC_IP) o = ~o;
                    kk = ((unsigned long long)o << 32) | co[lane];
                }
#pragma unroll
                for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                    for (int j = k2 >> 1; j > 0; j >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xffffffffu, kk, j);
                        const bool up = ((lane & k2) == 0), lower = ((lane & j) == 0);
                        const unsigned long long mn = kk < other ? kk : other, mx = kk < other ? other : kk;
                        kk = (lower == up) ? mn : mx;
                    }
                if (lane < min(rc, K)) {
                    uint32_t o = (uint32_t)(kk >> 32);
                    if (metric == METRIC_IP) o = ~o;
                    cd[lane] = ord2f(o);
                    co[lane] = (unsigned)(kk & 0xffffffffu);
                }
            } else {
                int P = 64;
                while (P < rc) P <<= 1;
                for (int i = lane; i < P; i += 32) {
                    unsigned long long kk = ~0ull;
                    if (i < rc) {
                        uint32_t o = f2ord(cd[i]);
                        if (metric == METRIC_IP) o = ~o;
                        kk = ((unsigned long long)o << 32) | co[i];
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
                for (int i = lane; i < min(rc, K); i += 32) {
                    const unsigned long long kk = key[i];
                    uint32_t o = (uint32_t)(kk >> 32);
                    if (metric == METRIC_IP) o = ~o;
                    cd[i] = ord2f(o);
                    co[i] = (unsigned)(kk & 0xffffffffu);
                }
                __syncwarp();
            }
            if (lane == 0) rp.slot_cnt[slot] = min(rc, K) | SLOT_SORTED;
        }
    }
}

// Exact rounds with few queries split every list into S segments x nsub row subsets so that all SMs have
// work; each (query, rank) then owns S * nsub partial results.  Merging them inside merge_check would put up
// to w * S * nsub dependent merges behind one warp (batch 1: ~1000), so they are merged here first, one warp
// per (query, rank): the K best by (distance, offset) end up in the pair's first sub-slot.
constexpr int SM_WARPS = 8;

__global__ void __launch_bounds__(SM_WARPS * 32) stage_merge_kernel(RoundParams rp, int KP) {
    __shared__ unsigned long long s_key[SM_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* key = s_key[warp];
    const int K = rp.K, cap = rp.cap, metric = rp.metric, nseg = rp.S * rp.nsub;
    const long npairs = (long)rp.n_active * rp.w;
    for (long pidx = (long)blockIdx.x * SM_WARPS + warp; pidx < npairs; pidx += (long)gridDim.x * SM_WARPS) {
        const long slot0 = pidx * nseg;
        int total = 0;
        bool any = false;
        for (int sgm = 0; sgm < nseg; sgm += 32) {  // 32 counts per load
            const int c = sgm + lane < nseg ? (rp.slot_cnt[slot0 + sgm + lane] & ~SLOT_SORTED) : 0;
            any |= __any_sync(0xffffffffu, c > 0 && sgm + lane > 0);
        }
        if (!any) continue;  // nothing outside the first sub-slot
        int c0 = min(rp.slot_cnt[slot0] & ~SLOT_SORTED, K);
        for (int i = lane; i < KP; i += 32) {
            unsigned long long kk = ~0ull;
            if (i < c0) {
                uint32_t o = f2ord(rp.cand_d[slot0 * cap + i]);
                if (metric == METRIC_IP) o = ~o;
                kk = ((unsigned long long)o << 32) | rp.cand_off[slot0 * cap + i];
            }
            key[i] = kk;
        }
        total = c0;
        for (int sgm = 1; sgm < nseg; sgm++) {
            const long sl = slot0 + sgm;
            const int c = min(rp.slot_cnt[sl] & ~SLOT_SORTED, K);
            if (c == 0) continue;
            for (int i = lane; i < KP; i += 32) {  // second run reversed: the 2 KP keys form a bitonic sequence
                unsigned long long kk = ~0ull;
                if (i < c) {
                    uint32_t o = f2ord(rp.cand_d[sl * cap + i]);
                    if (metric == METRIC_IP) o = ~o;
                    kk = ((unsigned long long)o << 32) | rp.cand_off[sl * cap + i];
                }
                key[2 * KP - 1 - i] = kk;
            }
            __syncwarp();
            for (int stride = KP; stride > 0; stride >>= 1) {
                for (int t = lane; t < KP; t += 32) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const unsigned long long x = key[lo], y = key[hi];
                    if (x > y) {
                        key[lo] = y;
                        key[hi] = x;
                    }
                }
                __syncwarp();
            }
            total = min(K, total + c);
            if (lane == 0) rp.slot_cnt[sl] = 0;
        }
        __syncwarp();
        for (int i = lane; i < total; i += 32) {
            const unsigned long long kk = key[i];
            uint32_t o = (uint32_t)(kk >> 32);
            if (metric == METRIC_IP) o = ~o;
            rp.cand_d[slot0 * cap + i] = ord2f(o);
            rp.cand_off[slot0 * cap + i] = (unsigned)(kk & 0xffffffffu);
        }
        if (lane == 0) rp.slot_cnt[slot0] = total | SLOT_SORTED;
        __syncwarp();
    }
}

void launch_stage_merge(const RoundParams& rp, int num_sms, cudaStream_t s) {
    const long npairs = (long)rp.n_active * rp.w;
    if (npairs == 0 || rp.S * rp.nsub <= 1) return;
    int KP = 16;
    while (KP < rp.K) KP <<= 1;
    const unsigned blocks = (unsigned)std::min<long>((npairs + SM_WARPS - 1) / SM_WARPS, (long)num_sms * 8);
    stage_merge_kernel<<<blocks, SM_WARPS * 32, 0, s>>>(rp, KP);
    CUDA_CHECK(cudaGetLastError());
}

void launch_slot_sort(const RoundParams& rp, int num_sms, cudaStream_t s) {
    const long nslots = (long)rp.n_active * rp.w * rp.S * rp.nsub;
    if (nslots == 0) return;
    AUNCEL_CHECK(rp.cap <= SS_MAX, "slot capacity too large");
    const unsigned blocks = (unsigned)std::min<long>((nslots + SS_WARPS * 32 - 1) / (SS_WARPS * 32), (long)num_sms * 8);
    slot_sort_kernel<<<blocks, SS_WARPS * 32, 0, s>>>(rp, nslots);
    CUDA_CHECK(cudaGetLastError());
}

// --------------------------------------------------------------------------- warp-level networks in registers
// 32 E keys, E per lane (lane l owns positions E l .. E l + E - 1).  Compare-exchange distances below E stay in
// the lane, the others are a shuffle pair: a lone warp pays no shared-memory round trip and no barrier per stage.
template <int E>
__device__ __forceinline__ void warp_merge_reg(unsigned long long (&e)[E], int lane) {  // bitonic input -> ascending
#pragma unroll
    for (int lm = 16; lm > 0; lm >>= 1) {
        const bool keep_min = (lane & lm) == 0;
#pragma unroll
        for (int r = 0; r < E; r++) {
            const unsigned lo32 = __shfl_xor_sync(0xffffffffu, (unsigned)e[r], lm);
            const unsigned hi32 = __shfl_xor_sync(0xffffffffu, (unsigned)(e[r] >> 32), lm);
            const unsigned long long o = ((unsigned long long)hi32 << 32) | lo32;
            e[r] = keep_min ? (o < e[r] ? o : e[r]) : (o > e[r] ? o : e[r]);
        }
    }
#pragma unroll
    for (int st = E / 2; st > 0; st >>= 1)
#pragma unroll
        for (int r = 0; r < E; r++)
            if ((r & st) == 0) {
                const unsigned long long a = e[r], b = e[r | st];
                if (a > b) {
                    e[r] = b;
                    e[r | st] = a;
                }
            }
}

template <int E>
__device__ __forceinline__ void warp_sort_reg(unsigned long long (&e)[E], int lane) {  // any input -> ascending
    const int base = lane * E;
#pragma unroll 1
    for (int size = 2; size <= 32 * E; size <<= 1) {
        int stride = size >> 1;
        for (; stride >= E; stride >>= 1) {
            const int lm = stride / E;
            const bool keep_min = ((lane & lm) == 0) == ((base & size) == 0);
#pragma unroll
            for (int r = 0; r < E; r++) {
                const unsigned lo32 = __shfl_xor_sync(0xffffffffu, (unsigned)e[r], lm);
                const unsigned hi32 = __shfl_xor_sync(0xffffffffu, (unsigned)(e[r] >> 32), lm);
                const unsigned long long o = ((unsigned long long)hi32 << 32) | lo32;
                e[r] = keep_min ? (o < e[r] ? o : e[r]) : (o > e[r] ? o : e[r]);
            }
        }
#pragma unroll
        for (int st = E / 2; st > 0; st >>= 1) {
            if (st <= stride) {
#pragma unroll
                for (int r = 0; r < E; r++) {
                    if ((r & st) == 0) {
                        const bool up = ((base + r) & size) == 0;
                        const unsigned long long a = e[r], b = e[r | st];
                        if ((a > b) == up) {
                            e[r] = b;
                            e[r | st] = a;
                        }
                    }
                }
            }
        }
    }
}

// the 2 KP keys of a merge (ascending held results ++ descending candidates) -> ascending, through registers
template <int KP>
__device__ __forceinline__ void merge_keys(unsigned long long* key, int lane) {
    constexpr int E = 2 * KP / 32;
    unsigned long long e[E];
#pragma unroll
    for (int r = 0; r < E; r++) e[r] = key[lane * E + r];
    warp_merge_reg<E>(e, lane);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < E; r++) key[lane * E + r] = e[r];
    __syncwarp();
}

// error_pro::cur_num (IVF_pro.cpp:258-291) evaluated by a whole warp.  The reference probes U(phi(D[j])) for
// j = query_k - 1 and then along a binary search; every probe is 15 arccos-LUT terms summed in order
// (sum_angle, :162-177) plus a bucket search (Trace::search, :84-107) -- scalar work that would otherwise run
// once per probe on a full warp.  Here lane j computes U for rank j (all ranks at once, same arithmetic and
// summation order per rank), then the reference's decisions are replayed on the precomputed values.  Error
// bits are taken only from the ranks the reference would have probed.
__device__ __forceinline__ unsigned cur_num_warp(const ErrModelView& m, const float* D, const float* dtb, int index,
                                                 unsigned query_k, int* err, int lane) {
    const int start = (1 << index) - 1;
    float u[4];    // ranks lane, lane + 32, ... (query_k <= MAX_K = 128)
    int e_l[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
        u[b] = 0.f;
        e_l[b] = 0;
        const unsigned j = b * 32 + lane;
        if (b * 32 < (int)query_k) {
            if (j < query_k) u[b] = model_U(m, index, sum_angle(D[j], dtb, 15, start, m.arcos, m.arcos_size, &e_l[b]));
        }
    }
    int used_err = 0;
    auto probe = [&](long j) -> float {  // value and error bits of rank j, as if computed now
        float v = 0.f;
        int e = 0;
#pragma unroll
        for (int b = 0; b < 4; b++)
            if ((j >> 5) == b) {
                v = __shfl_sync(0xffffffffu, u[b], (int)(j & 31));
                e = __shfl_sync(0xffffffffu, e_l[b], (int)(j & 31));
            }
        used_err |= e;
        return v;
    };
    long high = (long)query_k - 1, low = 0;
    unsigned result;
    {
        const float uh = probe(high);
        // size_t*float -> float; size_t*1.005 -> double
        if ((double)fmul((float)query_k, uh) <= dmul((double)query_k, 1.005)) {
            *err |= used_err;
            return query_k;
        }
    }
    result = 0xffffffffu;
    while (low <= high) {
        const long middle = (low + high) / 2;
        if (middle <= 0) {
            result = 0;
            break;
        }
        const float um = probe(middle);
        if (fmul((float)(middle + 1), um) <= (float)query_k)
            low = middle + 1;
        else
            high = middle - 1;
    }
    *err |= used_err;
    return result == 0xffffffffu ? (unsigned)(low + 1) : result;
}

// --------------------------------------------------------------------------- merge + check
constexpr int MC_WARPS = 4;
#ifdef MC_MIN_BLOCKS
#define MC_BOUNDS __launch_bounds__(MC_WARPS * 32, MC_MIN_BLOCKS)
#else
#define MC_BOUNDS __launch_bounds__(MC_WARPS * 32)
#endif
constexpr int MC_INSERT_MAX = 8;  // slots with at most this many candidates are merged by insertion

template <int KP>
struct MergeSmem {
    unsigned long long key[2 * KP];
    unsigned long long code[2][KP];
    unsigned long long ccode[KP];
    float Rd[KP];
    float ang[KP];
};

template <int KP>
__global__ void MC_BOUNDS merge_check_kernel(RoundParams rp, TuneParams tp) {
    __shared__ MergeSmem<KP> sm_all[MC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    MergeSmem<KP>& sm = sm_all[warp];
    // persistent warps: queries are handed out through a counter, so a warp that drew a short
    // (decided, nothing to merge) query takes the next one instead of idling until the wave ends
    for (;;) {  // (body not re-indented)
    int a = 0;
    if (lane == 0) a = atomicAdd(&rp.ctl[CTL_MERGE_NEXT], 1);
    a = __shfl_sync(0xffffffffu, a, 0);
    if (a >= rp.n_active) break;
    const int q = rp.active[a];
    const int K = rp.K, metric = rp.metric;
    const float neut = neutral(metric);

    int rcnt = rp.st.rcnt[q];
    int bound = rp.st.bound[q], decided = rp.st.decided[q];
    const int limit = rp.st.limit[q], cut = rp.st.cut[q];
    int stoped = rp.st.stoped[q];
    float pre_val = rp.st.pre_val[q];
    unsigned long long mynp = rp.st.mynp[q];
    int cur = 0;  // code ping-pong
    int err = 0;

    for (int i = lane; i < KP; i += 32) {
        sm.Rd[i] = i < rcnt ? rp.st.Rd[(long)q * K + i] : neut;
        sm.code[0][i] = i < rcnt ? rp.st.Rcode[(long)q * K + i] : 0ull;
    }
    __syncwarp();

    const float* dtb = tp.dtb ? tp.dtb + (long)q * tp.max_num : nullptr;

    unsigned nzmask = 0;
    int grp_base = -1, cnt_grp = 0, flg_grp = 0;  // slot counts / redo flags of the current group of 32 stages
    // cur_num (IVF_pro.cpp:258-291) reads only the first query_topk entries of the sorted heap, the
    // boundary distances and the trace of `ind`: its value is reused until one of them changes
    int cached_ind = -1, topq_dirty = 1;
    unsigned cached_pre = 0;
    const int qk_i = tp.query_topk;
    for (int p_rel = 0; p_rel < rp.w; p_rel++) {
        const int stage = rp.r0 + p_rel + 1;
        if (stage > bound) break;
        // ---- merge the candidates of probe rank stage-1 (all segments)
        const int nseg = rp.S * rp.nsub;
        if (decided && tp.mode == 1 && nseg == 1 && stage < bound) {
            // ---- bulk path.  The stop stage of this query is known, the tune block no longer runs (:615-632),
            // so the stages before the last one only merge: their candidates are gathered (up to KP at a time),
            // ordered by (distance, arrival) -- arrival = (rank, offset), which is what one merge per stage
            // would produce -- and merged in one go.  The last stage takes the regular path (break / profile).
            const int p_end = min(rp.w, bound - rp.r0 - 1);  // stages [p_rel, p_end) are < bound
            int T = 0, p = p_rel;
            int cnt_l = 0, flg_l = 0;
            bool stop = false;
            auto flush = [&]() {
                if (T == 0) return;
                // candidate keys (ord << 32 | KP + arrival) are in sm.key[KP .. KP + T): sort them descending
                // over the whole upper half (pads = largest first), the lower half gets the held results
                for (int i = KP + T + lane; i < 2 * KP; i += 32) sm.key[i] = ~0ull;
                __syncwarp();
                if (KP >= 32) {
                    constexpr int E2 = KP >= 32 ? KP / 32 : 1;
                    unsigned long long e2[E2];
#pragma unroll
                    for (int r = 0; r < E2; r++) e2[r] = sm.key[KP + lane * E2 + r];
                    warp_sort_reg<E2>(e2, lane);
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < E2; r++) sm.key[2 * KP - 1 - (lane * E2 + r)] = e2[r];  // descending
                    __syncwarp();
                } else {
                    for (int size = 2; size <= KP; size <<= 1)
                        for (int stride = size >> 1; stride > 0; stride >>= 1) {
                            for (int t = lane; t < KP / 2; t += 32) {
                                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                                const bool down = ((lo & size) == 0);  // descending overall
                                const unsigned long long x = sm.key[KP + lo], y = sm.key[KP + hi];
                                if ((x < y) == down) {
                                    sm.key[KP + lo] = y;
                                    sm.key[KP + hi] = x;
                                }
                            }
                            __syncwarp();
                        }
                }
                for (int i = lane; i < KP; i += 32) {
                    unsigned long long k1 = ~0ull;
                    if (i < rcnt) {
                        uint32_t o = f2ord(sm.Rd[i]);
                        if (metric == METRIC_IP) o = ~o;
                        k1 = ((unsigned long long)o << 32) | (unsigned)i;
                    }
                    sm.key[i] = k1;
                }
                __syncwarp();
                merge_keys<KP>(sm.key, lane);
                rcnt = min(K, rcnt + T);
                for (int i = lane; i < KP; i += 32) {
                    if (i < rcnt) {
                        const unsigned long long k = sm.key[i];
                        uint32_t o = (uint32_t)(k >> 32);
                        if (metric == METRIC_IP) o = ~o;
                        const unsigned idx = (unsigned)(k & 0xffffffffu);
                        sm.Rd[i] = ord2f(o);
                        sm.code[cur ^ 1][i] = idx < (unsigned)KP ? sm.code[cur][idx] : sm.ccode[idx - KP];
                    } else {
                        sm.Rd[i] = neut;
                    }
                }
                cur ^= 1;
                T = 0;
                __syncwarp();
            };
            while (p < p_end && !stop) {
                if (p == p_rel || (p & 31) == 0) {
                    const int pl = (p & ~31) + lane;
                    cnt_l = pl < rp.w ? rp.slot_cnt[(long)a * rp.w + pl] : 0;
                    flg_l = (pl < rp.w && rp.pair_flag) ? rp.pair_flag[(long)a * rp.w + pl] : 0;
                }
                // next stage with candidates in this group of 32
                const unsigned nz = __ballot_sync(0xffffffffu, (cnt_l & ~SLOT_SORTED) > 0) & (~0u << (p & 31));
                const int grp_end = min(p_end, (p & ~31) + 32);
                if (nz == 0 || (p & ~31) + (__ffs(nz) - 1) >= grp_end) {
                    p = grp_end;
                    continue;
                }
                p = (p & ~31) + (__ffs(nz) - 1);
                const int raw = __shfl_sync(0xffffffffu, cnt_l, p & 31);
                const int flg = __shfl_sync(0xffffffffu, flg_l, p & 31);
                if (flg || !(raw & SLOT_SORTED)) {  // redone pair / unordered slot: regular path
                    stop = true;
                    break;
                }
                int c = min(raw & ~SLOT_SORTED, K);
                const long slot = (long)a * rp.w + p;
                if (rcnt == K && c > 4) {
                    // only the prefix that beats the K-th value held at the last flush can still enter
                    const float kth = sm.Rd[K - 1];
                    int nb = 0;
                    for (int i = lane; i < c; i += 32) {
                        const float v = rp.cand_d[slot * rp.cap + i];
                        nb += (metric == METRIC_L2 ? v < kth : v > kth) ? 1 : 0;
                    }
                    for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, o);
                    c = nb;
                }
                if (T + c > KP) flush();
                for (int i = lane; i < c; i += 32) {
                    uint32_t o = f2ord(rp.cand_d[slot * rp.cap + i]);
                    if (metric == METRIC_IP) o = ~o;
                    sm.key[KP + T + i] = ((unsigned long long)o << 32) | (unsigned)(KP + T + i);
                    sm.ccode[T + i] = ((unsigned long long)(unsigned)(rp.r0 + p) << 32) | rp.cand_off[slot * rp.cap + i];
                }
                T += c;
                p++;
            }
            flush();
            if (p > p_rel) {
                p_rel = p - 1;  // the loop increment moves on to stage p
                grp_base = -1;  // (the bulk path keeps its own group registers)
                continue;
            }
        }
        // One slot per pair (tensor-core rounds; exact rounds after stage_merge_kernel): the counts and
        // redo flags of 32 consecutive stages come from one coalesced load each, so a stage costs no
        // dependent global load of its own unless it has candidates.
        const bool single = nseg == 1 || rp.merged;
        int raw0 = 0, flagged = 0;
        if (single) {
            if ((p_rel & ~31) != grp_base) {
                grp_base = p_rel & ~31;
                const int pl = grp_base + lane;
                cnt_grp = pl < rp.w ? rp.slot_cnt[((long)a * rp.w + pl) * nseg] : 0;
                flg_grp = (pl < rp.w && rp.pair_flag) ? rp.pair_flag[(long)a * rp.w + pl] : 0;
                nzmask = __ballot_sync(0xffffffffu, cnt_grp > 0);
            }
            raw0 = __shfl_sync(0xffffffffu, cnt_grp, p_rel & 31);
            flagged = __shfl_sync(0xffffffffu, flg_grp, p_rel & 31);
            // a decided query has nothing to do at a stage without candidates (the tune block cannot change
            // its state any more, :615-632)
            if (raw0 == 0 && decided && stage < bound) continue;
        }
        // flagged pair of a tensor-core round: its candidates are the exact redo's, in the compact pool
        const long pidx = (long)a * rp.w + p_rel;
        const bool redo = rp.redo_ord != nullptr && (single ? flagged != 0 : (rp.pair_flag != nullptr && rp.pair_flag[pidx] != 0));
        const int nseg_s = (single && raw0 == 0) ? 0 : redo ? 4 : (rp.merged ? 1 : nseg);  // merged: one slot per pair
        const int cap = redo ? K : rp.cap;
        float* const cand_d = redo ? rp.redo_d : rp.cand_d;
        unsigned* const cand_off = redo ? rp.redo_off : rp.cand_off;
        const int* const slot_cnt = redo ? rp.redo_cnt : rp.slot_cnt;
        for (int seg = 0; seg < nseg_s; seg++) {
            const long slot = redo ? (long)rp.redo_ord[pidx] * 4 + seg : pidx * nseg + seg;
            const int raw_cnt = (single && !redo) ? raw0 : slot_cnt[slot];
            int rc = min(raw_cnt & ~SLOT_SORTED, cap);  // entries present (unsorted wide slots: up to cap)
            if ((raw_cnt & SLOT_SORTED) && rcnt == K && rc > MC_INSERT_MAX) {
                // The slot was filled against the threshold of the round's START; stages merged since then
                // have tightened the K-th value.  Only candidates that beat it NOW can enter (the strict
                // test of IndexIVFFlat.cpp:129), and in an ordered slot those are a prefix: usually a
                // handful, which go through the insertion path instead of a 2 KP merge.
                const float kth = sm.Rd[K - 1];
                int nb = 0;
                for (int i = lane; i < min(rc, K); i += 32) {
                    const float v = cand_d[slot * cap + i];
                    nb += (metric == METRIC_L2 ? v < kth : v > kth) ? 1 : 0;
                }
                for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, o);
                rc = nb;
            }
            const int c = min(rc, K);                          // entries that can matter
            if (c == 0) continue;
            const int rcnt_before = rcnt;
            const float kth_before = (qk_i >= 1 && rcnt >= qk_i) ? sm.Rd[qk_i - 1] : neut;
            if (rc <= MC_INSERT_MAX) {
                // A handful of candidates (the usual case after the tensor-core filter): insert them one
                // by one into the sorted top-k instead of a 2*KP bitonic merge.  A candidate goes behind
                // every held value <= it, like the (value, arrival) order of the merge below.
                unsigned long long kk = ~0ull;
                if (lane < rc) {
                    uint32_t o = f2ord(cand_d[slot * cap + lane]);
                    if (metric == METRIC_IP) o = ~o;
                    kk = ((unsigned long long)o << 32) | cand_off[slot * cap + lane];
                }
                if (!(raw_cnt & SLOT_SORTED)) {
#pragma unroll
                    for (int k2 = 2; k2 <= MC_INSERT_MAX; k2 <<= 1)
#pragma unroll
                        for (int j = k2 >> 1; j > 0; j >>= 1) {
                            unsigned long long other = __shfl_xor_sync(0xffffffffu, kk, j);
                            bool up = ((lane & k2) == 0), lower = ((lane & j) == 0);
                            unsigned long long mn = kk < other ? kk : other, mx = kk < other ? other : kk;
                            kk = (lower == up) ? mn : mx;
                        }
                }
                constexpr int EPL = (KP + 31) / 32;
                float best = 0.f;
                for (int t = 0; t < c; t++) {
                    const unsigned long long ck = __shfl_sync(0xffffffffu, kk, t);
                    const uint32_t co = (uint32_t)(ck >> 32);
                    const float cd = ord2f(metric == METRIC_IP ? ~co : co);
                    if (t == 0) best = cd;
                    int pos = 0;
#pragma unroll
                    for (int j = 0; j < EPL; j++) {
                        const int i = lane + 32 * j;
                        bool le = false;
                        if (i < rcnt) {
                            uint32_t o = f2ord(sm.Rd[i]);
                            if (metric == METRIC_IP) o = ~o;
                            le = o <= co;
                        }
                        pos += __popc(__ballot_sync(0xffffffffu, le));
                    }
                    if (pos >= K) break;  // the heap is full of values <= this one; the rest is no better
                    float nd[EPL];
                    unsigned long long nc[EPL];
#pragma unroll
                    for (int j = 0; j < EPL; j++) {
                        const int i = lane + 32 * j;
                        if (i < KP) {
                            const int src = i > pos ? i - 1 : i;
                            nd[j] = sm.Rd[src];
                            nc[j] = sm.code[cur][src];
                            if (i == pos) {
                                nd[j] = cd;
                                nc[j] = ((unsigned long long)(unsigned)(stage - 1) << 32) | (unsigned)(ck & 0xffffffffu);
                            }
                        }
                    }
                    __syncwarp();
                    rcnt = min(K, rcnt + 1);
#pragma unroll
                    for (int j = 0; j < EPL; j++) {
                        const int i = lane + 32 * j;
                        if (i < KP && i >= pos && i < rcnt) {
                            sm.Rd[i] = nd[j];
                            sm.code[cur][i] = nc[j];
                        }
                    }
                    __syncwarp();
                }
                if (tp.mode == 1 && qk_i >= 1) {
                    if (rcnt_before < qk_i || (metric == METRIC_L2 ? best < kth_before : best > kth_before)) topq_dirty = 1;
                }
                continue;
            }
            if (!(raw_cnt & SLOT_SORTED)) {
                // rerank_kernel (tensor-core rounds) appends survivors in arrival order and the exact
                // scan hands over short lists as they are: order them by (distance, offset), in place
                if (rc <= 32) {
                    unsigned long long kk = ~0ull;
                    if (lane < rc) {
                        uint32_t o = f2ord(cand_d[slot * cap + lane]);
                        if (metric == METRIC_IP) o = ~o;
                        kk = ((unsigned long long)o << 32) | cand_off[slot * cap + lane];
                    }
#pragma unroll
                    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                        for (int j = k2 >> 1; j > 0; j >>= 1) {
                            unsigned long long other = __shfl_xor_sync(0xffffffffu, kk, j);
                            bool up = ((lane & k2) == 0), lower = ((lane & j) == 0);
                            unsigned long long mn = kk < other ? kk : other, mx = kk < other ? other : kk;
                            kk = (lower == up) ? mn : mx;
                        }
                    if (lane < c) {
                        uint32_t o = (uint32_t)(kk >> 32);
                        if (metric == METRIC_IP) o = ~o;
                        cand_d[slot * cap + lane] = ord2f(o);
                        cand_off[slot * cap + lane] = (unsigned)(kk & 0xffffffffu);
                    }
                    __syncwarp();
                } else {
                    // wide slots (first tensor-core round) hold up to cap <= 2 KP survivors: sort them all,
                    // the K best go on
                    const int P = rc <= KP ? KP : 2 * KP;
                    for (int i = lane; i < P; i += 32) {
                        unsigned long long kk = ~0ull;
                        if (i < rc) {
                            uint32_t o = f2ord(cand_d[slot * cap + i]);
                            if (metric == METRIC_IP) o = ~o;
                            kk = ((unsigned long long)o << 32) | cand_off[slot * cap + i];
                        }
                        sm.key[i] = kk;
                    }
                    __syncwarp();
                    for (int size = 2; size <= P; size <<= 1)
                        for (int stride = size >> 1; stride > 0; stride >>= 1) {
                            for (int t = lane; t < P / 2; t += 32) {
                                int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                                bool up = ((lo & size) == 0);
                                unsigned long long x = sm.key[lo], y = sm.key[hi];
                                if ((x > y) == up) {
                                    sm.key[lo] = y;
                                    sm.key[hi] = x;
                                }
                            }
                            __syncwarp();
                        }
                    for (int i = lane; i < c; i += 32) {
                        unsigned long long kk = sm.key[i];
                        uint32_t o = (uint32_t)(kk >> 32);
                        if (metric == METRIC_IP) o = ~o;
                        cand_d[slot * cap + i] = ord2f(o);
                        cand_off[slot * cap + i] = (unsigned)(kk & 0xffffffffu);
                    }
                    __syncwarp();
                }
            }
            for (int i = lane; i < KP; i += 32) {
                unsigned long long k1 = ~0ull, k2 = ~0ull;
                if (i < rcnt) {
                    uint32_t o = f2ord(sm.Rd[i]);
                    if (metric == METRIC_IP) o = ~o;
                    k1 = ((unsigned long long)o << 32) | (unsigned)i;
                }
                if (i < c) {
                    uint32_t o = f2ord(cand_d[slot * cap + i]);
                    if (metric == METRIC_IP) o = ~o;
                    k2 = ((unsigned long long)o << 32) | (unsigned)(KP + i);
                    sm.ccode[i] = ((unsigned long long)(unsigned)(stage - 1) << 32) | cand_off[slot * cap + i];
                }
                sm.key[i] = k1;
                sm.key[2 * KP - 1 - i] = k2;
            }
            __syncwarp();
            merge_keys<KP>(sm.key, lane);
            rcnt = min(K, rcnt + c);
            for (int i = lane; i < KP; i += 32) {
                if (i < rcnt) {
                    unsigned long long k = sm.key[i];
                    uint32_t o = (uint32_t)(k >> 32);
                    if (metric == METRIC_IP) o = ~o;
                    unsigned idx = (unsigned)(k & 0xffffffffu);
                    sm.Rd[i] = ord2f(o);
                    sm.code[cur ^ 1][i] = idx < (unsigned)KP ? sm.code[cur][idx] : sm.ccode[idx - KP];
                } else {
                    sm.Rd[i] = neut;
                }
            }
            cur ^= 1;
            __syncwarp();
            if (tp.mode == 1 && qk_i >= 1) {
                // did a candidate enter the first query_topk positions?  (its best one must beat the
                // old query_topk-th value; an equal value leaves the sorted values unchanged)
                const float best = cand_d[slot * cap];
                if (rcnt_before < qk_i || (metric == METRIC_L2 ? best < kth_before : best > kth_before)) topq_dirty = 1;
            }
        }

        const bool cut_here = cut && stage == limit;  // max_codes break precedes both blocks (:541)
        if (cut_here) break;

        if (tp.mode == 1 && !tp.overhead_profile) {
            if (!decided) {
                // ---- tune block, IndexIVF.cpp:551-626
                const int ind = stage_to_ind((size_t)stage, (size_t)rp.nlist);
                const unsigned qk = (unsigned)tp.query_topk;
                unsigned pre = cached_pre;
                if (topq_dirty || ind != cached_ind) {
                    const float* S = sm.Rd;
                    if (metric == METRIC_IP) {
                        for (int i = lane; i < K; i += 32)
                            sm.ang[i] = arcos_lookup(tp.model.arcos, tp.model.arcos_size, sm.Rd[i], &err);
                        __syncwarp();
                        S = sm.ang;
                    }
                    pre = cur_num_warp(tp.model, S, dtb, ind, qk, &err, lane);
                    cached_pre = pre;
                    cached_ind = ind;
                    topq_dirty = 0;
                }
                float recall = fdiv((float)pre, (float)qk);
                const float ext = rcnt == K ? sm.Rd[K - 1] : neut;  // heap extreme (:573-587)
                const float req = tp.require_acc[q];
                const unsigned long long stops = (unsigned long long)fmul(req, 12.f);
                if (stage > 1) {
                    stoped = (ext == pre_val) ? stoped + 1 : 0;
                    if ((unsigned long long)stoped >= stops) recall = 1.f;
                }
                pre_val = ext;
                if ((recall >= req && mynp == 0) || (stage >= rp.nlist / 8 && mynp == 0)) {
                    mynp = (unsigned long long)fmul((float)stage, tp.model.multipler);
                    if (mynp >= (unsigned long long)rp.nlist && tp.t_recalls && lane == 0) tp.t_recalls[q] = 1.f;
                }
                if (mynp != 0) {
                    decided = 1;
                    unsigned long long b = mynp > (unsigned long long)stage ? mynp : (unsigned long long)stage;
                    bound = (int)(b < (unsigned long long)limit ? b : (unsigned long long)limit);
                }
                __syncwarp();
            }
            if (decided && mynp != 0 && mynp <= (unsigned long long)stage) {
                // the break of :627-632; with `profile` the true recall is logged first
                if (tp.profile && tp.t_recalls && tp.gt_kth) {
                    const float kd = tp.gt_kth[q];
                    int cnt = 0;
                    for (int i = lane; i < K; i += 32) {
                        float v = sm.Rd[i];
                        bool hit = metric == METRIC_L2 ? ((double)v <= dmul((double)kd, 1.0005))
                                                       : ((double)v >= dmul((double)kd, 0.9995));
                        cnt += hit ? 1 : 0;
                    }
                    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                    if (lane == 0) tp.t_recalls[q] = fdiv((float)cnt, (float)tp.query_topk);
                }
                bound = stage;
                break;
            }
        } else if (tp.mode == 2) {
            // ---- calibration snapshot, IndexIVF.cpp:640-655
            if (stage <= rp.nlist / 8 && (stage & (stage - 1)) == 0) {
                int ind = 0;
                while ((1 << ind) != stage) ind++;
                float* out = tp.snapshots + ((long)q * tp.n_traces + ind) * K;
                for (int i = lane; i < K; i += 32) out[i] = sm.Rd[i];
            }
        }
    }

    // ---- write back
    for (int i = lane; i < rcnt; i += 32) {
        rp.st.Rd[(long)q * K + i] = sm.Rd[i];
        rp.st.Rcode[(long)q * K + i] = sm.code[cur][i];
    }
    if (lane == 0) {
        rp.st.rcnt[q] = rcnt;
        rp.st.bound[q] = bound;
        rp.st.decided[q] = decided;
        rp.st.stoped[q] = stoped;
        rp.st.pre_val[q] = pre_val;
        rp.st.mynp[q] = mynp;
        rp.st.tau[q] = rcnt == K ? sm.Rd[K - 1] : neut;
        if (err) atomicOr(&rp.ctl[CTL_ERR], err);
    }
    __syncwarp();
    }  // next query
}

void launch_merge_check(const RoundParams& rp, const TuneParams& tp, cudaStream_t s) {
    if (rp.n_active == 0) return;
    unsigned blocks = (unsigned)((rp.n_active + MC_WARPS - 1) / MC_WARPS);
    static int resident_blocks = 0;
    if (!resident_blocks) {
        int dev = 0, sms = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        resident_blocks = sms * 8;
    }
    blocks = std::min<unsigned>(blocks, (unsigned)resident_blocks);
    CUDA_CHECK(cudaMemsetAsync(rp.ctl + CTL_MERGE_NEXT, 0, sizeof(int), s));
    if (rp.K <= 16) merge_check_kernel<16><<<blocks, MC_WARPS * 32, 0, s>>>(rp, tp);
    else if (rp.K <= 32) merge_check_kernel<32><<<blocks, MC_WARPS * 32, 0, s>>>(rp, tp);
    else if (rp.K <= 64) merge_check_kernel<64><<<blocks, MC_WARPS * 32, 0, s>>>(rp, tp);
    else merge_check_kernel<128><<<blocks, MC_WARPS * 32, 0, s>>>(rp, tp);
    CUDA_CHECK(cudaGetLastError());
}

// --------------------------------------------------------------------------- active list
__global__ void compact_active_kernel(RoundParams rp, int r1, int* active_out) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= rp.n_active) return;
    int q = rp.active[a];
    if (rp.st.bound[q] > r1) {
        int pos = atomicAdd(&rp.ctl[CTL_N_ACTIVE], 1);
        active_out[pos] = q;
        atomicMin(&rp.ctl[CTL_MIN_RCNT], rp.st.rcnt[q]);
        if (rp.st.rcnt[q] < rp.K) atomicAdd(&rp.ctl[CTL_NOT_FULL], 1);  // heaps still holding neutral slots
        atomicAdd(&rp.ctl[CTL_REM_SUM], min(rp.st.bound[q] - r1, 4096));  // (the host sizes the next round's tiles with it)
    }
}

void launch_compact_active(const RoundParams& rp, int r1, int* active_out, int* h_ctl_pinned,
                           cudaStream_t s) {
    CUDA_CHECK(cudaMemsetAsync(rp.ctl + CTL_N_ACTIVE, 0, sizeof(int), s));
    CUDA_CHECK(cudaMemsetAsync(rp.ctl + CTL_MIN_RCNT, 0x7f, sizeof(int), s));
    CUDA_CHECK(cudaMemsetAsync(rp.ctl + CTL_NOT_FULL, 0, sizeof(int), s));
    CUDA_CHECK(cudaMemsetAsync(rp.ctl + CTL_REM_SUM, 0, sizeof(int), s));
    if (rp.n_active > 0) {
        compact_active_kernel<<<(unsigned)((rp.n_active + 255) / 256), 256, 0, s>>>(rp, r1, active_out);
        CUDA_CHECK(cudaGetLastError());
    }
    CUDA_CHECK(cudaMemcpyAsync(h_ctl_pinned, rp.ctl, CTL_SIZE * sizeof(int), cudaMemcpyDeviceToHost, s));
}

// --------------------------------------------------------------------------- finalize
// heap_reorder output (Heap.h:295-322): best first, padded with neutral / -1; labels from the
// inverted lists' id arrays; IndexIVFStats ndis / nlist (IndexIVF.cpp:676,731-734).
__global__ void finalize_kernel(RoundParams rp, TuneParams tp, float* D, long long* I,
                                unsigned long long* mynp_out, unsigned long long* stats) {
    long q = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (q >= rp.n) return;
    const int K = rp.K;
    const int rcnt = rp.st.rcnt[q];
    const float neut = neutral(rp.metric);
    for (int i = lane; i < K; i += 32) {
        float dv = neut;
        long long id = -1;
        if (i < rcnt) {
            dv = rp.st.Rd[q * K + i];
            unsigned long long code = rp.st.Rcode[q * K + i];
            int p = (int)(code >> 32);
            unsigned off = (unsigned)(code & 0xffffffffu);
            int l = rp.ckeys[q * rp.nlist + p];
            if (rp.shard_rank >= 0) {
                // the vector lives in exactly one shard; the others report -1 and an all-reduce(max) follows
                const int sh = (int)(off >> SHARD_CODE_SHIFT);
                off &= (1u << SHARD_CODE_SHIFT) - 1u;
                id = sh == rp.shard_rank ? rp.ids[rp.list_off[l] + off] : -1;
            } else {
                id = rp.ids[rp.list_off[l] + off];
            }
        }
        D[q * K + i] = dv;
        I[q * K + i] = id;
    }
    // stats: lists visited / codes scanned up to the stop stage
    const int stop = rp.st.bound[q];
    unsigned long long nd = 0, nl = 0;
    for (int p = lane; p < stop; p += 32) {
        int l = rp.ckeys[q * rp.nlist + p];
        long long sz = rp.list_off[l + 1] - rp.list_off[l];
        nd += (unsigned long long)sz;
        nl += sz > 0 ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        nd += __shfl_xor_sync(0xffffffffu, nd, o);
        nl += __shfl_xor_sync(0xffffffffu, nl, o);
    }
    if (lane == 0) {
        atomicAdd(&stats[0], nl);
        atomicAdd(&stats[1], nd);
        if (mynp_out && tp.mode == 1 && !tp.overhead_profile) mynp_out[q] = rp.st.mynp[q];
    }
}

void launch_finalize(const RoundParams& rp, const TuneParams& tp, float* D, long long* I,
                     unsigned long long* mynp_out, unsigned long long* stats_dev, cudaStream_t s) {
    if (rp.n == 0) return;
    finalize_kernel<<<(unsigned)((rp.n * 32 + 127) / 128), 128, 0, s>>>(rp, tp, D, I, mynp_out, stats_dev);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
