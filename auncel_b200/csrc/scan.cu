// Subsystems (2) + (3): the inverted-list scan with fused per-(query, list) top-k selection.
//
// Replaces IVFFlatScanner::scan_codes + heap_pop/heap_push
// (/root/reference/Auncel/IndexIVFFlat.cpp:117-137, Heap.h:88-142).  The reference walks
// one query through its lists; here every (query, probe-rank) pair of the current round is
// grouped by inverted list, so a list tile staged in shared memory serves up to 32 queries.
// Distances use the reference's exact arithmetic (exact.cuh).  Selection keeps, per
// (query, list segment), the K best candidates that beat the query's threshold tau (the K-th
// best distance it already holds): exactly the candidates the reference's strict
// `C::cmp(simi[0], dis)` test could ever accept (IndexIVFFlat.cpp:129).
//
// Kernel structure (sm_100a): persistent CTAs, warp specialised.
//   warp 8      producer: claims tiles (atomic counter), decodes them, and per k-chunk issues
//               one TMA 2D tile load (128 list rows x 32 floats, SWIZZLE_128B) plus one
//               cp.async.bulk row per query of the tile, all completing on the stage's mbarrier
//   warps 0..7  consumers: wait on the stage's `full` mbarrier, accumulate a 4-query x
//               4-vector register tile each (lanes <-> vectors, warps <-> queries), release the
//               stage on its `empty` mbarrier, and run selection privately -- no CTA barrier
//               in the steady state, so a warp that is sorting never stalls the others.
#include <cuda.h>

#include "exact.cuh"
#include "scan.cuh"
#include "tcfilter.cuh"

namespace auncel {

// ------------------------------------------------------------------------- planning
__global__ void plan_count_kernel(RoundParams rp) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long tot = (long)rp.n_active * rp.w;
    unsigned long long work = 0;
    if (idx < tot) {
        int a = (int)(idx / rp.w), p_rel = (int)(idx - (long)a * rp.w);
        int q = rp.active[a];
        int p = rp.r0 + p_rel;
        if (p < rp.st.bound[q]) {
            int l = rp.ckeys[(long)q * rp.nlist + p];
            long long sz = rp.list_off[l + 1] - rp.list_off[l];
            if (sz > 0 && !(rp.filtered && !rp.pair_flag[idx])) {  // IndexIVF.cpp:452-455; filtered: slot == idx
                atomicAdd(&rp.list_cnt[l], 1);
                if (!rp.filtered) work = (unsigned long long)sz;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) work += __shfl_xor_sync(0xffffffffu, work, o);
    if ((threadIdx.x & 31) == 0 && work) atomicAdd(rp.round_work, work);
}

// single block: exclusive scans over the lists
__global__ void __launch_bounds__(1024) plan_offsets_kernel(RoundParams rp) {
    __shared__ int s_pair[1024], s_tile[1024];
    __shared__ int carry_pair, carry_tile;
    if (threadIdx.x == 0) carry_pair = carry_tile = 0;
    __syncthreads();
    for (long base = 0; base < rp.nlist; base += 1024) {
        long l = base + threadIdx.x;
        int c = l < rp.nlist ? rp.list_cnt[l] : 0;
        int t = ((c + rp.qt - 1) / rp.qt) * rp.S;
        s_pair[threadIdx.x] = c;
        s_tile[threadIdx.x] = t;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int vp = threadIdx.x >= off ? s_pair[threadIdx.x - off] : 0;
            int vt = threadIdx.x >= off ? s_tile[threadIdx.x - off] : 0;
            __syncthreads();
            s_pair[threadIdx.x] += vp;
            s_tile[threadIdx.x] += vt;
            __syncthreads();
        }
        if (l < rp.nlist) {
            rp.list_pair_off[l] = carry_pair + s_pair[threadIdx.x] - c;
            rp.list_tile_off[l] = carry_tile + s_tile[threadIdx.x] - t;
            rp.list_cursor[l] = 0;
            if (c > 0 && !rp.filtered) {
                // vectors of the lists this round touches (compulsory traffic) and vectors staged
                // into shared memory (one pass per query tile)
                unsigned long long L = (unsigned long long)(rp.list_off[l + 1] - rp.list_off[l]);
                atomicAdd(rp.round_work + 1, L);
                atomicAdd(rp.round_work + 2, L * (unsigned long long)((c + rp.qt - 1) / rp.qt));
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry_pair += s_pair[1023];
            carry_tile += s_tile[1023];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rp.list_pair_off[rp.nlist] = carry_pair;
        rp.list_tile_off[rp.nlist] = carry_tile;
        rp.ctl[CTL_TOTAL_PAIRS] = carry_pair;
        rp.ctl[CTL_TOTAL_TILES] = carry_tile;
        rp.ctl[CTL_TILE_COUNTER] = 0;
    }
}

__global__ void plan_fill_kernel(RoundParams rp) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long tot = (long)rp.n_active * rp.w;
    if (idx >= tot) return;
    int a = (int)(idx / rp.w), p_rel = (int)(idx - (long)a * rp.w);
    int q = rp.active[a];
    int p = rp.r0 + p_rel;
    if (p >= rp.st.bound[q]) return;
    int l = rp.ckeys[(long)q * rp.nlist + p];
    if (rp.list_off[l + 1] == rp.list_off[l]) return;
    if (rp.filtered) {
        if (!rp.pair_flag[idx]) return;
        if (rp.redo_ord) rp.redo_ord[idx] = atomicAdd(&rp.ctl[CTL_REDO_N], 1);  // compact redo pool
        else rp.slot_cnt[idx] = 0;  // the exact scan rewrites this slot
    }
    int pos = rp.list_pair_off[l] + atomicAdd(&rp.list_cursor[l], 1);
    rp.pairs[pos] = ((unsigned long long)(unsigned)a << 32) | (unsigned)p_rel;
}

// queries of the round, gathered in pair order: the query tile of a scan tile is then one
// contiguous 32-row box that a single TMA load can fetch
__global__ void gather_queries_kernel(RoundParams rp) {
    // warp-stride loop over the pairs actually planned (a filtered plan keeps a small fraction of
    // n_active * w, so the grid is sized for the hardware, not for the worst case)
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long total = rp.ctl[CTL_TOTAL_PAIRS];
    for (long pos = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; pos < total; pos += nwarps) {
        const int a = (int)(rp.pairs[pos] >> 32);
        const float4* src = reinterpret_cast<const float4*>(rp.xq + (long long)rp.active[a] * rp.dpad);
        float4* dst = reinterpret_cast<float4*>(rp.xq_sorted + pos * rp.dpad);
        for (int c = lane; c < rp.dpad / 4; c += 32) dst[c] = src[c];
    }
}

void launch_plan(const RoundParams& rp, cudaStream_t s) {
    long tot = (long)rp.n_active * rp.w;
    CUDA_CHECK(cudaMemsetAsync(rp.list_cnt, 0, rp.nlist * sizeof(int), s));
    if (!rp.filtered) CUDA_CHECK(cudaMemsetAsync(rp.slot_cnt, 0, (size_t)tot * rp.S * rp.nsub * sizeof(int), s));
    if (rp.filtered && rp.redo_ord) CUDA_CHECK(cudaMemsetAsync(rp.ctl + CTL_REDO_N, 0, sizeof(int), s));
    unsigned blocks = (unsigned)((tot + 255) / 256);
    plan_count_kernel<<<blocks, 256, 0, s>>>(rp);
    plan_offsets_kernel<<<1, 1024, 0, s>>>(rp);
    plan_fill_kernel<<<blocks, 256, 0, s>>>(rp);
    gather_queries_kernel<<<(unsigned)std::min<long>((tot * 32 + 255) / 256, 148L * 64), 256, 0, s>>>(rp);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------- scan
constexpr int MAX_STAGES = 6;
constexpr int QLD = SCAN_DK;                      // query row in smem, floats (dense TMA box)
constexpr int VT_BYTES = SCAN_VT * SCAN_DK * 4;   // 16384: 128 rows x 128 B, 128B-swizzled by TMA
constexpr int QT_BYTES = SCAN_QT * QLD * 4;       // 4096
constexpr int HDR_BYTES = 512;
constexpr int STAGE_BYTES = VT_BYTES + QT_BYTES + HDR_BYTES + 512;  // 21504 = 21 * 1024
constexpr int CAP = 256;                          // candidate buffer per query (>= MAX_K + 32)
constexpr size_t scan_smem(int nc) { return 1024 + (size_t)(nc == 8 ? 6 : 3) * STAGE_BYTES + (size_t)nc * 4 * CAP * 8; }
static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024 B alignment SWIZZLE_128B needs");

struct StageHdr {
    int flags;       // 1 = end of work
    int first;       // first stage of a tile
    int last_chunk;  // last k-chunk of a vector block -> epilogue
    int last_iter;   // last stage of the tile -> write results
    int blk;         // vector block inside the segment
    int nk;          // valid floats in this k-chunk (multiple of 4, <= 32)
    int Qt;          // queries in the tile
    int v_begin;     // first list offset of the segment
    int nvec;        // vectors in the segment
    int pad[7];      // (the first eight ints are read by the consumers as two 16-byte loads)
    int slot[SCAN_QT];
    float tau[SCAN_QT];
};
static_assert(sizeof(StageHdr) <= HDR_BYTES, "header too large");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    const unsigned addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// warp-cooperative bitonic sort of N (power of two) 64-bit keys in shared memory
template <int N>
__device__ __forceinline__ void warp_sort_smem(unsigned long long* key, int lane) {
#pragma unroll 1
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int t = lane; t < N / 2; t += 32) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                }
            }
            __syncwarp();
        }
    }
}

// one key per lane, ascending across lanes
__device__ __forceinline__ unsigned long long warp_sort_reg(unsigned long long key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
            bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
            unsigned long long mn = key < other ? key : other, mx = key < other ? other : key;
            key = (lower == up) ? mn : mx;
        }
    }
    return key;
}

template <int METRIC>
__device__ __forceinline__ unsigned long long make_key(float d, unsigned off) {
    uint32_t o = f2ord(d);
    if (METRIC == METRIC_IP) o = ~o;
    return ((unsigned long long)o << 32) | off;
}
template <int METRIC>
__device__ __forceinline__ float key_dist(unsigned long long key) {
    uint32_t o = (uint32_t)(key >> 32);
    if (METRIC == METRIC_IP) o = ~o;
    return ord2f(o);
}

// one 32-dimension chunk of NQ queries x TV rows per lane, in the reference's accumulation order
template <int METRIC, int TV, int NQ>
__device__ __forceinline__ void dist_chunk(float (&acc)[4][TV][4], const float* sq, const unsigned char* sv, int nkc,
                                           int xr) {
    // few (query, row) pairs per lane: unroll deeper so enough loads are in flight
#pragma unroll(NQ * TV <= 2 ? 8 : NQ * TV <= 4 ? 4 : 2)
    for (int kc = 0; kc < nkc; kc++) {
        float4 a[NQ], b[TV];
#pragma unroll
        for (int i = 0; i < NQ; i++) a[i] = *reinterpret_cast<const float4*>(sq + i * QLD + kc * 4);
#pragma unroll
        for (int j = 0; j < TV; j++) b[j] = *reinterpret_cast<const float4*>(sv + j * (32 * 128) + ((kc ^ xr) << 4));
#pragma unroll
        for (int i = 0; i < NQ; i++)
#pragma unroll
            for (int j = 0; j < TV; j++) exact_step<METRIC>(acc[i][j], a[i], b[j]);
    }
}

// threshold as an unsigned key where smaller = tighter (atomicMin publishes improvements)
template <int METRIC>
__device__ __forceinline__ unsigned tau_key(float tau) {
    uint32_t o = f2ord(tau);
    return METRIC == METRIC_L2 ? o : ~o;
}
template <int METRIC>
__device__ __forceinline__ float key_tau(unsigned k) {
    return ord2f(METRIC == METRIC_L2 ? k : ~k);
}

// 256 keys, 8 per lane (lane l owns positions 8l .. 8l+7), sorted ascending in registers: compare-exchange
// distances below 8 never leave the lane, the others are one shuffle pair each -- no shared-memory round trip
// and no warp barrier per network stage, which is what a lone warp waits for in warp_sort_smem.
__device__ __forceinline__ void warp_sort_reg8(unsigned long long (&e)[8], int lane) {
    const int base = lane * 8;
#pragma unroll 1
    for (int size = 2; size <= 256; size <<= 1) {
        int stride = size >> 1;
        for (; stride >= 8; stride >>= 1) {
            const int lm = stride >> 3;
            const bool keep_min = ((lane & lm) == 0) == ((base & size) == 0);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const unsigned lo32 = __shfl_xor_sync(0xffffffffu, (unsigned)e[r], lm);
                const unsigned hi32 = __shfl_xor_sync(0xffffffffu, (unsigned)(e[r] >> 32), lm);
                const unsigned long long o = ((unsigned long long)hi32 << 32) | lo32;
                e[r] = keep_min ? (o < e[r] ? o : e[r]) : (o > e[r] ? o : e[r]);
            }
        }
#pragma unroll
        for (int st = 4; st > 0; st >>= 1) {
            if (st <= stride) {
#pragma unroll
                for (int r = 0; r < 8; r++) {
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

// sort the buffer, keep the K best, tighten tau when K are held
template <int METRIC>
__device__ __forceinline__ void compact(unsigned long long* buf, int& cnt, float& tau, int K, int lane) {
    if (cnt <= 64) {
        for (int i = cnt + lane; i < 64; i += 32) buf[i] = ~0ull;
        __syncwarp();
        warp_sort_smem<64>(buf, lane);
    } else {
        static_assert(CAP == 256, "warp_sort_reg8 sorts exactly 256 keys");
        unsigned long long e[8];
#pragma unroll
        for (int r = 0; r < 8; r++) e[r] = lane * 8 + r < cnt ? buf[lane * 8 + r] : ~0ull;
        warp_sort_reg8(e, lane);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; r += 2)
            *reinterpret_cast<ulonglong2*>(&buf[lane * 8 + r]) = make_ulonglong2(e[r], e[r + 1]);
    }
    if (cnt > K) cnt = K;
    if (cnt == K) tau = key_dist<METRIC>(buf[K - 1]);
    __syncwarp();
}

// RSPLIT = 1: warp w owns queries 4w..4w+3 and all 128 rows of a block (4 per lane); 32 queries/tile.
// RSPLIT = 2: 16 queries/tile; warp w owns queries 4(w/2)..+3 and 64 rows (2 per lane).
// RSPLIT = 4: 8 queries/tile; warp w owns queries 4(w/4)..+3 and 32 rows (1 per lane).
// Splitting the rows keeps all warps busy when a list is probed by few queries; the RSPLIT row
// subsets of a query are written as RSPLIT sub-slots.
// NC = consumer warps: 8 (one CTA per SM, 6 stages) or 4 (two CTAs per SM, 3 stages each).  With few queries
// per list a tile is a latency-bound chain of stages; two independent chains per SM nearly double the rate.
template <int METRIC, int RSPLIT, int NC>
__global__ void __launch_bounds__((NC + 1) * 32, NC == 8 ? 1 : 2)
scan_kernel(RoundParams rp, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap qmap) {
    constexpr int NCONS = NC;
    constexpr int STAGES = NC == 8 ? 6 : 3;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES];
    // thresholds shared by the RSPLIT warps that scan different rows for the same query: a warp
    // that has K candidates publishes its K-th best, the others filter with the tightest value.
    // Ring of 8 tiles (> STAGES, the furthest a warp can run ahead), reset by the producer.
    constexpr int TAU_RING = 8;
    static_assert(TAU_RING > STAGES + 1, "a consumer can run at most STAGES stages (hence < STAGES + 1 tiles) ahead "
                                         "of the slowest one: the threshold ring must be deeper than that");
    __shared__ unsigned s_tau[TAU_RING][SCAN_QT];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(smem + (size_t)STAGES * STAGE_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = rp.K, dpad = rp.dpad;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCONS) {
        // =========================== producer ===========================
        const int nchunk = (dpad + SCAN_DK - 1) / SCAN_DK;
        const int total_tiles = rp.ctl[CTL_TOTAL_TILES];
        unsigned it = 0, tcount = 0;
        while (true) {
            int T = 0;
            if (lane == 0) T = atomicAdd(&rp.ctl[CTL_TILE_COUNTER], 1);
            T = __shfl_sync(0xffffffffu, T, 0);
            if (T >= total_tiles) break;
            // decode tile -> (list, segment, query tile); segment-major so that neighbouring
            // tiles share list rows in L2
            int lo = 0, hi = (int)rp.nlist;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (rp.list_tile_off[mid] <= T) lo = mid; else hi = mid;
            }
            const int l = lo;
            const int cnt_l = rp.list_pair_off[l + 1] - rp.list_pair_off[l];
            const int QT = rp.qt;
            const int nqt = (cnt_l + QT - 1) / QT;
            const int tl = T - rp.list_tile_off[l];
            const int seg = tl / nqt, qt = tl - seg * nqt;
            const long long L0 = rp.list_off[l];
            const int L = (int)(rp.list_off[l + 1] - L0);
            int seg_len = (L + rp.S - 1) / rp.S;
            seg_len = (seg_len + 31) / 32 * 32;
            const int v_begin = seg * seg_len;
            const int v_end = min(L, v_begin + seg_len);
            if (v_begin >= v_end) continue;  // empty segment: slot_cnt stays 0
            const int Qt = min(QT, cnt_l - qt * QT);
            int slot = 0;
            float tau = 0.f;
            const int pair0 = rp.list_pair_off[l] + qt * QT;
            if (lane < Qt) {
                unsigned long long pr = rp.pairs[pair0 + lane];
                int a = (int)(pr >> 32), p_rel = (int)(pr & 0xffffffffu);
                slot = rp.redo_ord ? rp.redo_ord[a * rp.w + p_rel] : (a * rp.w + p_rel) * rp.S + seg;
                tau = rp.st.tau[rp.active[a]];
            }
            const int nvec = v_end - v_begin;
            const int nblk = (nvec + SCAN_VT - 1) / SCAN_VT;
            const long long row0 = L0 + v_begin;
            for (int blk = 0; blk < nblk; blk++) {
                for (int c = 0; c < nchunk; c++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    unsigned char* st = smem + (size_t)s * STAGE_BYTES;
                    StageHdr* h = reinterpret_cast<StageHdr*>(st + VT_BYTES + QT_BYTES);
                    const int k0 = c * SCAN_DK;
                    const int nk = min(SCAN_DK, dpad - k0);
                    if (lane == 0) {
                        h->flags = 0;
                        h->first = (blk == 0 && c == 0);
                        h->last_chunk = (c == nchunk - 1);
                        h->last_iter = (blk == nblk - 1 && c == nchunk - 1);
                        h->blk = blk;
                        h->nk = nk;
                        h->Qt = Qt;
                        h->v_begin = v_begin;
                        h->nvec = nvec;
                    }
                    h->slot[lane] = slot;
                    h->tau[lane] = tau;
                    if (blk == 0 && c == 0) s_tau[tcount & 7][lane] = tau_key<METRIC>(tau);
                    __syncwarp();
                    if (lane == 0) {
                        mbar_expect_tx(&full_bar[s], (unsigned)(VT_BYTES + QT_BYTES));
                        tma_load_2d(st, &tmap, k0, (int)(row0 + (long long)blk * SCAN_VT), &full_bar[s]);
                        tma_load_2d(st + VT_BYTES, &qmap, k0, pair0, &full_bar[s]);
                    }
                }
            }
            tcount++;
        }
        // end marker
        {
            const int s = it % STAGES;
            mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
            StageHdr* h = reinterpret_cast<StageHdr*>(smem + (size_t)s * STAGE_BYTES + VT_BYTES + QT_BYTES);
            if (lane == 0) {
                h->flags = 1;
                mbar_arrive(&full_bar[s]);
            }
        }
        return;
    }

    // =========================== consumers ===========================
    constexpr int TV = 4 / RSPLIT;                       // list rows per lane
    constexpr int NGROUP = NC / RSPLIT;                  // warps that share a row subset split the tile's queries
    const int group = warp / RSPLIT;
    const int rbase = (warp % RSPLIT) * (SCAN_VT / RSPLIT);  // first row of this warp (rows rbase + lane + 32 j)
    int cnt[4] = {0, 0, 0, 0};
    float tau[4] = {0.f, 0.f, 0.f, 0.f};
    float acc[4][TV][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < TV; j++)
#pragma unroll
            for (int x = 0; x < 4; x++) acc[i][j][x] = 0.f;
    const int xr = lane & 7;  // SWIZZLE_128B: 16-byte chunk index is XORed with (row & 7)
    unsigned tcount = 0;
    int tslot = 0;

#ifdef SCAN_TIMING
    long long tm[5] = {0, 0, 0, 0, 0}, tm_prev = clock64();  // wait, header, distance, filter, write-out
#define SCAN_TICK(k)                      \
    do {                                  \
        const long long now_ = clock64(); \
        tm[k] += now_ - tm_prev;          \
        tm_prev = now_;                   \
    } while (0)
#else
#define SCAN_TICK(k)
#endif
    for (unsigned it = 0;; it++) {
        const int s = it % STAGES;
        SCAN_TICK(4);
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        SCAN_TICK(0);
        const unsigned char* st = smem + (size_t)s * STAGE_BYTES;
        const StageHdr* h = reinterpret_cast<const StageHdr*>(st + VT_BYTES + QT_BYTES);
        const int4 h0 = *reinterpret_cast<const int4*>(&h->flags);  // flags, first, last_chunk, last_iter
        if (h0.x) break;
        const int4 h1 = *reinterpret_cast<const int4*>(&h->blk);    // blk, nk, Qt, v_begin
        // the tile's Qt queries are dealt evenly to the query groups (<= 4 each), so partially
        // filled tiles keep every warp busy instead of filling group 0 first
        const int Qt = h1.z;
        const int per = (Qt + NGROUP - 1) / NGROUP;
        const int q0 = group * per;
        const int nq = max(0, min(per, Qt - q0));
        const bool has_q = nq > 0;
        const int last_chunk = h0.z, last_iter = h0.w;
        const int blk = h1.x, nvec = h->nvec, v_begin = h1.w, nk = h1.y;
        int slot[4];
        if (h0.y) {
            tslot = tcount & 7;
            tcount++;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                tau[i] = h->tau[min(q0 + i, SCAN_QT - 1)];
                cnt[i] = 0;
            }
        }
        if (last_iter) {
#pragma unroll
            for (int i = 0; i < 4; i++) slot[i] = h->slot[min(q0 + i, SCAN_QT - 1)];
        }

        SCAN_TICK(1);
#ifdef SCAN_EXP_NODIST
        if (false) {
#else
        if (nq > 0) {
#endif
            const unsigned char* sv = st + (rbase + lane) * 128;
            const float* sq = reinterpret_cast<const float*>(st + VT_BYTES) + q0 * QLD;
            const int nkc = nk >> 2;
            switch (nq) {  // warp-uniform: only the queries this warp owns are computed
                case 1: dist_chunk<METRIC, TV, 1>(acc, sq, sv, nkc, xr); break;
                case 2: dist_chunk<METRIC, TV, 2>(acc, sq, sv, nkc, xr); break;
                case 3: dist_chunk<METRIC, TV, 3>(acc, sq, sv, nkc, xr); break;
                default: dist_chunk<METRIC, TV, 4>(acc, sq, sv, nkc, xr); break;
            }
        }
        SCAN_TICK(2);
        // the stage's data and header are consumed: hand the slot back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);

        if (has_q && last_chunk) {
            // ---- epilogue: filter against tau, append, compact when the buffer fills
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i >= nq) break;
                unsigned long long* buf = cand + (size_t)(warp * 4 + i) * CAP;
                const bool qok = true;
                if (RSPLIT > 1) {  // the tightest threshold any sibling warp has published
                    const float shared_tau = key_tau<METRIC>(s_tau[tslot][q0 + i]);
                    if (METRIC == METRIC_L2 ? shared_tau < tau[i] : shared_tau > tau[i]) tau[i] = shared_tau;
                }
#pragma unroll
                for (int j = 0; j < TV; j++) {
                    float dist = exact_finish(acc[i][j]);
                    acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
                    int v = blk * SCAN_VT + rbase + lane + 32 * j;
                    bool pass = qok && v < nvec && (METRIC == METRIC_L2 ? dist < tau[i] : dist > tau[i]);
#ifdef SCAN_EXP_NOSELECT
                    pass = pass && dist == -12345.f;
#endif
                    unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (m) {
                        if (cnt[i] + 32 > CAP) {
                            compact<METRIC>(buf, cnt[i], tau[i], K, lane);
                            if (RSPLIT > 1 && lane == 0) atomicMin(&s_tau[tslot][q0 + i], tau_key<METRIC>(tau[i]));
                            pass = pass && (METRIC == METRIC_L2 ? dist < tau[i] : dist > tau[i]);
                            m = __ballot_sync(0xffffffffu, pass);
                        }
                        if (pass) buf[cnt[i] + __popc(m & ((1u << lane) - 1))] = make_key<METRIC>(dist, (unsigned)(v_begin + v));
                        cnt[i] += __popc(m);
                        __syncwarp();
                    }
                }
            }
        }
        SCAN_TICK(3);
        if (has_q && last_iter) {
            // ---- per-(query, segment[, row subset]) result: sorted, at most K candidates
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i < nq && cnt[i] > 0) {
                    unsigned long long* buf = cand + (size_t)(warp * 4 + i) * CAP;
                    const long sl = (long)slot[i] * RSPLIT + (warp % RSPLIT);
                    if (cnt[i] <= 32) {
                        unsigned long long key = lane < cnt[i] ? buf[lane] : ~0ull;
                        key = warp_sort_reg(key, lane);
                        const int c = min(cnt[i], K);
                        if (lane < c) {
                            rp.cand_d[sl * rp.cap + lane] = key_dist<METRIC>(key);
                            rp.cand_off[sl * rp.cap + lane] = (unsigned)(key & 0xffffffffu);
                        }
                        if (lane == 0) rp.slot_cnt[sl] = c | SLOT_SORTED;
                    } else {
                        // up to K candidates go out as they are: sorting them here would stall the
                        // whole CTA's stage ring behind one warp, merge_check orders them instead
                        const bool sorted = cnt[i] > K || !rp.defer_sort;
                        if (sorted) compact<METRIC>(buf, cnt[i], tau[i], K, lane);
                        for (int t = lane; t < cnt[i]; t += 32) {
                            unsigned long long key = buf[t];
                            rp.cand_d[sl * rp.cap + t] = key_dist<METRIC>(key);
                            rp.cand_off[sl * rp.cap + t] = (unsigned)(key & 0xffffffffu);
                        }
                        if (lane == 0) rp.slot_cnt[sl] = cnt[i] | (sorted ? SLOT_SORTED : 0);
                    }
                    __syncwarp();
                }
            }
        }
    }
#ifdef SCAN_TIMING
    SCAN_TICK(4);
    if (blockIdx.x == 0 && lane == 0)
        printf("scan<%d> warp %d: wait %lld header %lld distance %lld filter %lld writeout %lld\n", RSPLIT, warp, tm[0],
               tm[1], tm[2], tm[3], tm[4]);
#endif
}

// -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_tensor_map_2d(void* out_map, const float* base, long long nrows, int dpad, int box_rows,
                               bool swizzle) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        AUNCEL_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
        fn = (EncodeTiledFn)p;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)dpad, (cuuint64_t)std::max<long long>(nrows, 1)};
    cuuint64_t gstride[1] = {(cuuint64_t)dpad * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)SCAN_DK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AUNCEL_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
}

void make_codes_tensor_map(void* out_map, const float* codes, long long nrows, int dpad, int box_rows) {
    make_tensor_map_2d(out_map, codes, nrows, dpad, box_rows, true);
}

void make_queries_tensor_map(void* out_map, const float* xq_sorted, long long nrows, int dpad) {
    make_tensor_map_2d(out_map, xq_sorted, nrows, dpad, SCAN_QT, false);
}

void make_queries_tensor_map_tc(void* out_map, const float* xq_sorted, long long nrows, int dpad, int N) {
    make_tensor_map_2d(out_map, xq_sorted, nrows, dpad, N, true);
}

void launch_scan(const RoundParams& rp, const void* tmap, const void* qmap, int num_sms, cudaStream_t s) {
    void (*kern)(RoundParams, const CUtensorMap, const CUtensorMap);
    const int nc = rp.nsub == 4 && rp.nc == 4 ? 4 : 8;
    if (rp.metric == METRIC_L2)
        kern = rp.nsub == 4 ? (nc == 4 ? scan_kernel<METRIC_L2, 4, 4> : scan_kernel<METRIC_L2, 4, 8>)
                            : rp.nsub == 2 ? scan_kernel<METRIC_L2, 2, 8> : scan_kernel<METRIC_L2, 1, 8>;
    else
        kern = rp.nsub == 4 ? (nc == 4 ? scan_kernel<METRIC_IP, 4, 4> : scan_kernel<METRIC_IP, 4, 8>)
                            : rp.nsub == 2 ? scan_kernel<METRIC_IP, 2, 8> : scan_kernel<METRIC_IP, 1, 8>;
    AUNCEL_CHECK(rp.qt == 4 * nc / rp.nsub || rp.nsub == 1, "scan tile shape and queries per tile disagree");
    const size_t smem = scan_smem(nc);
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<num_sms * (nc == 4 ? 2 : 1), (nc + 1) * 32, smem, s>>>(rp, *reinterpret_cast<const CUtensorMap*>(tmap),
                                                                 *reinterpret_cast<const CUtensorMap*>(qmap));
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
