// Tensor-core candidate filter on CTA PAIRS (tcgen05 cta_group::2).
//
// tcfilter.cu is bound by how fast list tiles reach shared memory: the resident query tile (256 x d f32 = 128 KB
// at d = 128) leaves room for five 16 KB stages, and a ring that short is latency-bound at ~5.3 TB/s for all SMs
// together, from HBM or from L2 alike (with seven / nine stages and narrower query tiles the same
// pipeline stages 6.5 / 7.2 TB/s).  Here two CTAs on neighbouring SMs run ONE tile together:
//   D[256 rows x N queries] = A[256 x K] * B[N x K]^T   issued by the leader CTA as tcgen05.mma.cta_group::2;
//   each CTA streams ITS 128 rows of every 256-row block of the list (A), holds HALF of the tile's queries (B, the
//   tensor cores read the other half from the peer's shared memory) and owns the accumulator rows of its A rows.
// Half a query tile is 64 KB, so the ring grows from five to nine 16 KB stages per SM at the same bytes per MMA.
// The bound, the survivor list, the epilogue and rerank_kernel are those of tcfilter.cu.
// MEASURED (B200, bench step, three filter rounds): 1.04 / 2.05 / 1.00 ms against 0.94 / 1.83 / 0.90 ms of tcfilter.cu,
// results bit-identical.  The ring no longer stalls the MMA thread (9 % waiting for a_full), and without the epilogue the
// launches take 0.87 / 1.69 / 0.87 ms -- the 2-3-tile round is at the tensor pipe's real TF32 rate (~175 cycles per
// N = 256 instruction), which no staging scheme improves; what the pair adds is cross-CTA latency on every accumulator
// hand-over and half as many independent tile streams.  Kept as option "tc_kernel" = 3 (tests cover it); the engine uses
// tcfilter.cu.  docs/TC_FILTER_VARIANTS.md has the comparison.
//
// Protocol (every barrier lives at the same shared-memory offset in both CTAs):
//   a_full[s], b_full     leader's copy only: both producers' TMA loads complete_tx on it (its shared::cluster address,
//                         `mapa`), the leader's producer arms it with the bytes of both
//   a_empty[s], b_empty, t_full[i]   both copies, signalled by the leader's tcgen05.commit ... multicast::cluster
//   t_empty[i]            leader's copy: its epilogue warps arrive on it, and the peer's forwarder thread once the peer's
//                         epilogue warps have arrived on the peer's local t_done[i]
//   mail                  the leader's scheduler claims a tile and posts its index into the peer's mailbox; both
//                         schedulers then decode the same tile on their own
// Warps (480 threads per CTA) as in tcfilter.cu: 0 TMA producer, 1 MMA issuer (works in the leader only), 2-13 epilogue,
// 14 tile scheduler.
#include <cuda.h>

#include "exact.cuh"
#include "scan.cuh"
#include "tcfilter.cuh"
#include "tc_ptx.cuh"

namespace auncel {

constexpr int T3_EG = 3;                       // epilogue groups of four warps
constexpr int T3_SCHED_WARP = 2 + 4 * T3_EG;
constexpr int T3_THREADS = 32 * (T3_SCHED_WARP + 1);
constexpr int T3_ASTAGES = 9;
constexpr int T3_A_BYTES = 128 * 128;          // 128 rows x 32 f32
constexpr int T3_B_MAX = 64 * 1024;            // this CTA's half of the query tile: nchunk x N/2 x 128 B
constexpr int T3_NMAX = 256;
constexpr int T3_RES = 256;
constexpr size_t T3_SMEM = 1024 + T3_B_MAX + (size_t)T3_ASTAGES * T3_A_BYTES + 2 * (T3_NMAX * 8 + 64);

struct TileMeta3 {
    int flags;  // 1 = no more tiles
    int nblk;   // 256-row blocks of the list
    int Qt;
    int pair0;
    long long row0;
    int L;
    int pad;
    float kq[T3_NMAX];  // per query column: the side of the filter test that does not depend on the row (see the scheduler)
};

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `target` of the pair
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned target) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(target));
    return r;
}
// TMA load whose bytes are counted on the LEADER's barrier (CUTLASS SM100_TMA_2SM_LOAD)
__device__ __forceinline__ void tma2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(mapa(s32(bar), 0))
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                              unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}
// arrive on the leader's copy of a barrier (from either CTA)
__device__ __forceinline__ void mb_arrive_leader(unsigned long long* b) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(s32(b), 0)) : "memory");
}
// wait with cluster-scope acquire (the barrier is arrived on by the other CTA)
__device__ __forceinline__ void mb_wait_cluster(unsigned long long* b, unsigned parity) {
    unsigned ok = 0;
    const unsigned a = s32(b);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!ok);
}

#ifdef TC3_DEBUG
__device__ __forceinline__ void mb_wait_dbg3(unsigned long long* b, unsigned parity, int site, unsigned t) {
    const unsigned a = s32(b);
    unsigned long long t_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    bool said = false;
    for (;;) {
        unsigned ok = 0;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (ok) return;
        unsigned long long t_now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
        if (!said && t_now - t_start > 300000000ull && (threadIdx.x & 31) == 0 && blockIdx.x < 2) {
            printf("tc_filter3 STUCK block %d warp %d site %d t %u parity %u bar %u\n", blockIdx.x, threadIdx.x >> 5, site, t, parity,
                   a & 0xffffffu);
            said = true;
        }
        if (t_now - t_start > 3000000000ull) __trap();
    }
}
#define MBW3(site, b, par) mb_wait_dbg3(b, par, site, t)
#elif defined(TC3_TIMING)
#define MBW3(site, b, par)                     \
    do {                                       \
        long long t0_ = clock64();             \
        mb_wait(b, par);                       \
        t3_wait[site] += clock64() - t0_;      \
    } while (0)
#else
#define MBW3(site, b, par) mb_wait(b, par)
#endif

template <int METRIC>
__global__ void __launch_bounds__(T3_THREADS, 1)
tc_filter3_kernel(RoundParams rp, TcArgs ta, const __grid_constant__ CUtensorMap amap,
                  const __grid_constant__ CUtensorMap bmap) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long a_full[T3_ASTAGES], a_empty[T3_ASTAGES], b_full, b_empty, t_full[2],
        t_empty[2], t_done[2], m_full[2], m_empty[2], mail_full[4];
    __shared__ int mail[4];
    __shared__ unsigned tmem_base_s;
    unsigned char* smem = smem_dyn + ((1024u - (s32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* Bsm = smem;
    unsigned char* Asm = smem + T3_B_MAX;
    TileMeta3* meta = reinterpret_cast<TileMeta3*>(Asm + (size_t)T3_ASTAGES * T3_A_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster_ctarank();
    const bool leader = rank == 0;
#ifdef TC3_TIMING
    long long t3_wait[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long t3_begin = clock64();
#endif
    const int N = ta.N, NH = ta.N / 2, dpad = rp.dpad;  // queries per tile / per CTA
    const int nchunk = (dpad + 31) / 32;

    if (tid == 0) {
        for (int s = 0; s < T3_ASTAGES; s++) {
            mb_init(&a_full[s], 1);
            mb_init(&a_empty[s], 1);
        }
        mb_init(&b_full, 1);
        mb_init(&b_empty, 1);
        for (int i = 0; i < 2; i++) {
            mb_init(&t_full[i], 1);
            mb_init(&t_empty[i], 4 * T3_EG + 1);  // (leader's copy) its epilogue warps + the peer's forwarder
            mb_init(&t_done[i], 4 * T3_EG);       // (peer's copy) its epilogue warps
            mb_init(&m_full[i], 1);
            mb_init(&m_empty[i], 2 + 4 * T3_EG);
        }
        for (int i = 0; i < 4; i++) mb_init(&mail_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // one allocation for the pair: the same warp of both CTAs issues it
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs are initialised before anything remote touches them
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;

    if (warp == T3_SCHED_WARP) {
        // =========================== tile scheduler ===========================
        const int total_tiles = rp.ctl[CTL_TOTAL_TILES];
        for (unsigned t = 0;; t++) {
            int T = 0;
            if (leader) {
                if (lane == 0) {
                    T = atomicAdd(&rp.ctl[CTL_TILE_COUNTER], 1);
                    // post the tile to the peer: value, then a release arrive on the peer's mailbox barrier
                    const unsigned rm = mapa(s32(&mail[t & 3]), 1), rb = mapa(s32(&mail_full[t & 3]), 1);
                    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(rm), "r"(T) : "memory");
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb) : "memory");
                }
            } else {
                mb_wait_cluster(&mail_full[t & 3], (t >> 2) & 1);  // (the tile index was written by the other CTA)
                if (lane == 0) T = *reinterpret_cast<volatile int*>(&mail[t & 3]);
            }
            T = __shfl_sync(0xffffffffu, T, 0);
            const int m = t & 1;
            TileMeta3* mt = &meta[m];
            int l = 0, cnt_l = 0, qt = 0, L = 0, Qt = 0, pair0 = 0, nblk = 0;
            long long L0 = 0;
            float cq[T3_NMAX / 32];
            if (T < total_tiles) {
                int lo = 0, hi = (int)rp.nlist;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (rp.list_tile_off[mid] <= T) lo = mid; else hi = mid;
                }
                l = lo;
                cnt_l = rp.list_pair_off[l + 1] - rp.list_pair_off[l];
                qt = T - rp.list_tile_off[l];
                L0 = rp.list_off[l];
                L = (int)(rp.list_off[l + 1] - L0);
                Qt = min(N, cnt_l - qt * N);
                pair0 = rp.list_pair_off[l] + qt * N;
                nblk = (L + 255) / 256;
                // Filter test of a (row, query) pair, rows' part on the right:
                //   L2:  ||v||^2 (1-c2) - 2 dot - c1 |q||v| < tau + c3|tau| - ||q||^2 (1-c2)
                //        <=  dot + kq > 0.5 ||v||^2 (1-c2),   kq = 0.5 (rhs + c1 |q| max|v|) + slack
                //   IP:  dot + c1/2 |q||v| > tau - c3|tau|   <=  dot > kq,   kq = tau - c3|tau| - c1/2 |q| max|v| - slack
                // with max|v| over the LIST instead of the row's own norm (only the margin grows, nothing the exact
                // test accepts is lost) and a slack for the roundings of this evaluation: one add and one compare
                // per pair instead of two FMAs and a compare.
                const float nmax = ta.list_nmax[l], snmax = sqrtf(nmax);
#pragma unroll
                for (int jj = 0; jj < T3_NMAX / 32; jj++) {
                    const int j = jj * 32 + lane;
                    float c = METRIC == METRIC_L2 ? -FLT_MAX : FLT_MAX;  // never passes
                    if (j < Qt) {
                        unsigned long long pr = rp.pairs[pair0 + j];
                        int q = rp.active[(int)(pr >> 32)];
                        float tau = rp.st.tau[q], nq = ta.qnorm[q];
                        if (METRIC == METRIC_L2) {
                            const float rhs = tau + ta.c3 * fabsf(tau) - nq * (1.f - ta.c2);
                            c = 0.5f * (rhs + ta.c1 * sqrtf(nq) * snmax) + (nq + nmax) * (1.f / 1048576.f);
                        } else {
                            const float sq = sqrtf(nq) * snmax;  // >= |dot|
                            c = tau - ta.c3 * fabsf(tau) - 0.5f * ta.c1 * sq - sq * (1.f / 1048576.f);
                        }
                    }
                    cq[jj] = c;
                }
            }
            MBW3(0, &m_empty[m], ((t >> 1) & 1) ^ 1);
            if (T >= total_tiles) {
                if (lane == 0) {
                    mt->flags = 1;
                    mb_arrive(&m_full[m]);
                }
                break;
            }
#pragma unroll
            for (int jj = 0; jj < T3_NMAX / 32; jj++)
                if (jj * 32 + lane < N) mt->kq[jj * 32 + lane] = cq[jj];
            if (lane == 0) {
                mt->flags = 0;
                mt->nblk = nblk;
                mt->Qt = Qt;
                mt->pair0 = pair0;
                mt->row0 = L0;
                mt->L = L;
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&m_full[m]);
        }
    } else if (warp == 0) {
        // =========================== TMA producer ===========================
        unsigned ita = 0;
        for (unsigned t = 0;; t++) {
            const int m = t & 1;
            MBW3(1, &m_full[m], (t >> 1) & 1);
            const int flags = meta[m].flags, nblk = meta[m].nblk, pair0 = meta[m].pair0, Qt = meta[m].Qt;
            const long long L0 = meta[m].row0;
            __syncwarp();
            if (lane == 0) mb_arrive(&m_empty[m]);
            if (flags) break;
            if (lane == 0) {
                // this CTA's half of the query columns the MMA uses: [rank * Nt / 2, (rank + 1) * Nt / 2)
                const int Nt = min(N, (Qt + 31) / 32 * 32);
                MBW3(2, &b_empty, (t & 1) ^ 1);
                if (leader) mb_expect_tx(&b_full, (unsigned)(2 * nchunk * NH * 128));
                for (int c = 0; c < nchunk; c++)
                    tma2d_2sm(Bsm + (size_t)c * NH * 128, &bmap, c * 32, pair0 + (int)rank * (Nt / 2), &b_full);
                for (int blk = 0; blk < nblk; blk++)
                    for (int c = 0; c < nchunk; c++, ita++) {
                        const int s = ita % T3_ASTAGES;
                        MBW3(3, &a_empty[s], ((ita / T3_ASTAGES) & 1) ^ 1);
                        if (leader) mb_expect_tx(&a_full[s], 2 * T3_A_BYTES);
                        tma2d_2sm(Asm + (size_t)s * T3_A_BYTES, &amap, c * 32, (int)(L0 + (long long)blk * 256 + rank * 128),
                                  &a_full[s]);
                    }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (leader) ===========================
        // idesc: D = f32, A = B = tf32, both K-major, N >> 3, M = 256 >> 4 (128 rows in each CTA)
        if (elect_one()) {
            const unsigned idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 4) << 24);
            const unsigned long long adesc0 = umma_desc(s32(Asm)), bdesc0 = umma_desc(s32(Bsm));
            const unsigned b_chunk_inc = (unsigned)(NH * 128) >> 4;
            unsigned s = 0, a_phase = 0, blkc = 0;
            for (unsigned t = 0;; t++) {
                const int m = t & 1;
                MBW3(4, &m_full[m], (t >> 1) & 1);
                const int flags = meta[m].flags, nblk = meta[m].nblk;
                const int Nt = min(N, (meta[m].Qt + 31) / 32 * 32);
                const unsigned idesc = idesc0 | ((unsigned)(Nt >> 3) << 17);
                mb_arrive(&m_empty[m]);
                if (flags) break;
                if (!leader) {
                    // Forwarder: the peer's epilogue warps arrive on a LOCAL barrier; this otherwise idle thread
                    // passes each completed buffer on to the leader.  (A release at cluster scope issued by the
                    // epilogue warps themselves has to order their survivor stores first: 58 % of the MMA thread's
                    // time went into waiting for t_empty that way.)
                    for (int blk = 0; blk < nblk; blk++, blkc++) {
                        const int buf = blkc & 1;
                        MBW3(10, &t_done[buf], (blkc >> 1) & 1);
                        mb_arrive_leader(&t_empty[buf]);
                    }
                    continue;
                }
                MBW3(5, &b_full, t & 1);
                for (int blk = 0; blk < nblk; blk++, blkc++) {
                    const int buf = blkc & 1;
#ifdef TC3_TIMING
                    {
                        long long t0_ = clock64();
                        mb_wait_cluster(&t_empty[buf], ((blkc >> 1) & 1) ^ 1);
                        t3_wait[6] += clock64() - t0_;
                    }
#else
                    mb_wait_cluster(&t_empty[buf], ((blkc >> 1) & 1) ^ 1);  // (one arrival comes from the other CTA)
#endif
                    tc_fence_after();
                    const unsigned d_tmem = tmem_base + buf * 256;
                    unsigned long long bdesc = bdesc0;
                    for (int c = 0; c < nchunk; c++, bdesc += b_chunk_inc) {
                        MBW3(7, &a_full[s], a_phase);
                        tc_fence_after();
                        const unsigned long long adesc = adesc0 + s * (unsigned)(T3_A_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 4; k++) umma_tf32_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (c | k) != 0);
                        umma_commit_2sm(&a_empty[s]);
                        if (++s == T3_ASTAGES) {
                            s = 0;
                            a_phase ^= 1;
                        }
                    }
                    umma_commit_2sm(&t_full[buf]);
                }
                umma_commit_2sm(&b_empty);
            }
        }
        __syncwarp();
    } else if (warp >= 2 && warp < T3_SCHED_WARP) {
        // =========================== epilogue ===========================
        const int wq = warp & 3;
        const int grp = (warp - 2) >> 2;
        unsigned blkc = 0;
        int res_pos = 0, res_end = 0;
        for (unsigned t = 0;; t++) {
            const int m = t & 1;
            MBW3(8, &m_full[m], (t >> 1) & 1);
            const TileMeta3* mt = &meta[m];
            if (mt->flags) {
                __syncwarp();
                if (lane == 0) mb_arrive(&m_empty[m]);
                break;
            }
            const int nblk = mt->nblk, L = mt->L, pair0 = mt->pair0;
            const bool dead = *reinterpret_cast<volatile int*>(&rp.ctl[CTL_OVERFLOW]) < 0;
            const int ncg = min(N, (mt->Qt + 31) / 32 * 32) / 32;
            const long long row0 = mt->row0;
            float nv_next = (nblk > 0 && (int)rank * 128 + wq * 32 + lane < L) ? ta.vnorm[row0 + (int)rank * 128 + wq * 32 + lane] : 0.f;
            for (int blk = 0; blk < nblk; blk++, blkc++) {
                const int buf = blkc & 1;
                const int v = blk * 256 + (int)rank * 128 + wq * 32 + lane;
                const bool valid = v < L;
                const float nv = nv_next;  // (fetched one block ahead: the load's latency stays off the t_full -> t_empty path)
                {
                    const int vn = v + 256;
                    nv_next = (blk + 1 < nblk && vn < L) ? ta.vnorm[row0 + vn] : 0.f;
                }
                const float nvh = METRIC == METRIC_L2 ? 0.5f * (nv * (1.f - ta.c2)) : 0.f;  // the row's side of the test
                MBW3(9, &t_full[buf], (blkc >> 1) & 1);
                tc_fence_after();
                for (int cg = ta.dry == 1 ? ncg : grp; cg < ncg; cg += T3_EG) {
                    unsigned r[32];
                    tmem_ld32(tmem_base + ((unsigned)(wq * 32) << 16) + buf * 256 + cg * 32, r);
                    unsigned hits = 0;
                    const float4* k4 = reinterpret_cast<const float4*>(mt->kq + cg * 32);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 k = k4[j >> 2];  // the constants of four query columns per shared-memory load
                        const float d0 = __uint_as_float(r[j]), d1 = __uint_as_float(r[j + 1]);
                        const float d2 = __uint_as_float(r[j + 2]), d3 = __uint_as_float(r[j + 3]);
                        bool p0, p1, p2, p3;
                        if (METRIC == METRIC_L2) {
                            p0 = __fadd_rn(d0, k.x) > nvh;
                            p1 = __fadd_rn(d1, k.y) > nvh;
                            p2 = __fadd_rn(d2, k.z) > nvh;
                            p3 = __fadd_rn(d3, k.w) > nvh;
                        } else {
                            p0 = d0 > k.x;
                            p1 = d1 > k.y;
                            p2 = d2 > k.z;
                            p3 = d3 > k.w;
                        }
                        hits |= ((p0 ? 1u : 0u) << j) | ((p1 ? 1u : 0u) << (j + 1)) | ((p2 ? 1u : 0u) << (j + 2)) |
                                ((p3 ? 1u : 0u) << (j + 3));
                    }
                    if (!valid || dead) hits = 0;
                    if (ta.dry == 2 && hits != 0x12345678u) hits = 0;
                    if (__any_sync(0xffffffffu, hits != 0)) {
                        const int mine = __popc(hits);
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int up = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += up;
                        }
                        const int total = __shfl_sync(0xffffffffu, incl, 31);
                        if (res_pos + total > res_end) {
                            for (int t2 = res_pos + lane; t2 < res_end; t2 += 32)
                                if ((unsigned)t2 < (unsigned)ta.cand_cap) ta.cand[t2] = ~0ull;
                            const int want = max(total, T3_RES);
                            int b2 = 0;
                            if (lane == 0) b2 = atomicAdd(&rp.ctl[CTL_NCAND], want);
                            res_pos = __shfl_sync(0xffffffffu, b2, 0);
                            res_end = res_pos + want;
                        }
                        int pos = res_pos + incl - mine;
                        res_pos += total;
                        while (hits) {
                            const int j = __ffs(hits) - 1;
                            hits &= hits - 1;
                            if ((unsigned)pos < (unsigned)ta.cand_cap)
                                ta.cand[pos] = ((unsigned long long)(unsigned)(pair0 + cg * 32 + j) << 32) | (unsigned)v;
                            else
                                rp.ctl[CTL_OVERFLOW] = -(1 << 30);
                            pos++;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {  // the leader's MMA thread reuses the buffer in both CTAs
                    if (leader) mb_arrive(&t_empty[buf]);
                    else mb_arrive(&t_done[buf]);
                }
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&m_empty[m]);
        }
        for (int t2 = res_pos + lane; t2 < res_end; t2 += 32)
            if ((unsigned)t2 < (unsigned)ta.cand_cap) ta.cand[t2] = ~0ull;
    }
#ifdef TC3_TIMING
    if (blockIdx.x < 2 && (warp <= 2 || warp == T3_SCHED_WARP)) {
        bool me = lane == 0;
        if (warp == 1) me = t3_wait[4] > 0;  // the elected lane
        if (me)
            printf("tc_filter3 block %d warp %d: total %lld | sched m_empty %lld | tma m_full %lld b_empty %lld a_empty %lld | mma m_full %lld b_full %lld t_empty %lld a_full %lld | epi m_full %lld t_full %lld\n",
                   blockIdx.x, warp, clock64() - t3_begin, t3_wait[0], t3_wait[1], t3_wait[2], t3_wait[3], t3_wait[4], t3_wait[5],
                   t3_wait[6], t3_wait[7], t3_wait[8], t3_wait[9]);
    }
#endif
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the other may still touch it
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// queries per tile of the CTA-pair kernel (each CTA keeps half of them resident)
int tc3_tile_queries(int dpad) {
    const int nchunk = (dpad + 31) / 32;
    int nh = T3_B_MAX / (nchunk * 128);
    nh = std::min(nh, T3_NMAX / 2) / 16 * 16;
    return 2 * nh;  // a multiple of 32
}

void launch_tc_filter3(const RoundParams& rp, const TcArgs& ta, const void* amap, const void* bmap_half, int num_sms,
                       cudaStream_t s) {
    AUNCEL_CHECK(ta.N == tc3_tile_queries(rp.dpad) && ta.N >= 32, "tc_filter3: tile size");
    auto kern = rp.metric == METRIC_L2 ? tc_filter3_kernel<METRIC_L2> : tc_filter3_kernel<METRIC_IP>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(num_sms / 2 * 2));
    cfg.blockDim = dim3(T3_THREADS);
    cfg.dynamicSmemBytes = T3_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, rp, ta, *reinterpret_cast<const CUtensorMap*>(amap),
                                  *reinterpret_cast<const CUtensorMap*>(bmap_half)));
}

}  // namespace auncel
