// Tensor-core candidate filter, second layout: the QUERIES are the TMEM-resident operand.
//
// tcfilter.cu keeps a tile's queries in shared memory (128 KB at d = 128) and streams the list through the
// 80 KB that are left -- five 16 KB stages.  That ring is latency-bound: with 148 x 80 KB in flight the
// lists arrive at ~5.3 TB/s whether they come from HBM or from L2, and the round in which every list is
// probed by two query tiles (the tensor-bound one) stalls on a_full.  Here the roles
// of the MMA operands are swapped:
//   A (M = 128)  = 128 queries, read by the tensor core from TENSOR MEMORY (tcgen05.mma with a TMEM A operand);
//                  a tile holds G = 1 or 2 such groups, G * nchunk * 32 columns
//   B (N = NB)   = NB list rows (64 for G = 2, 128 for G = 1), K-major SWIZZLE_128B in shared memory
//   D            = [128 queries x NB rows] per group, 2 x G x NB TMEM columns (double buffered)
// so ALL shared memory (208 KB) is one ring of 8 / 16 KB stages.  The queries of a tile travel through the same
// ring (TMA, ahead of the tile's first list block) and four loader warps move them from the stage into
// TMEM with tcgen05.st once the previous tile's MMAs have retired.  Column budget (512): A at [0, G * KC),
// KC = nchunk * 32 <= 256 / G; accumulators at [256, 512).  d > 256 stays with tcfilter.cu.
// MEASURED (B200, bench step, three filter rounds): 1.95 / 3.13 / 1.57 ms against 0.97 / 1.96 / 0.97 ms of tcfilter.cu --
// results bit-identical, but SLOWER, also with one query group and N = 128 (1.90 / 2.91 / 1.46).  The wait
// attribution (-DTC2_TIMING) shows the MMA thread and the epilogue warps each busy ~60 % of the time and
// waiting for each other the rest: an A operand in TMEM costs 128 lanes x 8 columns x 4 B = 4 KB of tensor-memory
// reads per K = 8 instruction, on the read path the epilogue's tcgen05.ld (64 B per cycle) also uses; with only
// d / 8 = 16 instructions per accumulator the operand reads are as large as the accumulator reads.  The layout
// pays for GEMMs with a long K loop, not for this filter.  Kept as option "tc_kernel" = 2 (tests cover it);
// the engine uses tcfilter.cu.
// The bound, the survivor list and rerank_kernel are unchanged (tcfilter.cu); the epilogue is the
// transpose of the old one: a thread owns a QUERY (TMEM lane) and walks over 32 list rows per tcgen05.ld.
//
// Warps (480 threads): 0 TMA producer, 1 MMA issuer (one elected lane), 2-5 query loaders (one per TMEM lane
// quarter), 6-13 epilogue (two per lane quarter), 14 tile scheduler.
#include <cuda.h>

#include "exact.cuh"
#include "scan.cuh"
#include "tcfilter.cuh"
#include "tc_ptx.cuh"

namespace auncel {

constexpr int T2_LOAD_WARP0 = 2;
constexpr int T2_EPI_WARP0 = 6;
constexpr int T2_EPI_WARPS = 8;
constexpr int T2_SCHED_WARP = T2_EPI_WARP0 + T2_EPI_WARPS;
constexpr int T2_THREADS = 32 * (T2_SCHED_WARP + 1);
constexpr int T2_SLOT = 64 * 128;            // one TMA box: 64 rows x 32 f32
constexpr int T2_SLOTS = 26;                 // ring: 208 KB
constexpr int T2_RING = T2_SLOTS * T2_SLOT;
constexpr int T2_QMAX = 256;
constexpr int T2_ACC0 = 256;                 // first accumulator column
constexpr size_t T2_SMEM = 1024 + (size_t)T2_RING + 2 * (T2_QMAX * 8 + 64) + T2_EPI_WARPS * 32 * 8;

struct TileMeta2 {
    int flags;  // 1 = no more tiles
    int nblk;
    int Qt;
    int pair0;
    long long row0;  // first arena row of the list
    int L;           // list length
    int pad;
    float2 q[T2_QMAX];  // (c1 * ||q||, rhs) per query
};

// D[tmem] (+)= A[tmem] * B[smem descriptor]   (kind::tf32, one CTA)
__device__ __forceinline__ void umma_tf32_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long bdesc, unsigned idesc,
                                             unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(unsigned taddr, const unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#ifdef TC2_DEBUG
// debug builds: a wait that gives up after ~1 s, names its site and traps (finds protocol deadlocks)
__device__ __forceinline__ void mb_wait_dbg(unsigned long long* b, unsigned parity, int site, unsigned t) {
    const unsigned a = s32(b);
    unsigned long long t_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    bool said = false;
    for (;;) {
        unsigned ok = 0;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (ok) return;
        unsigned long long t_now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
        if (!said && t_now - t_start > 300000000ull && (threadIdx.x & 31) == 0 && blockIdx.x < 2) {
            printf("tc_filter2 STUCK block %d warp %d site %d t %u parity %u bar %u\n", blockIdx.x, threadIdx.x >> 5, site, t,
                   parity, a);
            said = true;
        }
        if (t_now - t_start > 3000000000ull) __trap();
    }
}
#define MBW(site, b, par) mb_wait_dbg(b, par, site, t)
#elif defined(TC2_TIMING)
// wait-time attribution: cycles spent at each wait site, printed by CTA 0 at the end
#define MBW(site, b, par)                      \
    do {                                       \
        long long t0_ = clock64();             \
        mb_wait(b, par);                       \
        t2_wait[site] += clock64() - t0_;      \
    } while (0)
#else
#define MBW(site, b, par) mb_wait(b, par)
#endif

// position in the stage ring: slot index and phase parity, advanced without divisions in the hot loops
struct RingPos {
    unsigned s = 0, ph = 0;
    __device__ __forceinline__ void step(unsigned ns) {
        if (++s == ns) {
            s = 0;
            ph ^= 1;
        }
    }
    __device__ __forceinline__ void skip(unsigned long long n, unsigned ns) {
        const unsigned long long tot = (unsigned long long)s + n;
        ph ^= (unsigned)((tot / ns) & 1);
        s = (unsigned)(tot % ns);
    }
};

template <int METRIC>
__global__ void __launch_bounds__(T2_THREADS, 1)
tc_filter2_kernel(RoundParams rp, TcArgs ta, const __grid_constant__ CUtensorMap amap,
                  const __grid_constant__ CUtensorMap qmap, int G, int NB) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long a_full[T2_SLOTS], a_empty[T2_SLOTS], q_full, q_free, t_full[2], t_empty[2],
        m_full[2], m_empty[2];
    __shared__ unsigned tmem_base_s;
    unsigned char* ring = smem_dyn + ((1024u - (s32(smem_dyn) & 1023u)) & 1023u);
    TileMeta2* meta = reinterpret_cast<TileMeta2*>(ring + T2_RING);
    float2* rowc_all = reinterpret_cast<float2*>(ring + T2_RING + 2 * (T2_QMAX * 8 + 64));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef TC2_TIMING
    long long t2_wait[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long t2_begin = clock64();
#endif
    const int dpad = rp.dpad;
    const int nchunk = (dpad + 31) / 32;
    const int KC = nchunk * 32;                 // TMEM columns of one query group
    const unsigned stage_bytes = (unsigned)NB * 128u;
    const unsigned NS = T2_RING / stage_bytes;  // 26 stages of 8 KB or 13 of 16 KB
    const int boxes = NB / 64;                  // TMA boxes per stage
    const int qhalves = 128 / NB;               // stages that hold one k-chunk of one query group
    const int Ntile = 128 * G;

    if (tid == 0) {
        for (int s = 0; s < T2_SLOTS; s++) {
            mb_init(&a_full[s], 1);
            mb_init(&a_empty[s], 1);
        }
        mb_init(&q_full, 4);
        mb_init(&q_free, 1);
        for (int i = 0; i < 2; i++) {
            mb_init(&t_full[i], 1);
            mb_init(&t_empty[i], T2_EPI_WARPS);
            mb_init(&m_full[i], 1);
            mb_init(&m_empty[i], 2 + 4 + T2_EPI_WARPS);  // TMA, MMA, loaders, epilogue
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;
#ifdef TC2_DEBUG
    if (tid == 0 && blockIdx.x < 4)
        printf("tc_filter2 block %d: a_full %u a_empty %u q_full %u q_free %u t_full %u t_empty %u m_full %u m_empty %u NS %u G %d NB %d nchunk %d tiles %d\n",
               blockIdx.x, s32(a_full), s32(a_empty), s32(&q_full), s32(&q_free), s32(t_full), s32(t_empty), s32(m_full),
               s32(m_empty), NS, G, NB, nchunk, rp.ctl[CTL_TOTAL_TILES]);
#endif

    if (warp == T2_SCHED_WARP) {
        // =========================== tile scheduler ===========================
        const int total_tiles = rp.ctl[CTL_TOTAL_TILES];
        for (unsigned t = 0;; t++) {
            int T = 0;
            if (lane == 0) T = atomicAdd(&rp.ctl[CTL_TILE_COUNTER], 1);
            T = __shfl_sync(0xffffffffu, T, 0);
            const int m = t & 1;
            TileMeta2* mt = &meta[m];
            int l = 0, cnt_l = 0, qt = 0, L = 0, Qt = 0, pair0 = 0, nblk = 0;
            long long L0 = 0;
            float2 cq[T2_QMAX / 32];
            if (T < total_tiles) {  // decode and gather the constants BEFORE waiting for the meta slot
                int lo = 0, hi = (int)rp.nlist;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (rp.list_tile_off[mid] <= T) lo = mid; else hi = mid;
                }
                l = lo;
                cnt_l = rp.list_pair_off[l + 1] - rp.list_pair_off[l];
                qt = T - rp.list_tile_off[l];  // S == 1 in tensor-core rounds
                L0 = rp.list_off[l];
                L = (int)(rp.list_off[l + 1] - L0);
                Qt = min(Ntile, cnt_l - qt * Ntile);
                pair0 = rp.list_pair_off[l] + qt * Ntile;
                nblk = (L + NB - 1) / NB;
#pragma unroll
                for (int jj = 0; jj < T2_QMAX / 32; jj++) {
                    const int j = jj * 32 + lane;
                    float2 c = make_float2(0.f, METRIC == METRIC_L2 ? -FLT_MAX : FLT_MAX);  // never passes
                    if (j < Qt) {
                        unsigned long long pr = rp.pairs[pair0 + j];
                        int q = rp.active[(int)(pr >> 32)];
                        float tau = rp.st.tau[q], nq = ta.qnorm[q];
                        if (METRIC == METRIC_L2) {
                            // pass <=> nv(1-c2) - 2 dot - c1 |q||v|  <  tau + c3|tau| - nq(1-c2)
                            c.x = ta.c1 * sqrtf(nq);
                            c.y = tau + ta.c3 * fabsf(tau) - nq * (1.f - ta.c2);
                        } else {
                            // pass <=> dot + c1/2 |q||v|  >  tau - c3|tau|
                            c.x = 0.5f * ta.c1 * sqrtf(nq);
                            c.y = tau - ta.c3 * fabsf(tau);
                        }
                    }
                    cq[jj] = c;
                }
            }
            MBW(0, &m_empty[m], ((t >> 1) & 1) ^ 1);
            if (T >= total_tiles) {
                if (lane == 0) {
                    mt->flags = 1;
                    mb_arrive(&m_full[m]);
                }
                break;
            }
#pragma unroll
            for (int jj = 0; jj < T2_QMAX / 32; jj++)
                if (jj * 32 + lane < Ntile) mt->q[jj * 32 + lane] = cq[jj];
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
        // ring order per tile: the query stages (group, k-chunk, row half), then the list (row block, k-chunk)
        RingPos rg;
        for (unsigned t = 0;; t++) {
            const int m = t & 1;
            MBW(1, &m_full[m], (t >> 1) & 1);
            const int flags = meta[m].flags, nblk = meta[m].nblk, pair0 = meta[m].pair0, Qt = meta[m].Qt;
            const long long L0 = meta[m].row0;
            __syncwarp();
            if (lane == 0) mb_arrive(&m_empty[m]);
            if (flags) break;
            if (lane == 0) {
                const int Gt = Qt > 128 ? 2 : 1;
                for (int g = 0; g < Gt; g++)
                    for (int c = 0; c < nchunk; c++)
                        for (int h = 0; h < qhalves; h++) {
                            MBW(2, &a_empty[rg.s], rg.ph ^ 1);
                            mb_expect_tx(&a_full[rg.s], stage_bytes);
                            for (int b = 0; b < boxes; b++)
                                tma2d(ring + (size_t)rg.s * stage_bytes + (size_t)b * T2_SLOT, &qmap, c * 32,
                                      pair0 + g * 128 + h * NB + b * 64, &a_full[rg.s]);
                            rg.step(NS);
                        }
                for (int blk = 0; blk < nblk; blk++)
                    for (int c = 0; c < nchunk; c++) {
                        MBW(3, &a_empty[rg.s], rg.ph ^ 1);
                        mb_expect_tx(&a_full[rg.s], stage_bytes);
                        for (int b = 0; b < boxes; b++)
                            tma2d(ring + (size_t)rg.s * stage_bytes + (size_t)b * T2_SLOT, &amap, c * 32,
                                  (int)(L0 + (long long)blk * NB + b * 64), &a_full[rg.s]);
                        rg.step(NS);
                    }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        // idesc: D = f32, A = B = tf32, K-major, N = NB (>> 3), M = 128 (>> 4)
        if (elect_one()) {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(NB >> 3) << 17) | ((128u >> 4) << 24);
            const unsigned long long bdesc0 = umma_desc(s32(ring));
            const unsigned stage_units = stage_bytes >> 4;  // descriptor address units are 16 bytes
            const unsigned acc_cols = (unsigned)(G * NB);   // one accumulator buffer
            RingPos rg;
            unsigned blkc = 0;
            for (unsigned t = 0;; t++) {
                const int m = t & 1;
                MBW(4, &m_full[m], (t >> 1) & 1);
                const int flags = meta[m].flags, nblk = meta[m].nblk, Qt = meta[m].Qt;
                mb_arrive(&m_empty[m]);
                if (flags) break;
                const int Gt = Qt > 128 ? 2 : 1;
                rg.skip((unsigned long long)Gt * nchunk * qhalves, NS);  // the tile's query stages belong to the loaders
                MBW(11, &q_full, t & 1);                              // the tile's queries are in TMEM
                tc_fence_after();
                for (int blk = 0; blk < nblk; blk++, blkc++) {
                    const unsigned buf = blkc & 1;
                    MBW(5, &t_empty[buf], ((blkc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const unsigned d0 = tmem_base + T2_ACC0 + buf * acc_cols;
                    for (int c = 0; c < nchunk; c++) {
                        MBW(6, &a_full[rg.s], rg.ph);
                        tc_fence_after();
                        const unsigned long long bdesc = bdesc0 + (unsigned long long)(rg.s * stage_units);
                        for (int g = 0; g < Gt; g++) {
                            const unsigned a0 = tmem_base + (unsigned)(g * KC + c * 32);
#pragma unroll
                            for (int k = 0; k < 4; k++)  // K = 8 tf32: 8 TMEM columns of A, 32 bytes of every B row
                                umma_tf32_ts(d0 + (unsigned)(g * NB), a0 + 8 * k, bdesc + 2 * k, idesc, (c | k) != 0);
                        }
                        umma_commit(&a_empty[rg.s]);  // frees the stage when these MMAs have read it
                        rg.step(NS);
                    }
                    umma_commit(&t_full[buf]);
                }
                umma_commit(&q_free);  // the tile's MMAs are done with the queries in TMEM
            }
        }
        __syncwarp();
    } else if (warp >= T2_LOAD_WARP0 && warp < T2_EPI_WARP0) {
        // =========================== query loaders ===========================
        // A query stage holds NB rows x 32 floats (SWIZZLE_128B).  The warp whose TMEM lane quarter the
        // rows belong to reads them (thread = row: eight conflict-free 128-bit loads), the stage is
        // released, and the 32 values go to TMEM columns [g * KC + c * 32, +32) of the thread's lane.
        const int wq = warp & 3;
        RingPos rg;
        for (unsigned t = 0;; t++) {
            const int m = t & 1;
            MBW(7, &m_full[m], (t >> 1) & 1);
            const int flags = meta[m].flags, nblk = meta[m].nblk, Qt = meta[m].Qt;
            __syncwarp();
            if (lane == 0) mb_arrive(&m_empty[m]);
            if (flags) break;
            const int Gt = Qt > 128 ? 2 : 1;
            // The previous tile's MMAs have read their queries -- and its list stages, which this warp skipped
            // without waiting.  Waiting here (not just before the first tcgen05.st) keeps the warp less than one
            // lap of the ring ahead of the producer; a parity wait cannot tell lap n from lap n + 2.
            MBW(12, &q_free, (t & 1) ^ 1);
            tc_fence_after();
            for (int g = 0; g < Gt; g++)
                for (int c = 0; c < nchunk; c++)
                    for (int h = 0; h < qhalves; h++) {
                        MBW(8, &a_full[rg.s], rg.ph);
                        const int r = wq * 32 + lane - h * NB;  // row inside the stage
                        const bool mine = r >= 0 && r < NB;     // (warp-uniform)
                        unsigned v[32];
                        if (mine) {
                            const unsigned char* row = ring + (size_t)rg.s * stage_bytes + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const uint4 x = *reinterpret_cast<const uint4*>(row + ((j ^ (r & 7)) << 4));
                                v[4 * j] = x.x;
                                v[4 * j + 1] = x.y;
                                v[4 * j + 2] = x.z;
                                v[4 * j + 3] = x.w;
                            }
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");  // all four loader warps are done with the stage
                        if (warp == T2_LOAD_WARP0 && lane == 0) mb_arrive(&a_empty[rg.s]);
                        rg.step(NS);
                        if (mine) {
                            tmem_st32(tmem_base + ((unsigned)(wq * 32) << 16) + (unsigned)(g * KC + c * 32), v);
                        }
                    }
            rg.skip((unsigned long long)nblk * nchunk, NS);  // the list stages belong to the MMA warp
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mb_arrive(&q_full);
        }
    } else if (warp >= T2_EPI_WARP0 && warp < T2_SCHED_WARP) {
        // =========================== epilogue ===========================
        // Two warps per TMEM lane quarter.  A unit = (query group g, 32 list rows): thread = query, the
        // 32 dots of its row chunk arrive with one tcgen05.ld; the rows' constants (||v||^2 (1 - c2), ||v||)
        // sit in a per-warp shared-memory line (lane j loads row j, one unit ahead).
        const int wq = warp & 3;
        const int sub = (warp - T2_EPI_WARP0) >> 2;
        float2* rowc = rowc_all + (warp - T2_EPI_WARP0) * 32;
        const int cpb = NB / 32;  // row chunks per block
        const unsigned acc_cols = (unsigned)(G * NB);
        unsigned blkc = 0;
        int res_pos = 0, res_end = 0;  // this warp's reserved range of the survivor list
        for (unsigned t = 0;; t++) {
            const int m = t & 1;
            MBW(9, &m_full[m], (t >> 1) & 1);
            const TileMeta2* mt = &meta[m];
            if (mt->flags) {
                __syncwarp();
                if (lane == 0) mb_arrive(&m_empty[m]);
                break;
            }
            const int nblk = mt->nblk, L = mt->L, pair0 = mt->pair0;
            const int Gt = mt->Qt > 128 ? 2 : 1;
            const long long row0 = mt->row0;
            const float2 cq0 = mt->q[wq * 32 + lane];
            const float2 cq1 = mt->q[(Gt - 1) * 128 + wq * 32 + lane];
            // once the survivor list is full the round is redone anyway: stop counting (checked once per tile)
            const bool dead = *reinterpret_cast<volatile int*>(&rp.ctl[CTL_OVERFLOW]) < 0;
            const int U = Gt * cpb;  // units per block; this warp takes sub, sub + 2, ...
            // norms one unit ahead
            auto load_nv = [&](int blk, int u) -> float {
                const int v = blk * NB + (u % cpb) * 32 + lane;
                return (blk < nblk && v < L) ? ta.vnorm[row0 + v] : 0.f;
            };
            float nv_next = load_nv(0, sub);
            for (int blk = 0; blk < nblk; blk++, blkc++) {
                const unsigned buf = blkc & 1;
                MBW(10, &t_full[buf], (blkc >> 1) & 1);
                tc_fence_after();
                for (int u = sub; u < U; u += 2) {
                    const float nv = nv_next;
                    nv_next = (u + 2 < U) ? load_nv(blk, u + 2) : load_nv(blk + 1, sub);
                    const int g = u / cpb, ch = u % cpb;
                    const int v0 = blk * NB + ch * 32;
                    const int nvalid = L - v0;
                    if (nvalid <= 0) continue;  // (warp-uniform)
                    rowc[lane] = make_float2(METRIC == METRIC_L2 ? nv * (1.f - ta.c2) : 0.f, sqrtf(nv));
                    __syncwarp();
                    unsigned r[32];
                    tmem_ld32(tmem_base + ((unsigned)(wq * 32) << 16) + T2_ACC0 + buf * acc_cols + (unsigned)(g * NB + ch * 32), r);
                    const float2 cq = g == 0 ? cq0 : cq1;
                    unsigned hits = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float2 c = rowc[j];
                        const float dot = __uint_as_float(r[j]);
                        bool pass;
                        if (METRIC == METRIC_L2)
                            pass = __fmaf_rn(-cq.x, c.y, __fmaf_rn(-2.f, dot, c.x)) < cq.y;
                        else
                            pass = __fmaf_rn(cq.x, c.y, dot) > cq.y;
                        hits |= (pass ? 1u : 0u) << j;
                    }
                    if (nvalid < 32) hits &= (1u << nvalid) - 1u;
                    if (dead) hits = 0;
                    __syncwarp();  // everyone has read rowc before the next unit rewrites it
                    if (__any_sync(0xffffffffu, hits != 0)) {
                        // one atomic per ~TC_RES survivors: lane offsets by an inclusive scan of the hit counts,
                        // inside a range of the survivor list the warp reserved (see tcfilter.cu)
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
                            const int want = max(total, 256);
                            int b2 = 0;
                            if (lane == 0) b2 = atomicAdd(&rp.ctl[CTL_NCAND], want);
                            res_pos = __shfl_sync(0xffffffffu, b2, 0);
                            res_end = res_pos + want;
                        }
                        int pos = res_pos + incl - mine;
                        res_pos += total;
                        const unsigned long long hi = (unsigned long long)(unsigned)(pair0 + g * 128 + wq * 32 + lane) << 32;
                        while (hits) {
                            const int j = __ffs(hits) - 1;
                            hits &= hits - 1;
                            if ((unsigned)pos < (unsigned)ta.cand_cap)
                                ta.cand[pos] = hi | (unsigned)(v0 + j);
                            else
                                rp.ctl[CTL_OVERFLOW] = -(1 << 30);  // survivor list full: the whole round is redone
                            pos++;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mb_arrive(&t_empty[buf]);
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&m_empty[m]);
        }
        for (int t2 = res_pos + lane; t2 < res_end; t2 += 32)  // unused tail of the last reservation: holes
            if ((unsigned)t2 < (unsigned)ta.cand_cap) ta.cand[t2] = ~0ull;
    }
#ifdef TC2_TIMING
    if (blockIdx.x == 0 && (warp <= 2 || warp == 6 || warp == T2_SCHED_WARP) && (lane == 0 || warp == 1)) {
        bool me = lane == 0;
        if (warp == 1) me = t2_wait[4] + t2_wait[11] + t2_wait[5] + t2_wait[6] > 0;  // the elected lane
        if (me)
            printf("tc_filter2 warp %d: total %lld | sched m_empty %lld | tma m_full %lld a_empty(q) %lld a_empty(list) %lld | mma m_full %lld q_full %lld t_empty %lld a_full %lld | load m_full %lld a_full %lld q_free %lld | epi m_full %lld t_full %lld\n",
                   warp, clock64() - t2_begin, t2_wait[0], t2_wait[1], t2_wait[2], t2_wait[3], t2_wait[4], t2_wait[11], t2_wait[5],
                   t2_wait[6], t2_wait[7], t2_wait[8], t2_wait[12], t2_wait[9], t2_wait[10]);
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// queries per tile of the TMEM-resident layout; 0 = this dimension needs tcfilter.cu
int tc2_tile_queries(int dpad) {
    const int kc = (dpad + 31) / 32 * 32;
    static const int g_env = getenv("AUNCEL_TC2_G") ? atoi(getenv("AUNCEL_TC2_G")) : 0;  // experiment: 1 = one query group per tile
    if (g_env == 1) return kc <= 256 ? 128 : 0;
    return kc <= 128 ? 256 : kc <= 256 ? 128 : 0;
}

void launch_tc_filter2(const RoundParams& rp, const TcArgs& ta, const void* amap64, const void* qmap64, int num_sms,
                       cudaStream_t s) {
    const int N = tc2_tile_queries(rp.dpad);
    AUNCEL_CHECK(N > 0 && ta.N == N, "tc_filter2: unsupported dimension");
    const int G = N / 128, NB = G == 2 ? 64 : 128;
    auto kern = rp.metric == METRIC_L2 ? tc_filter2_kernel<METRIC_L2> : tc_filter2_kernel<METRIC_IP>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM));
    kern<<<num_sms, T2_THREADS, T2_SMEM, s>>>(rp, ta, *reinterpret_cast<const CUtensorMap*>(amap64),
                                              *reinterpret_cast<const CUtensorMap*>(qmap64), G, NB);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
