// Tensor-core candidate filter for the bulk rounds of the scan (subsystem (2), large batches).
//
// When every query of a round already holds K results, almost no vector of a newly probed
// list can enter its top-K: only those with distance < tau (IndexIVFFlat.cpp:129).  Instead of
// evaluating all (query, vector) distances with the exact 3-op FP32 sequence, this kernel
//   1. computes dot(q, v) for a whole (256-query x 128-vector) tile with tcgen05.mma
//      kind::tf32 (operands are the same fp32 rows, staged by TMA with 128B swizzle, read by
//      the tensor core with the low 13 mantissa bits ignored), accumulating in TMEM;
//   2. turns it into a LOWER bound of the reference distance,
//        ||q||^2 + ||v||^2 - 2 dot  -  E(q, v),   E = c1 ||q|| ||v|| + c2 (||q||^2+||v||^2) + c3 |tau|
//      (c1 covers two tf32 truncations per product, 2 * 2^-10 relative, plus the accumulation;
//      see DESIGN.md section 4) and keeps the pair only if the bound is below tau;
//   3. emits the survivors (a few per query per round) to a global list; rerank_kernel
//      recomputes them with the reference's exact arithmetic (exact.cuh) and applies the
//      reference's strict test.  Results are therefore bit-identical to the SIMT path; the
//      tensor cores only decide what is worth computing exactly.
// Warp roles (352 threads): warp 0 TMA producer, warp 1 MMA issuer (one elected lane),
// warps 2-9 epilogue (TMEM -> registers -> filter; warps 2-5 take the even 32-column chunks of
// each accumulator, warps 6-9 the odd ones), warp 10 tile scheduler (claims and decodes the next tile and its per-query
// constants while the current one streams).  TMEM: 2 x 256 fp32 columns, double buffered so the
// MMAs of block b+1 overlap the filter of block b.
#include <cuda.h>

#include "exact.cuh"
#include "scan.cuh"
#include "tcfilter.cuh"
#include "tc_ptx.cuh"

namespace auncel {

#ifndef TC_EPI_GROUPS
#define TC_EPI_GROUPS 3
#endif
constexpr int TC_EG = TC_EPI_GROUPS;            // epilogue groups of four warps (one warp per TMEM lane quarter)
constexpr int TC_SCHED_WARP = 2 + 4 * TC_EG;     // warps: 0 TMA, 1 MMA, 2 .. 1 + 4 EG epilogue, then the tile scheduler
constexpr int TC_THREADS = 32 * (TC_SCHED_WARP + 1);
#ifndef TC_PREFETCH
#define TC_PREFETCH 0  // experiment: row blocks requested into L2 ahead of the stage ring (measured: 0 is best, see below)
#endif
#ifndef TC_ASTAGES_CFG
#define TC_ASTAGES_CFG 5
#endif
#ifndef TC_BKB_CFG
#define TC_BKB_CFG 128
#endif
constexpr int TC_ASTAGES = TC_ASTAGES_CFG;
constexpr int TC_A_BYTES = 128 * 128;        // 128 rows x 32 f32
constexpr int TC_B_MAX = TC_BKB_CFG * 1024;  // resident query tile: nchunk x N x 128 B
constexpr int TC_NMAX = 256;
constexpr int TC_SB_STAGES = 4;                  // streamed query tiles: stage = 16 KB of list rows + 256 queries x 128 B
constexpr int TC_SB_BYTES = TC_A_BYTES + TC_NMAX * 128;
static_assert(TC_SB_STAGES * TC_SB_BYTES <= TC_B_MAX + TC_ASTAGES * TC_A_BYTES, "streamed stages must fit the same shared memory");
constexpr int TC_RES = 256;  // survivor-list entries an epilogue warp reserves per global atomic
constexpr size_t TC_SMEM = 1024 + TC_B_MAX + (size_t)TC_ASTAGES * TC_A_BYTES + 2 * (TC_NMAX * 8 + 64);

struct TileMeta {
    int flags;  // 1 = no more tiles
    int nblk;
    int Qt;
    int pair0;
    long long row0;  // first arena row of the list
    int L;           // list length
    int pad;
    float kq[TC_NMAX];  // per query column: the side of the filter test that does not depend on the row (see the scheduler)
};

#ifdef TC_TIMING
// wait-time attribution (debug builds only): cycles spent at each barrier site, printed by CTA 0
__device__ long long g_tc_wait[16];
#define MB_WAIT(site, b, par)                        \
    do {                                             \
        long long t0_ = clock64();                   \
        mb_wait(b, par);                             \
        tc_wait_acc[site] += clock64() - t0_;        \
    } while (0)
#else
#define MB_WAIT(site, b, par) mb_wait(b, par)
#endif
template <int METRIC>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_filter_kernel(RoundParams rp, TcArgs ta, const __grid_constant__ CUtensorMap amap,
                 const __grid_constant__ CUtensorMap bmap) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long a_full[TC_ASTAGES], a_empty[TC_ASTAGES], b_full, b_empty, t_full[2],
        t_empty[2], m_full[2], m_empty[2];
    __shared__ unsigned tmem_base_s;
    unsigned char* smem = smem_dyn + ((1024u - (s32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* Bsm = smem;
    unsigned char* Asm = smem + TC_B_MAX;
    TileMeta* meta = reinterpret_cast<TileMeta*>(Asm + (size_t)TC_ASTAGES * TC_A_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef TC_TIMING
    long long tc_wait_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long tc_t_begin = clock64();
#endif
    const int N = ta.N, dpad = rp.dpad;
    const int nchunk = (dpad + 31) / 32;

    if (tid == 0) {
        for (int s = 0; s < TC_ASTAGES; s++) {
            mb_init(&a_full[s], 1);
            mb_init(&a_empty[s], 1);
        }
        mb_init(&b_full, 1);
        mb_init(&b_empty, 1);
        for (int i = 0; i < 2; i++) {
            mb_init(&t_full[i], 1);
            mb_init(&t_empty[i], 4 * TC_EG);
            mb_init(&m_full[i], 1);
            mb_init(&m_empty[i], 2 + 4 * TC_EG);  // TMA warp + MMA warp + the epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // TMEM: 512 columns (2 accumulators of up to 256 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;

    if (warp == TC_SCHED_WARP) {
        // =========================== tile scheduler ===========================
        const int total_tiles = rp.ctl[CTL_TOTAL_TILES];
        for (unsigned t = 0;; t++) {
            int T = 0;
            if (lane == 0) T = atomicAdd(&rp.ctl[CTL_TILE_COUNTER], 1);
            T = __shfl_sync(0xffffffffu, T, 0);
            const int m = t & 1;
            TileMeta* mt = &meta[m];
            int l = 0, cnt_l = 0, qt = 0, L = 0, Qt = 0, pair0 = 0, nblk = 0;
            long long L0 = 0;
            float cq[TC_NMAX / 32];
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
                Qt = min(N, cnt_l - qt * N);
                pair0 = rp.list_pair_off[l] + qt * N;
                nblk = (L + 127) / 128;
                // Filter test of a (row, query) pair, rows' part on the right:
                //   L2:  ||v||^2 (1-c2) - 2 dot - c1 |q||v| < tau + c3|tau| - ||q||^2 (1-c2)
                //        <=  dot + kq > 0.5 ||v||^2 (1-c2),   kq = 0.5 (rhs + c1 |q| max|v|) + slack
                //   IP:  dot + c1/2 |q||v| > tau - c3|tau|   <=  dot > kq,   kq = tau - c3|tau| - c1/2 |q| max|v| - slack
                // with max|v| over the LIST instead of the row's own norm (only the margin grows, nothing the exact
                // test accepts is lost) and a slack for the roundings of this evaluation: one add and one compare
                // per pair instead of two FMAs and a compare.
                const float nmax = ta.list_nmax[l], snmax = sqrtf(nmax);
#pragma unroll
                for (int jj = 0; jj < TC_NMAX / 32; jj++) {
                    const int j = jj * 32 + lane;
                    float c = METRIC == METRIC_L2 ? -FLT_MAX : FLT_MAX;  // never passes
                    if (j < Qt) {
                        unsigned long long pr = rp.pairs[pair0 + j];
                        int q = rp.active[(int)(pr >> 32)];
                        float tau = rp.st.tau[q], nq = ta.qnorm[q];
                        if (tau == (METRIC == METRIC_L2 ? FLT_MAX : -FLT_MAX)) {
                            // The query's heap is not full yet (its first lists held fewer than K vectors): no
                            // threshold, every vector of the list would survive, overflow the slot and be redone
                            // exactly anyway -- 73 such queries of the bench batch produced 5.5 M of the first filter
                            // round's 5.6 M survivors.  Hand the pair to the exact redo right away: the column never
                            // passes, the slot is flagged, the host sees a non-zero overflow count.
                            rp.pair_flag[(long)(pr >> 32) * rp.w + (long)(pr & 0xffffffffu)] = 1;
                            atomicAdd(&rp.ctl[CTL_OVERFLOW], 1);
                        } else if (METRIC == METRIC_L2) {
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
            MB_WAIT(0, &m_empty[m], ((t >> 1) & 1) ^ 1);
            if (T >= total_tiles) {
                if (lane == 0) {
                    mt->flags = 1;
                    mb_arrive(&m_full[m]);
                }
                break;
            }
#pragma unroll
            for (int jj = 0; jj < TC_NMAX / 32; jj++)
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
            MB_WAIT(1, &m_full[m], (t >> 1) & 1);
            const int flags = meta[m].flags, nblk = meta[m].nblk, pair0 = meta[m].pair0;
            const long long L0 = meta[m].row0;
            __syncwarp();
            if (lane == 0) mb_arrive(&m_empty[m]);
            if (flags) break;
            if (lane == 0 && ta.stream_b) {
                // d > 256: a k-chunk of the query tile (256 x 32 f32) rides in every stage next to the list rows'
                // k-chunk; the tile is re-read (from L2) for every 128-row block, but a list pass serves 256 queries
                // instead of the 32 that fit resident at d = 960
                for (int blk = 0; blk < nblk; blk++)
                    for (int c = 0; c < nchunk; c++, ita++) {
                        const int s = ita % TC_SB_STAGES;
                        MB_WAIT(3, &a_empty[s], ((ita / TC_SB_STAGES) & 1) ^ 1);
                        mb_expect_tx(&a_full[s], (unsigned)(TC_A_BYTES + N * 128));
                        tma2d(smem + (size_t)s * TC_SB_BYTES, &amap, c * 32, (int)(L0 + (long long)blk * 128), &a_full[s]);
                        tma2d(smem + (size_t)s * TC_SB_BYTES + TC_A_BYTES, &bmap, c * 32, pair0, &a_full[s]);
                    }
            } else if (lane == 0) {
                // queries: resident for the whole tile, one swizzled [N x 32] block per k-chunk
                MB_WAIT(2, &b_empty, (t & 1) ^ 1);
                mb_expect_tx(&b_full, (unsigned)(nchunk * N * 128));
                for (int c = 0; c < nchunk; c++) tma2d(Bsm + (size_t)c * N * 128, &bmap, c * 32, pair0, &b_full);
#if TC_PREFETCH > 0
                // row blocks [0, TC_PREFETCH) of the list go to L2 now, block blk + TC_PREFETCH when block blk is staged
                for (int pb = 0; pb < min(nblk, TC_PREFETCH); pb++)
                    for (int c = 0; c < nchunk; c++) tma2d_prefetch(&amap, c * 32, (int)(L0 + (long long)pb * 128));
#endif
                for (int blk = 0; blk < nblk; blk++)
                    for (int c = 0; c < nchunk; c++, ita++) {
#if TC_PREFETCH > 0
                        if (blk + TC_PREFETCH < nblk) tma2d_prefetch(&amap, c * 32, (int)(L0 + (long long)(blk + TC_PREFETCH) * 128));
#endif
                        const int s = ita % TC_ASTAGES;
                        MB_WAIT(3, &a_empty[s], ((ita / TC_ASTAGES) & 1) ^ 1);
#ifdef TC_EXP_SKIP_A  // experiment: only every TC_EXP_SKIP_A-th stage is really loaded (results are garbage)
                        if (ita % TC_EXP_SKIP_A != 0) {
                            mb_arrive(&a_full[s]);
                            continue;
                        }
#endif
                        mb_expect_tx(&a_full[s], TC_A_BYTES);
                        tma2d(Asm + (size_t)s * TC_A_BYTES, &amap, c * 32, (int)(L0 + (long long)blk * 128), &a_full[s]);
                    }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        // One lane runs the whole loop: this warp's instruction latency is what feeds the tensor
        // pipe (four MMAs of ~180 cycles per 16 KB stage), so the per-stage instruction count is
        // kept minimal -- descriptors are base + increments, the stage ring is a counter.
        // idesc: D = f32, A = B = tf32, both K-major, N >> 3, M = 128 >> 4
        if (elect_one()) {
            const unsigned idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
            const unsigned long long adesc0 = umma_desc(s32(Asm)), bdesc0 = umma_desc(s32(Bsm));
            const unsigned b_chunk_inc = (unsigned)(N * 128) >> 4;  // descriptor address units are 16 bytes
            unsigned s = 0, a_phase = 0, blkc = 0;
            for (unsigned t = 0;; t++) {
                const int m = t & 1;
                MB_WAIT(4, &m_full[m], (t >> 1) & 1);
                const int flags = meta[m].flags, nblk = meta[m].nblk;
                const int Nt = min(N, (meta[m].Qt + 31) / 32 * 32);  // MMA N: only the columns that hold queries
                const unsigned idesc = idesc0 | ((unsigned)(Nt >> 3) << 17);
                mb_arrive(&m_empty[m]);
                if (flags) break;
                if (ta.stream_b) {
                    const unsigned long long sdesc0 = umma_desc(s32(smem));
                    for (int blk = 0; blk < nblk; blk++, blkc++) {
                        const int buf = blkc & 1;
                        MB_WAIT(6, &t_empty[buf], ((blkc >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const unsigned d_tmem = tmem_base + buf * 256;
                        for (int c = 0; c < nchunk; c++) {
                            MB_WAIT(7, &a_full[s], a_phase);
                            tc_fence_after();
                            const unsigned long long adesc = sdesc0 + s * (unsigned)(TC_SB_BYTES >> 4);
                            const unsigned long long bdesc = adesc + (unsigned)(TC_A_BYTES >> 4);
#pragma unroll
                            for (int k = 0; k < 4; k++) umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (c | k) != 0);
                            umma_commit(&a_empty[s]);
                            if (++s == TC_SB_STAGES) {
                                s = 0;
                                a_phase ^= 1;
                            }
                        }
                        umma_commit(&t_full[buf]);
                    }
                    continue;
                }
                MB_WAIT(5, &b_full, t & 1);
                for (int blk = 0; blk < nblk; blk++, blkc++) {
                    const int buf = blkc & 1;
                    MB_WAIT(6, &t_empty[buf], ((blkc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const unsigned d_tmem = tmem_base + buf * 256;
                    unsigned long long bdesc = bdesc0;
                    for (int c = 0; c < nchunk; c++, bdesc += b_chunk_inc) {
                        MB_WAIT(7, &a_full[s], a_phase);
                        tc_fence_after();
                        const unsigned long long adesc = adesc0 + s * (unsigned)(TC_A_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 4; k++)  // K = 8 tf32 = 32 bytes (2 address units) per instruction
                            umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (c | k) != 0);
                        umma_commit(&a_empty[s]);  // frees the A stage when these MMAs have read it
                        if (++s == TC_ASTAGES) {
                            s = 0;
                            a_phase ^= 1;
                        }
                    }
                    umma_commit(&t_full[buf]);
                }
                umma_commit(&b_empty);  // the tile's MMAs are done with the query block
            }
        }
        __syncwarp();
    } else if (warp >= 2 && warp < TC_SCHED_WARP) {
        // =========================== epilogue ===========================
        // TC_EG groups of four warps; all drain every accumulator, group g taking the 32-column
        // chunks g, g + TC_EG, ...  The filter of a block must finish within the MMA time of the next
        // one or the MMA warp stalls on t_empty; twelve warps give every scheduler three epilogue
        // warps to hide TMEM / shared-memory latencies (measured: 2 groups 2.07 ms for the 2-tile round,
        // 3 groups 1.98, 4 groups 2.00; two query columns per 128-bit constant load: no gain).
        const int wq = warp & 3;  // TMEM lane quarter this warp may read
        const int grp = (warp - 2) >> 2;
        unsigned blkc = 0;
        int res_pos = 0, res_end = 0;  // this warp's reserved range of the survivor list
        for (unsigned t = 0;; t++) {
            const int m = t & 1;
            MB_WAIT(8, &m_full[m], (t >> 1) & 1);
            const TileMeta* mt = &meta[m];
            if (mt->flags) {
                __syncwarp();
                if (lane == 0) mb_arrive(&m_empty[m]);
                break;
            }
            const int nblk = mt->nblk, L = mt->L, pair0 = mt->pair0;
            // once the survivor list is full the round is redone anyway: stop counting (checked once per
            // tile), so the signed counter can never wrap however many pairs pass a loose threshold
            const bool dead = *reinterpret_cast<volatile int*>(&rp.ctl[CTL_OVERFLOW]) < 0;
            const int ncg = min(N, (mt->Qt + 31) / 32 * 32) / 32;
            const long long row0 = mt->row0;
            float nv_next = (nblk > 0 && wq * 32 + lane < L) ? ta.vnorm[row0 + wq * 32 + lane] : 0.f;
            for (int blk = 0; blk < nblk; blk++, blkc++) {
                const int buf = blkc & 1;
                const int v = blk * 128 + wq * 32 + lane;
                const bool valid = v < L;
                const float nv = nv_next;  // (fetched one block ahead: the load's latency stays off the t_full -> t_empty path)
                {
                    const int vn = v + 128;
                    nv_next = (blk + 1 < nblk && vn < L) ? ta.vnorm[row0 + vn] : 0.f;
                }
                const float nvh = METRIC == METRIC_L2 ? 0.5f * (nv * (1.f - ta.c2)) : 0.f;  // the row's side of the test
                MB_WAIT(9, &t_full[buf], (blkc >> 1) & 1);
                tc_fence_after();
                for (int cg = grp; cg < ncg; cg += TC_EG) {
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
#ifdef TC_EXP_NOAPPEND  // experiment: the filter pipeline without the survivor appends
                    hits = 0;
#endif
                    if (__any_sync(0xffffffffu, hits != 0)) {
                        // one atomic per warp and chunk: lane offsets by an inclusive scan of the hit counts
                        const int mine = __popc(hits);
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int up = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += up;
                        }
                        // the warp owns a reserved range [res_pos, res_end) of the survivor list and refills it
                        // RES entries at a time: one global atomic per ~RES survivors instead of one per chunk
                        const int total = __shfl_sync(0xffffffffu, incl, 31);
                        if (res_pos + total > res_end) {
                            // hand the unused tail back as holes, then take a fresh range
                            for (int t2 = res_pos + lane; t2 < res_end; t2 += 32)
                                if ((unsigned)t2 < (unsigned)ta.cand_cap) ta.cand[t2] = ~0ull;
                            const int want = max(total, TC_RES);
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
#ifdef TC_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp <= 2 || warp == 6 || warp == TC_SCHED_WARP)) {
        const long long tot = clock64() - tc_t_begin;
        printf("tc_filter warp %d: total %lld | m_empty %lld m_full %lld/%lld/%lld b_empty %lld a_empty %lld b_full %lld t_empty %lld a_full %lld t_full %lld\n",
               warp, tot, tc_wait_acc[0], tc_wait_acc[1], tc_wait_acc[4], tc_wait_acc[8], tc_wait_acc[2], tc_wait_acc[3],
               tc_wait_acc[5], tc_wait_acc[6], tc_wait_acc[7], tc_wait_acc[9]);
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// Survivors of the filter, recomputed with the reference's arithmetic and test.
#ifndef RR_UNROLL
#define RR_UNROLL 4
#endif
#ifndef RR_BLOCKS
#define RR_BLOCKS 16
#endif
constexpr int RR_UNROLL_N = RR_UNROLL;  // row steps (float4 pairs) in flight per thread
template <int METRIC>
__global__ void rerank_kernel(RoundParams rp, TcArgs ta) {
    const int ncand = (int)min((unsigned)rp.ctl[CTL_NCAND], (unsigned)ta.cand_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncand; i += gridDim.x * blockDim.x) {
        const unsigned long long c = ta.cand[i];
        if (c == ~0ull) continue;  // hole: unused part of a warp's reservation
        const int pos = (int)(c >> 32);
        const unsigned v = (unsigned)(c & 0xffffffffu);
        const unsigned long long pr = rp.pairs[pos];
        const int a = (int)(pr >> 32), p_rel = (int)(pr & 0xffffffffu);
        const int q = rp.active[a];
        const int l = rp.ckeys[(long long)q * rp.nlist + rp.r0 + p_rel];
        const float4* x = reinterpret_cast<const float4*>(rp.xq_sorted + (long long)pos * rp.dpad);
        const float4* y = reinterpret_cast<const float4*>(rp.codes + (rp.list_off[l] + v) * (long long)rp.dpad);
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        // RR_UNROLL_N row steps are fetched before the first is used.  Measured (largest launch of the bench
        // step, 5.6 M survivors): 1 / 2 steps in flight 1.27 / 1.29 ms, 4 steps 0.97, 8 steps 1.11 (registers);
        // beyond that the kernel is bound by DRAM's rate for scattered 512-byte rows (2.2 GB in 0.97 ms, L2 hit 41 %).
        // A variant that fetched 32 rows per warp cooperatively into shared memory (coalesced 128-byte
        // pieces) was slower (+0.75 ms per step): the L1 wavefronts are not what limits it.
        const int n4 = rp.dpad / 4;
        for (int k0 = 0; k0 < n4; k0 += RR_UNROLL_N) {
            float4 xx[RR_UNROLL_N], yy[RR_UNROLL_N];
#pragma unroll
            for (int u = 0; u < RR_UNROLL_N; u++)
                if (k0 + u < n4) {
                    yy[u] = __ldg(y + k0 + u);
                    xx[u] = __ldg(x + k0 + u);
                }
#pragma unroll
            for (int u = 0; u < RR_UNROLL_N; u++)
                if (k0 + u < n4) exact_step<METRIC>(s, xx[u], yy[u]);
        }
        const float dist = exact_finish(s);
        const float tau = rp.st.tau[q];
        if (METRIC == METRIC_L2 ? dist < tau : dist > tau) {
            const long slot = (long)a * rp.w + p_rel;  // S == nsub == 1
            const int o = atomicAdd(&rp.slot_cnt[slot], 1);
            if (o < rp.cap) {
                rp.cand_d[slot * rp.cap + o] = dist;
                rp.cand_off[slot * rp.cap + o] = v;
            } else {
                rp.pair_flag[slot] = 1;  // more vectors of this list beat tau than a slot holds: exact scan redoes the pair
                atomicAdd(&rp.ctl[CTL_OVERFLOW], 1);
            }
        }
    }
}

// Audit (option "tc_audit", tests only): the exact scan has redone the whole round into a second
// pool; every slot the filter path did not flag as overflowed must hold exactly the same set of
// (distance, offset) candidates -- i.e. no pair the reference's strict test accepts was dropped.
__global__ void tc_audit_kernel(RoundParams tc, const float* __restrict__ ex_d, const unsigned* __restrict__ ex_off,
                                const int* __restrict__ ex_cnt, unsigned long long* __restrict__ ctr) {
    const long slot = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slot >= (long)tc.n_active * tc.w) return;
    if (tc.pair_flag[slot]) return;  // more than K survivors: the exact scan rewrites this slot anyway
    // the exact rescan keeps the K best candidates below tau, the filter path everything below tau (up to
    // the slot capacity): every exact entry must be present, and the counts must agree up to K
    const int K = tc.K, cap = tc.cap;
    const int c1 = tc.slot_cnt[slot] & ~SLOT_SORTED, c2 = ex_cnt[slot] & ~SLOT_SORTED;
    bool bad = min(c1, K) != c2;
    if (!bad)
        for (int i = lane; i < c2; i += 32) {
            const float dv = ex_d[slot * K + i];
            const unsigned ov = ex_off[slot * K + i];
            bool found = false;
            for (int j = 0; j < c1; j++)
                found |= (tc.cand_off[slot * cap + j] == ov) &&
                         (__float_as_uint(tc.cand_d[slot * cap + j]) == __float_as_uint(dv));
            bad |= !found;
        }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (bad) atomicAdd(ctr, 1ull);
        atomicAdd(ctr + 1, 1ull);
        atomicAdd(ctr + 2, (unsigned long long)c2);
    }
}

void launch_tc_audit(const RoundParams& tc, const float* ex_d, const unsigned* ex_off, const int* ex_cnt,
                     unsigned long long* ctr, cudaStream_t s) {
    const long slots = (long)tc.n_active * tc.w;
    if (slots == 0) return;
    tc_audit_kernel<<<(unsigned)((slots * 32 + 255) / 256), 256, 0, s>>>(tc, ex_d, ex_off, ex_cnt, ctr);
    CUDA_CHECK(cudaGetLastError());
}

// ||v||^2 of every arena row (double accumulation, rounded to float)
__global__ void row_norms_kernel(const float* __restrict__ x, long long n, int dpad, float* __restrict__ out) {
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= n) return;
    double s = 0.0;
    for (int c = lane; c < dpad; c += 32) {
        double v = x[r * dpad + c];
        s += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = (float)s;
}

void launch_row_norms(const float* x, long long n, int dpad, float* out, cudaStream_t s) {
    if (n == 0) return;
    row_norms_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(x, n, dpad, out);
    CUDA_CHECK(cudaGetLastError());
}

// max ||v||^2 of every inverted list (one warp per list)
__global__ void list_norm_max_kernel(const float* __restrict__ vnorm, const long long* __restrict__ list_off, long nlist,
                                     float* __restrict__ out) {
    const long l = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (l >= nlist) return;
    float m = 0.f;
    for (long long i = list_off[l] + lane; i < list_off[l + 1]; i += 32) m = fmaxf(m, vnorm[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[l] = m;
}

void launch_list_norm_max(const float* vnorm, const long long* list_off, long nlist, float* out, cudaStream_t s) {
    if (nlist == 0) return;
    list_norm_max_kernel<<<(unsigned)((nlist * 32 + 255) / 256), 256, 0, s>>>(vnorm, list_off, nlist, out);
    CUDA_CHECK(cudaGetLastError());
}

static int tc_resident_queries(int dpad) {
    int nchunk = (dpad + 31) / 32;
    int n = TC_B_MAX / (nchunk * 128);
    return std::min(n, TC_NMAX) / 32 * 32;  // the epilogue reads TMEM 32 columns at a time
}

// fewer than 128 queries fit resident (d > 256): stream the query tile instead
bool tc_stream_queries(int dpad) {
    static const int off = getenv("AUNCEL_TC_NO_STREAM") ? 1 : 0;  // experiment: resident tiles at every dimension
    return !off && tc_resident_queries(dpad) < 128;
}

int tc_tile_queries(int dpad, bool streamed) { return streamed ? TC_NMAX : tc_resident_queries(dpad); }

void launch_tc_filter(const RoundParams& rp, const TcArgs& ta, const void* amap, const void* bmap, int num_sms,
                      cudaStream_t s) {
    auto kern = rp.metric == METRIC_L2 ? tc_filter_kernel<METRIC_L2> : tc_filter_kernel<METRIC_IP>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    kern<<<num_sms, TC_THREADS, TC_SMEM, s>>>(rp, ta, *reinterpret_cast<const CUtensorMap*>(amap),
                                             *reinterpret_cast<const CUtensorMap*>(bmap));
    CUDA_CHECK(cudaGetLastError());
}

void launch_rerank(const RoundParams& rp, const TcArgs& ta, int num_sms, cudaStream_t s) {
    auto kern = rp.metric == METRIC_L2 ? rerank_kernel<METRIC_L2> : rerank_kernel<METRIC_IP>;
    kern<<<num_sms * RR_BLOCKS, 128, 0, s>>>(rp, ta);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
