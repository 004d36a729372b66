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
#include "scan.cuh"
#include "exact.cuh"

namespace auncel {

// ------------------------------------------------------------------------- planning
__global__ void plan_count_kernel(RoundParams rp) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long tot = (long)rp.n_active * rp.w;
    if (idx >= tot) return;
    int a = (int)(idx / rp.w), p_rel = (int)(idx - (long)a * rp.w);
    int q = rp.active[a];
    int p = rp.r0 + p_rel;
    if (p >= rp.st.bound[q]) return;
    int l = rp.ckeys[(long)q * rp.nlist + p];
    if (rp.list_off[l + 1] == rp.list_off[l]) return;  // IndexIVF.cpp:452-455
    atomicAdd(&rp.list_cnt[l], 1);
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
        int t = ((c + SCAN_QT - 1) / SCAN_QT) * rp.S;
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
    int pos = rp.list_pair_off[l] + atomicAdd(&rp.list_cursor[l], 1);
    rp.pairs[pos] = ((unsigned long long)(unsigned)a << 32) | (unsigned)p_rel;
}

void launch_plan(const RoundParams& rp, cudaStream_t s) {
    long tot = (long)rp.n_active * rp.w;
    CUDA_CHECK(cudaMemsetAsync(rp.list_cnt, 0, rp.nlist * sizeof(int), s));
    CUDA_CHECK(cudaMemsetAsync(rp.slot_cnt, 0, (size_t)tot * rp.S * sizeof(int), s));
    unsigned blocks = (unsigned)((tot + 255) / 256);
    plan_count_kernel<<<blocks, 256, 0, s>>>(rp);
    plan_offsets_kernel<<<1, 1024, 0, s>>>(rp);
    plan_fill_kernel<<<blocks, 256, 0, s>>>(rp);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------- scan
constexpr int STAGES = 4;
constexpr int LD = SCAN_DK + 4;                 // padded smem row, floats
constexpr int CAP = 256;                        // candidate buffer per query (>= MAX_K + 32)
constexpr int STAGE_FLOATS = (SCAN_VT + SCAN_QT) * LD;
constexpr size_t SCAN_SMEM = (size_t)STAGES * STAGE_FLOATS * 4 + (size_t)SCAN_QT * CAP * 8 + 1024;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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

// sort the buffer, keep the K best, tighten tau when K are held
template <int METRIC>
__device__ __forceinline__ void compact(unsigned long long* buf, int& cnt, float& tau, int K, int lane) {
    for (int i = cnt + lane; i < CAP; i += 32) buf[i] = ~0ull;
    __syncwarp();
    warp_sort_smem<CAP>(buf, lane);
    if (cnt > K) cnt = K;
    if (cnt == K) tau = key_dist<METRIC>(buf[K - 1]);
    __syncwarp();
}

template <int METRIC>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(RoundParams rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* stage_base = reinterpret_cast<float*>(smem_raw);
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)STAGES * STAGE_FLOATS * 4);
    __shared__ int s_tile;
    __shared__ int s_q[SCAN_QT];        // query index of each tile row (-1: none)
    __shared__ int s_slot[SCAN_QT];     // pool slot of each tile row
    __shared__ float s_tau[SCAN_QT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = rp.K, dpad = rp.dpad;
    const int nchunk = (dpad + SCAN_DK - 1) / SCAN_DK;
    const int total_tiles = rp.ctl[CTL_TOTAL_TILES];

    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(&rp.ctl[CTL_TILE_COUNTER], 1);
        __syncthreads();
        const int T = s_tile;
        if (T >= total_tiles) break;

        // ---- decode tile -> (list, query tile, segment)
        int lo = 0, hi = (int)rp.nlist;  // last l with list_tile_off[l] <= T
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (rp.list_tile_off[mid] <= T) lo = mid; else hi = mid;
        }
        const int l = lo;
        const int cnt_l = rp.list_pair_off[l + 1] - rp.list_pair_off[l];
        const int nqt = (cnt_l + SCAN_QT - 1) / SCAN_QT;
        const int tl = T - rp.list_tile_off[l];
        const int seg = tl / nqt, qt = tl - seg * nqt;
        const long long L0 = rp.list_off[l];
        const int L = (int)(rp.list_off[l + 1] - L0);
        int seg_len = (L + rp.S - 1) / rp.S;
        seg_len = (seg_len + 31) / 32 * 32;
        const int v_begin = seg * seg_len;
        const int v_end = min(L, v_begin + seg_len);
        const int Qt = min(SCAN_QT, cnt_l - qt * SCAN_QT);
        if (v_begin >= v_end) continue;  // empty segment: slot_cnt stays 0

        if (tid < SCAN_QT) {
            int q = -1, slot = 0;
            float tau = 0.f;
            if (tid < Qt) {
                unsigned long long pr = rp.pairs[rp.list_pair_off[l] + qt * SCAN_QT + tid];
                int a = (int)(pr >> 32), p_rel = (int)(pr & 0xffffffffu);
                q = rp.active[a];
                slot = (a * rp.w + p_rel) * rp.S + seg;
                tau = rp.st.tau[q];
            }
            s_q[tid] = q;
            s_slot[tid] = slot;
            s_tau[tid] = tau;
        }
        __syncthreads();

        const int nblk = (v_end - v_begin + SCAN_VT - 1) / SCAN_VT;
        const int total_it = nblk * nchunk;
        const float* lbase = rp.codes + (L0 + v_begin) * (long long)dpad;
        const int nvec = v_end - v_begin;

        auto issue = [&](int it) {
            if (it < total_it) {
                int blk = it / nchunk, c = it - blk * nchunk;
                float* sv = stage_base + (size_t)(it % STAGES) * STAGE_FLOATS;
                float* sq = sv + SCAN_VT * LD;
                int k0 = c * SCAN_DK;
                // vectors: 128 rows x 8 x 16B
#pragma unroll
                for (int t = 0; t < (SCAN_VT * 8) / SCAN_THREADS; t++) {
                    int idx = tid + t * SCAN_THREADS;
                    int r = idx >> 3, cc = (idx & 7) * 4;
                    int v = blk * SCAN_VT + r;
                    bool ok = v < nvec && k0 + cc < dpad;
                    const float* src = ok ? lbase + (long long)v * dpad + k0 + cc : rp.codes;
                    cp_async16(sv + r * LD + cc, src, ok ? 16 : 0);
                }
                // queries: 32 rows x 8 x 16B
                {
                    int r = tid >> 3, cc = (tid & 7) * 4;
                    int q = s_q[r];
                    bool ok = q >= 0 && k0 + cc < dpad;
                    const float* src = ok ? rp.xq + (long long)q * dpad + k0 + cc : rp.xq;
                    cp_async16(sq + r * LD + cc, src, ok ? 16 : 0);
                }
            }
            cp_async_commit();
        };

        // per-warp selection state for its 4 queries
        const bool warp_has_q = warp * 4 < Qt;
        int cnt[4] = {0, 0, 0, 0};
        float tau[4];
#pragma unroll
        for (int i = 0; i < 4; i++) tau[i] = s_tau[warp * 4 + i];
        float acc[4][4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int x = 0; x < 4; x++) acc[i][j][x] = 0.f;

#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) issue(s);

        for (int it = 0; it < total_it; it++) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            issue(it + STAGES - 1);
            const int blk = it / nchunk, c = it - blk * nchunk;
            if (warp_has_q) {
                const float* sv = stage_base + (size_t)(it % STAGES) * STAGE_FLOATS;
                const float* sq = sv + SCAN_VT * LD + warp * 4 * LD;
#pragma unroll
                for (int kk = 0; kk < SCAN_DK; kk += 4) {
                    float4 a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const float4*>(sq + i * LD + kk);
#pragma unroll
                    for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const float4*>(sv + (lane + 32 * j) * LD + kk);
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) exact_step<METRIC>(acc[i][j], a[i], b[j]);
                }
                if (c == nchunk - 1) {
                    // ---- epilogue: filter against tau, append, compact when needed
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        unsigned long long* buf = cand + (size_t)(warp * 4 + i) * CAP;
                        const bool qok = warp * 4 + i < Qt;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float dist = exact_finish(acc[i][j]);
                            acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
                            int v = blk * SCAN_VT + lane + 32 * j;
                            bool pass = qok && v < nvec &&
                                        (METRIC == METRIC_L2 ? dist < tau[i] : dist > tau[i]);
                            unsigned m = __ballot_sync(0xffffffffu, pass);
                            if (m) {
                                if (cnt[i] + 32 > CAP) compact<METRIC>(buf, cnt[i], tau[i], K, lane);
                                // tau may have tightened: re-test
                                pass = pass && (METRIC == METRIC_L2 ? dist < tau[i] : dist > tau[i]);
                                m = __ballot_sync(0xffffffffu, pass);
                                if (pass) {
                                    int pos = cnt[i] + __popc(m & ((1u << lane) - 1));
                                    buf[pos] = make_key<METRIC>(dist, (unsigned)(v_begin + v));
                                }
                                cnt[i] += __popc(m);
                                __syncwarp();
                            }
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();

        // ---- write the per-(query, segment) candidates: sorted, at most K
        if (warp_has_q) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (warp * 4 + i >= Qt) continue;
                unsigned long long* buf = cand + (size_t)(warp * 4 + i) * CAP;
                compact<METRIC>(buf, cnt[i], tau[i], K, lane);
                const long slot = s_slot[warp * 4 + i];
                for (int t = lane; t < cnt[i]; t += 32) {
                    unsigned long long key = buf[t];
                    rp.cand_d[slot * K + t] = key_dist<METRIC>(key);
                    rp.cand_off[slot * K + t] = (unsigned)(key & 0xffffffffu);
                }
                if (lane == 0) rp.slot_cnt[slot] = cnt[i];
                __syncwarp();
            }
        }
    }
}

void launch_scan(const RoundParams& rp, int num_sms, cudaStream_t s) {
    auto kern = rp.metric == METRIC_L2 ? scan_kernel<METRIC_L2> : scan_kernel<METRIC_IP>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM));
    kern<<<num_sms, SCAN_THREADS, SCAN_SMEM, s>>>(rp);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
