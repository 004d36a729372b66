// Subsystem (1): coarse assignment.
//   dense_exact_kernel : query x centroid distances with the reference's exact arithmetic
//                        (fvec_L2sqr / fvec_inner_product SSE order, utils_simd.cpp:391-443 --
//                        the knn_*_sse path of utils.cpp:417-490 that IndexFlat::search takes
//                        for nx < 20 and that the parity harness forces for batches)
//   rank_rows_kernel   : full ranking of all nlist centroids per query (heap_reorder output of
//                        IndexFlat::search with k = nprobe = nlist, profile.cpp:218-222)
//   the same tile kernel also fills interdis_cem (IVF_pro.cpp:21-39) and does add()'s k=1 assign.
#include <vector>

#include "engine.h"
#include "exact.cuh"

namespace auncel {

// ---------------------------------------------------------------------------------
__global__ void pad_rows_kernel(const float* __restrict__ src, long n, int d, float* __restrict__ dst,
                                int dpad) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long tot = n * (long)dpad;
    for (; i < tot; i += (long)gridDim.x * blockDim.x) {
        long r = i / dpad;
        int c = (int)(i - r * dpad);
        dst[i] = c < d ? src[r * (long)d + c] : 0.f;
    }
}

void launch_pad_rows(const float* src, long n, int d, float* dst, int dpad, cudaStream_t s) {
    if (n == 0) return;
    long tot = n * (long)dpad;
    int blocks = (int)std::min<long>((tot + 255) / 256, 148 * 16);
    pad_rows_kernel<<<blocks, 256, 0, s>>>(src, n, d, dst, dpad);
}

// ---------------------------------------------------------------------------------
// 64 x 64 tile, 256 threads, 4 x 4 pairs per thread, 4 lane-accumulators per pair.
constexpr int DT = 64;
constexpr int DKC = 32;            // floats per k chunk
constexpr int DLD = DKC + 4;       // padded smem row (conflict-free 128-bit reads)

enum { OUT_DENSE = 0, OUT_BEST = 1, OUT_TRI = 2 };

template <int METRIC, int OUT>
__global__ void __launch_bounds__(256)
dense_exact_kernel(const float* __restrict__ X, long nx, const float* __restrict__ Y, long ny, int dpad,
                   float* __restrict__ out, unsigned long long* __restrict__ best) {
    __shared__ __align__(16) float sx[DT][DLD];
    __shared__ __align__(16) float sy[DT][DLD];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const long x0 = (long)blockIdx.y * DT, y0 = (long)blockIdx.x * DT;
    if (OUT == OUT_TRI && y0 + DT <= x0) return;  // strictly-lower tiles hold no (i<j) pair

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int l = 0; l < 4; l++) acc[i][j][l] = 0.f;

    for (int k0 = 0; k0 < dpad; k0 += DKC) {
        // 64 rows x 8 float4 per operand = 512 float4 each, 2 per thread per operand
#pragma unroll
        for (int t = 0; t < 2; t++) {
            int idx = tid + t * 256;
            int r = idx >> 3, c = (idx & 7) * 4;
            float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vy = vx;
            if (x0 + r < nx && k0 + c < dpad) vx = *reinterpret_cast<const float4*>(X + (x0 + r) * dpad + k0 + c);
            if (y0 + r < ny && k0 + c < dpad) vy = *reinterpret_cast<const float4*>(Y + (y0 + r) * dpad + k0 + c);
            *reinterpret_cast<float4*>(&sx[r][c]) = vx;
            *reinterpret_cast<float4*>(&sy[r][c]) = vy;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DKC; kk += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const float4*>(&sx[ty * 4 + i][kk]);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const float4*>(&sy[tx + 16 * j][kk]);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) exact_step<METRIC>(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; i++) {
        long xi = x0 + ty * 4 + i;
        if (xi >= nx) continue;
        unsigned long long bk = ~0ull;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            long yj = y0 + tx + 16 * j;
            if (yj >= ny) continue;
            float v = exact_finish(acc[i][j]);
            if (OUT == OUT_DENSE) {
                out[xi * ny + yj] = v;
            } else if (OUT == OUT_TRI) {
                if (xi < yj) out[tri_index((size_t)nx, (size_t)xi, (size_t)yj)] = v;
            } else {
                uint32_t o = f2ord(v);
                if (METRIC == METRIC_IP) o = ~o;
                unsigned long long key = ((unsigned long long)o << 32) | (unsigned)yj;
                bk = key < bk ? key : bk;
            }
        }
        if (OUT == OUT_BEST && bk != ~0ull) atomicMin(&best[xi], bk);
    }
}

void launch_coarse_distances(int metric, const float* xq, long nq, const float* cent, long nlist,
                             int dpad, float* out_dis, unsigned long long* out_best, cudaStream_t s) {
    if (nq == 0) return;
    dim3 grid((unsigned)((nlist + DT - 1) / DT), (unsigned)((nq + DT - 1) / DT));
    AUNCEL_CHECK(grid.y <= 65535, "too many queries for one coarse launch");
    if (out_dis) {
        if (metric == METRIC_L2)
            dense_exact_kernel<METRIC_L2, OUT_DENSE><<<grid, 256, 0, s>>>(xq, nq, cent, nlist, dpad, out_dis, nullptr);
        else
            dense_exact_kernel<METRIC_IP, OUT_DENSE><<<grid, 256, 0, s>>>(xq, nq, cent, nlist, dpad, out_dis, nullptr);
    } else {
        CUDA_CHECK(cudaMemsetAsync(out_best, 0xff, nq * sizeof(unsigned long long), s));
        if (metric == METRIC_L2)
            dense_exact_kernel<METRIC_L2, OUT_BEST><<<grid, 256, 0, s>>>(xq, nq, cent, nlist, dpad, nullptr, out_best);
        else
            dense_exact_kernel<METRIC_IP, OUT_BEST><<<grid, 256, 0, s>>>(xq, nq, cent, nlist, dpad, nullptr, out_best);
    }
    CUDA_CHECK(cudaGetLastError());
}

void launch_interdis(int metric, const float* cent, long nlist, int dpad, float* out, cudaStream_t s) {
    dim3 grid((unsigned)((nlist + DT - 1) / DT), (unsigned)((nlist + DT - 1) / DT));
    if (metric == METRIC_L2)
        dense_exact_kernel<METRIC_L2, OUT_TRI><<<grid, 256, 0, s>>>(cent, nlist, cent, nlist, dpad, out, nullptr);
    else
        dense_exact_kernel<METRIC_IP, OUT_TRI><<<grid, 256, 0, s>>>(cent, nlist, cent, nlist, dpad, out, nullptr);
    CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------
// Full ranking: one CTA per query, bitonic sort of P = pow2 >= nlist 64-bit keys in smem.
// key = ord(dis) (L2, ascending) or ~ord(sim) (IP, descending) << 32 | centroid id: equal
// distances rank by centroid id.
__global__ void __launch_bounds__(1024)
rank_rows_kernel(int metric, const float* __restrict__ dis, long nlist, int P, float* __restrict__ out_dis,
                 int* __restrict__ out_keys, int* __restrict__ tie0) {
    extern __shared__ unsigned long long skey[];
    __shared__ int s_tie;
    if (threadIdx.x == 0) s_tie = 0x7fffffff;
    const long q = blockIdx.x;
    const float* row = dis + q * nlist;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < nlist) {
            uint32_t o = f2ord(row[i]);
            if (metric == METRIC_IP) o = ~o;
            key = ((unsigned long long)o << 32) | (unsigned)i;
        }
        skey[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                unsigned long long a = skey[lo], b = skey[hi];
                if ((a > b) == up) {
                    skey[lo] = b;
                    skey[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < nlist; i += blockDim.x) {
        unsigned long long key = skey[i];
        uint32_t o = (uint32_t)(key >> 32);
        if (metric == METRIC_IP) o = ~o;
        out_dis[q * nlist + i] = ord2f(o);
        out_keys[q * nlist + i] = (int)(key & 0xffffffffu);
        // equal distances: the reference's order among them is whatever its heap produces
        if (i + 1 < nlist && (uint32_t)(skey[i + 1] >> 32) == (uint32_t)(key >> 32)) atomicMin(&s_tie, i);
    }
    __syncthreads();
    if (threadIdx.x == 0) tie0[q] = s_tie;
}

// The same ranking for P >= 256 with E keys per thread held in registers: compare-exchange
// distances below E stay in registers, distances below 32 E go through warp shuffles, only the
// longer ones through shared memory (10 of the 78 network stages at P = 4096).  Shared-memory
// element i lives at i ^ (((i >> 4) & 7) << 1): both the blocked (thread t owns [tE, tE+E)) and
// the strided access patterns are then bank-conflict free.
__device__ __forceinline__ int rr_sw(int i) { return i ^ (((i >> 4) & 7) << 1); }

// The bitonic network on E keys per thread (T threads, P = T * E keys): compare-exchange distances below E stay
// in registers, distances below 32 E go through warp shuffles, only the longer ones through shared memory.
template <int E>
__device__ __forceinline__ void block_bitonic_sort(unsigned long long (&e)[E], unsigned long long* skey, int P, int t,
                                                   int T) {
    const int lane = t & 31, base = t * E;
    for (int size = 2; size <= P; size <<= 1) {
        int stride = size >> 1;
        if (stride >= 32 * E) {
            // long distances: through shared memory, E/2 pairs per thread and stage
#pragma unroll
            for (int r = 0; r < E; r += 2)
                *reinterpret_cast<ulonglong2*>(&skey[rr_sw(base + r)]) = make_ulonglong2(e[r], e[r + 1]);
            __syncthreads();
            for (; stride >= 32 * E; stride >>= 1) {
#pragma unroll
                for (int j = 0; j < E / 2; j++) {
                    const int p = t + j * T;
                    const int lo = 2 * p - (p & (stride - 1)), hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const unsigned long long a = skey[rr_sw(lo)], b = skey[rr_sw(hi)];
                    if ((a > b) == up) {
                        skey[rr_sw(lo)] = b;
                        skey[rr_sw(hi)] = a;
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < E; r += 2) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&skey[rr_sw(base + r)]);
                e[r] = v.x;
                e[r + 1] = v.y;
            }
        }
        for (; stride >= E; stride >>= 1) {  // partner element lives in lane ^ (stride / E), same register
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
    __syncthreads();
#pragma unroll
    for (int r = 0; r < E; r += 2)
        *reinterpret_cast<ulonglong2*>(&skey[rr_sw(base + r)]) = make_ulonglong2(e[r], e[r + 1]);
    __syncthreads();
}

// Full ranking.  qlist / sorted_upto (both nullable): "extension" launch -- block b serves query qlist[b]
// (or b), does nothing if the query's row is already ranked up to `need_upto` (or completely), and otherwise
// writes only the ranks >= sorted_upto[q]: the prefix a partial ranking (below) produced is kept as it is.
template <int E>
__global__ void __launch_bounds__(1024)
rank_rows_reg_kernel(int metric, const float* __restrict__ dis, long nlist, int P, float* __restrict__ out_dis,
                     int* __restrict__ out_keys, int* __restrict__ tie0, const int* __restrict__ qlist,
                     int* __restrict__ sorted_upto, const int* __restrict__ qbound, int need_upto) {
    extern __shared__ __align__(16) unsigned long long skey[];
    __shared__ int s_tie;
    const int t = threadIdx.x, T = blockDim.x;
    const long q = qlist ? qlist[blockIdx.x] : blockIdx.x;
    if (q < 0) return;
    int start = 0;
    if (sorted_upto) {
        start = sorted_upto[q];
        const int need = qbound ? min(need_upto, qbound[q] + 1) : need_upto;
        if (start >= nlist || start >= need) return;
    }
    if (t == 0) s_tie = 0x7fffffff;
    const float* row = dis + q * nlist;
    const int base = t * E;
    unsigned long long e[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int i = base + r;
        unsigned long long key = ~0ull;
        if (i < nlist) {
            uint32_t o = f2ord(row[i]);
            if (metric == METRIC_IP) o = ~o;
            key = ((unsigned long long)o << 32) | (unsigned)i;
        }
        e[r] = key;
    }
    block_bitonic_sort<E>(e, skey, P, t, T);
    for (int i = t; i < nlist; i += T) {
        if (i < start) continue;  // ranks a partial ranking (or nothing) already put in place
        const unsigned long long key = skey[rr_sw(i)];
        uint32_t o = (uint32_t)(key >> 32);
        if (metric == METRIC_IP) o = ~o;
        out_dis[q * nlist + i] = ord2f(o);
        out_keys[q * nlist + i] = (int)(key & 0xffffffffu);
        if (i + 1 < nlist && (uint32_t)(skey[rr_sw(i + 1)] >> 32) == (uint32_t)(key >> 32)) atomicMin(&s_tie, i);
    }
    __syncthreads();
    if (t == 0) {
        // extension: an earlier tie (still pending) stays the first one
        if (start == 0 || tie0[q] == 0x7fffffff) tie0[q] = s_tie;
        if (sorted_upto) sorted_upto[q] = (int)nlist;
    }
}

// Partial ranking for Auncel mode (nprobe = nlist): a query probes my_nprobe lists -- a few hundred of
// thousands --, so only the M best centroids are ranked up front.  The M-th smallest key is found by a radix
// select over the 32-bit order-preserving image of the distance (4 passes, 256-bin shared-memory histograms);
// every centroid at or below it (ties included, so the cut never splits a run of equal distances) is
// compacted and ranked by the same network.  sorted_upto[q] = number of ranks written; rows that need more
// are completed by the extension launch of rank_rows_reg_kernel before the round that reads them.
constexpr int RP_M = 1024, RP_T = 128, RP_E = 8;

__global__ void __launch_bounds__(RP_T)
rank_rows_partial_kernel(int metric, const float* __restrict__ dis, long nlist, float* __restrict__ out_dis,
                         int* __restrict__ out_keys, int* __restrict__ tie0, int* __restrict__ sorted_upto) {
    __shared__ __align__(16) unsigned long long skey[RP_M];
    __shared__ int hist[256];
    __shared__ int s_cnt, s_tie, s_digit, s_before;
    const int t = threadIdx.x;
    const long q = blockIdx.x;
    const float* row = dis + q * nlist;
    // ---- radix select of the RP_M-th smallest ord
    uint32_t prefix = 0, mask = 0;
    int remaining = RP_M;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = t; i < 256; i += RP_T) hist[i] = 0;
        __syncthreads();
        for (long i = t; i < nlist; i += RP_T) {
            uint32_t o = f2ord(row[i]);
            if (metric == METRIC_IP) o = ~o;
            if ((o & mask) == prefix) atomicAdd(&hist[(o >> shift) & 255], 1);
        }
        __syncthreads();
        if (t < 32) {  // one warp: 8 bins per lane, find the bin where the running count reaches `remaining`
            int loc[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                loc[j] = hist[t * 8 + j];
                sum += loc[j];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (t >= o) incl += up;
            }
            int before = incl - sum;
            if (before < remaining && remaining <= incl) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (before < remaining && remaining <= before + loc[j]) {
                        s_digit = t * 8 + j;
                        s_before = before;
                    }
                    before += loc[j];
                }
            }
        }
        __syncthreads();
        prefix |= (uint32_t)s_digit << shift;
        mask |= 0xffu << shift;
        remaining -= s_before;
        __syncthreads();
    }
    const uint32_t pivot = prefix;  // the RP_M-th smallest ord (exact)
    // ---- compact everything at or below the pivot
    if (t == 0) {
        s_cnt = 0;
        s_tie = 0x7fffffff;
    }
    for (int i = t; i < RP_M; i += RP_T) skey[i] = ~0ull;
    __syncthreads();
    for (long i = t; i < nlist; i += RP_T) {
        uint32_t o = f2ord(row[i]);
        if (metric == METRIC_IP) o = ~o;
        if (o <= pivot) {
            const int pos = atomicAdd(&s_cnt, 1);
            if (pos < RP_M) skey[pos] = ((unsigned long long)o << 32) | (unsigned)i;
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt > RP_M) {  // a long run of equal distances at the cut: leave the row to the full ranking
        if (t == 0) {
            sorted_upto[q] = 0;
            tie0[q] = 0x7fffffff;
        }
        return;
    }
    unsigned long long e[RP_E];
#pragma unroll
    for (int r = 0; r < RP_E; r++) e[r] = skey[t * RP_E + r];
    __syncthreads();
    block_bitonic_sort<RP_E>(e, skey, RP_M, t, RP_T);
    for (int i = t; i < cnt; i += RP_T) {
        const unsigned long long key = skey[rr_sw(i)];
        uint32_t o = (uint32_t)(key >> 32);
        if (metric == METRIC_IP) o = ~o;
        out_dis[q * nlist + i] = ord2f(o);
        out_keys[q * nlist + i] = (int)(key & 0xffffffffu);
        if (i + 1 < cnt && (uint32_t)(skey[rr_sw(i + 1)] >> 32) == (uint32_t)(key >> 32)) atomicMin(&s_tie, i);
    }
    __syncthreads();
    if (t == 0) {
        tie0[q] = s_tie;
        sorted_upto[q] = cnt;
    }
}

// ---------------------------------------------------------------------------------
// Exact tie order.  IndexFlat::search ranks centroids with a binary heap of size k (= nprobe;
// nlist in Auncel mode): knn_L2sqr_sse / knn_inner_product_sse (utils.cpp:417-490) push every
// candidate that beats the heap top, then heap_reorder (Heap.h:295-322) pops them out.  Among
// EQUAL distances the output order is a by-product of the heap's layout -- and with thousands
// of centroids at float resolution ties are common (birthday effect).  The probe order decides
// what the termination check sees at each stage, so for the queries that have a tie inside
// the range that matters the heap is replayed literally: one thread walks the reference's
// algorithm in shared memory (Heap.h:88-142).
// Heap nodes are packed as (ord << 32 | id) with ord = f2ord(dis) for L2 (CMax: a > b) and
// ~f2ord(sim) for IP (CMin: a < b), so both metrics are a max-heap on ord; comparisons look at
// ord only -- ids never break ties, exactly like the reference's cmp on values.
__device__ __forceinline__ bool ngt(unsigned long long a, unsigned long long b) {
    return (uint32_t)(a >> 32) > (uint32_t)(b >> 32);
}

// Heap.h:88-117 on a 1-based array, started at node `i` (1 = the literal algorithm).  Two tree
// levels are resolved per shared-memory round trip: children and grandchildren are loaded
// together, then the reference's decisions are applied in order -- same moves, half the
// dependent latency.
__device__ __forceinline__ void heap_pop_dev(int k, unsigned long long* h, int i = 1) {
    const unsigned long long v = h[k];
    while (4 * i + 3 <= k) {
        const unsigned long long c1 = h[2 * i], c2 = h[2 * i + 1];
        const ulonglong2 ga = *reinterpret_cast<const ulonglong2*>(&h[4 * i]);      // children of 2i
        const ulonglong2 gb = *reinterpret_cast<const ulonglong2*>(&h[4 * i + 2]);  // children of 2i+1
        const bool left = ngt(c1, c2);
        const unsigned long long c = left ? c1 : c2;
        if (ngt(v, c)) {
            h[i] = v;
            return;
        }
        h[i] = c;
        i = left ? 2 * i : 2 * i + 1;
        const unsigned long long g1 = left ? ga.x : gb.x, g2 = left ? ga.y : gb.y;
        const bool gleft = ngt(g1, g2);
        const unsigned long long g = gleft ? g1 : g2;
        if (ngt(v, g)) {
            h[i] = v;
            return;
        }
        h[i] = g;
        i = gleft ? 2 * i : 2 * i + 1;
    }
    while (true) {
        const int i1 = i << 1, i2 = i1 + 1;
        if (i1 > k) break;
        const unsigned long long c1 = h[i1], c2 = h[i2 <= k ? i2 : i1];
        const bool left = (i2 == k + 1) || ngt(c1, c2);
        const unsigned long long c = left ? c1 : c2;
        if (ngt(v, c)) break;
        h[i] = c;
        i = left ? i1 : i2;
    }
    h[i] = v;
}

// Heap.h:124-142
__device__ __forceinline__ void heap_push_dev(int k, unsigned long long* h, unsigned long long v) {
    int i = k;
    while (i > 1) {
        const int f = i >> 1;
        const unsigned long long p = h[f];
        if (!ngt(v, p)) break;
        h[i] = p;
        i = f;
    }
    h[i] = v;
}

// While neutral (FLT_MAX) slots remain, heap_pop walks from the root through neutral nodes --
// right child when both are neutral, the neutral one otherwise -- moving neutral onto neutral
// until it meets real values.  That prefix does not depend on the data, only on how many
// elements were inserted: `entry[j]` is the node where insertion j meets its first real
// comparison (precomputed on the host, heap_entry_table), so the replay starts there.
__device__ __forceinline__ uint32_t node_ord(unsigned long long a) { return (uint32_t)(a >> 32); }

// One level of heap_pop's sift-down (Heap.h:96-113) for the pop a lane carries; see the pipeline
// in heap_order_kernel.  Children are one aligned 16-byte load (the heap array is 1-based).
struct PopLane {
    unsigned long long v, top;
    int node, size, depth;
    bool busy;
};
__device__ __forceinline__ void pop_step(PopLane& p, unsigned long long* h, int k) {
    // written without branches: a lone warp pays ~20 cycles per taken branch
    const int i1 = p.node << 1;
    const ulonglong2 cc = *reinterpret_cast<const ulonglong2*>(&h[min(i1, k) & ~1]);  // in bounds; unused if i1 > size
    const bool has1 = i1 <= p.size, has2 = i1 < p.size;
    const bool left = !has2 | (node_ord(cc.x) > node_ord(cc.y));
    const unsigned long long c = left ? cc.x : cc.y;
    const bool stop = !has1 | (node_ord(p.v) > node_ord(c));
    if (p.busy) h[p.node] = stop ? p.v : c;
    if (p.busy & stop) h[p.size] = p.top;  // slot `size` is free after the pop: bh_val[k-ii-1] = top (Heap.h:305-317)
    p.busy = p.busy & !stop;
    p.node = stop ? p.node : i1 + (left ? 0 : 1);
    p.depth += stop ? 0 : 1;
}

// A single warp is latency-bound on its own dependent instructions (~5 cycles each), so the
// replay is written for few instructions per heap level, not for few memory round trips.
template <int METRIC>
__global__ void __launch_bounds__(32)
heap_order_kernel(const float* __restrict__ raw, long nlist, int k, const int* __restrict__ entry,
                  const int* __restrict__ fix_list, const int* __restrict__ nfix, float* __restrict__ out_dis,
                  int* __restrict__ out_keys, int* __restrict__ tie0, const int* __restrict__ qbound,
                  int* __restrict__ sorted_upto) {
    if ((int)blockIdx.x >= *nfix) return;
    extern __shared__ __align__(16) unsigned long long hp[];  // hp[0] unused: 1-based heap, 16 B aligned pairs
    unsigned long long* h = hp;
    const long q = fix_list[blockIdx.x] & 0x3fffffff;
    const bool straddle = (fix_list[blockIdx.x] >> 30) & 1;  // only the tie group around the query's stop stage
    const int lane = threadIdx.x;
    const float neut = METRIC == METRIC_L2 ? FLT_MAX : -FLT_MAX;
    uint32_t on = f2ord(neut);
    if (METRIC == METRIC_IP) on = ~on;
    const unsigned long long neutral = ((unsigned long long)on << 32) | 0xffffffffu;  // heapify, id -1
    for (int i = lane; i <= k + 1; i += 32) hp[i] = neutral;
    __syncwarp();
#ifdef HEAP_TIMING
    long long t0 = clock64();
#endif
    // Fill phase.  Insertion j (0 < j < k) pops with v = h[k] = d[j-1] (the previous push stays at
    // slot k as long as its parent is neutral), drops it into node entry[j] and sifts it down inside
    // that node's subtree, whose nodes were all filled earlier.  Sift-downs in disjoint subtrees
    // commute, so for the prefix j <= J whose entries stay off the root-to-k path (J = entry[k],
    // heap_entry_table; all but the last log2(k)+1 insertions when k is a power of two) the warp
    // places every d[j-1] at once and sifts level by level, deepest first -- Floyd's heapify with
    // the reference's comparisons.  The rest runs serially below.
    const float* row = raw + q * nlist;
    int J = min(entry[k], (int)min((long)k, nlist) - 1);
    {
        bool bad = false;  // a distance that does not beat the neutral top is not inserted: no shortcut
        for (int j = lane; j <= J; j += 32) {
            uint32_t o = f2ord(__ldg(row + j));
            if (METRIC == METRIC_IP) o = ~o;
            bad |= !(o < on);
        }
        if (__any_sync(0xffffffffu, bad)) J = 0;
    }
    if (J >= 1) {
        for (int j = 1 + lane; j <= J; j += 32) {
            uint32_t o = f2ord(__ldg(row + j - 1));
            if (METRIC == METRIC_IP) o = ~o;
            h[__ldg(entry + j)] = ((unsigned long long)o << 32) | (unsigned)(j - 1);
        }
        __syncwarp();
        const int dk = 31 - __clz(k);
        for (int t = dk - 1; t >= 0; t--) {
            const int first = 1 << t;
            for (int base = 0; base < first; base += 32) {
                int i = first + base + lane;
                const unsigned long long v = base + lane < first ? h[i] : neutral;
                bool active = v != neutral;
                for (int lev = t; lev < dk && __any_sync(0xffffffffu, active); lev++) {
                    const int i1 = i << 1;
                    const ulonglong2 cc = *reinterpret_cast<const ulonglong2*>(&h[min(i1, k) & ~1]);
                    const bool left = node_ord(cc.x) > node_ord(cc.y);
                    const unsigned long long c = left ? cc.x : cc.y;
                    const bool stop = (i1 > k) | (node_ord(v) > node_ord(c));
                    if (active) h[i] = stop ? v : c;
                    active = active & !stop;
                    i = stop ? i : i1 + (left ? 0 : 1);
                }
                if (active) h[i] = v;  // unreachable: a sift-down ends at a leaf at the latest
            }
            __syncwarp();
        }
        if (lane == 0) {
            uint32_t o = f2ord(__ldg(row + J));
            if (METRIC == METRIC_IP) o = ~o;
            h[k] = ((unsigned long long)o << 32) | (unsigned)J;  // push J stays at slot k
        }
        __syncwarp();
    }
    if (lane == 0) {
        constexpr int PF = 8;  // software prefetch of the distance row
        const long jstart = J >= 1 ? J + 1 : 0;
        float buf[PF];
        int ebuf[PF];
#pragma unroll
        for (int t = 0; t < PF; t++) {
            buf[t] = jstart + t < nlist ? __ldg(row + jstart + t) : 0.f;
            ebuf[t] = jstart + t < k ? __ldg(entry + jstart + t) : 1;
        }
        unsigned long long last = h[k];  // kept in a register
        for (long j0 = jstart; j0 < nlist; j0 += PF) {
            float cur[PF];
            int ecur[PF];
#pragma unroll
            for (int t = 0; t < PF; t++) {
                cur[t] = buf[t];
                ecur[t] = ebuf[t];
                long nj = j0 + PF + t;
                buf[t] = nj < nlist ? __ldg(row + nj) : 0.f;
                ebuf[t] = nj < k ? __ldg(entry + nj) : 1;
            }
#pragma unroll
            for (int t = 0; t < PF; t++) {
                long j = j0 + t;
                if (j >= nlist) break;
                uint32_t o = f2ord(cur[t]);
                if (METRIC == METRIC_IP) o = ~o;
                const unsigned long long nv = ((unsigned long long)o << 32) | (unsigned)j;
                if (j < k) {
                    // the root is still neutral: dis < simi[0] (utils.cpp:441,479) unless dis is FLT_MAX itself
                    if (!(o < on)) continue;
                    // heap_pop from the end of the neutral prefix, one level per iteration
                    int i = ecur[t];
                    while (true) {
                        const int i1 = i << 1;
                        if (i1 > k) break;
                        const ulonglong2 cc = *reinterpret_cast<const ulonglong2*>(&h[i1]);
                        const bool left = (i1 + 1 > k) || node_ord(cc.x) > node_ord(cc.y);
                        const unsigned long long c = left ? cc.x : cc.y;
                        if (node_ord(last) > node_ord(c)) break;
                        h[i] = c;
                        i = i1 + (left ? 0 : 1);
                    }
                    h[i] = last;
                    // heap_push at slot k (Heap.h:124-142)
                    i = k;
                    last = nv;
                    while (i > 1) {
                        const int f = i >> 1;
                        const unsigned long long par = h[f];
                        if (!(o > node_ord(par))) break;
                        h[i] = par;
                        if (i == k) last = par;
                        i = f;
                    }
                    h[i] = nv;
                } else if (o < node_ord(h[1])) {  // replacement phase (nprobe < nlist): full-depth pops
                    heap_pop_dev(k, h, 1);
                    heap_push_dev(k, h, nv);
                }
            }
        }
    }
    __syncwarp();
#ifdef HEAP_TIMING
    long long t1 = clock64();
#endif
    // heap_reorder, Heap.h:295-322 (every slot holds a real element because k <= nlist): pop p
    // takes top = h[1], sifts v = h[k-p] down a heap of size k-p, then stores top in slot k-p.
    // The k sift-downs are the bulk of the replay and each is a chain of dependent steps, so they
    // are software-pipelined over the lanes: a new pop enters at the root every second tick while
    // the earlier ones are still on their way down.  Every lane moves one level per tick; lanes are
    // therefore >= 2 levels apart, a lane at depth t reads depth t+1 (final: the lane ahead wrote
    // it a tick ago) and writes depth t.  The only other dependence is v = h[k-p], a bottom slot an
    // earlier pop may still be heading for: a pop starts only when no lane in flight sits on an
    // ancestor of that slot (a lane that has left the ancestor chain can never come back to it).
    // Pop p runs on lane p % 32: a pop lasts at most log2(k)+2 ticks, so that lane is idle again.
    {
        PopLane pl;
        pl.busy = false;
        pl.node = 1;
        pl.size = 0;
        pl.depth = 0;
        pl.v = pl.top = 0;
        int next_pop = 0;
        while (next_pop < k) {
            const int s = k - next_pop;
            const int ds = 31 - __clz(s);
            const bool anc = pl.busy && pl.depth <= ds && (s >> (ds - pl.depth)) == pl.node;
            if (!__any_sync(0xffffffffu, anc)) {
                if (lane == (next_pop & 31)) {
                    pl.busy = true;
                    pl.node = 1;
                    pl.depth = 0;
                    pl.size = s;
                    pl.v = h[s];
                    pl.top = h[1];
                }
                next_pop++;
            }
            pop_step(pl, h, k);
            __syncwarp();
            pop_step(pl, h, k);
            __syncwarp();
        }
        const int drain = 34 - __clz(k);
        for (int t = 0; t < drain; t++) {
            pop_step(pl, h, k);
            __syncwarp();
        }
    }
    __syncwarp();
#ifdef HEAP_TIMING
    if (lane == 0 && blockIdx.x == 0)
        printf("heap_order: nfix %d k %d phase1 %lld cycles, phase2 %lld cycles\n", *nfix, k, t1 - t0, clock64() - t1);
#endif
    int w_lo = 0, w_hi = k - 1;
    if (straddle) {
        // A decided query scans ranks [0, b): equal distances strictly inside that range cannot change what
        // is scanned, only a tie between rank b-1 and rank b can.  Ranks below the group may already have
        // been scanned in the provisional order (finalize maps rank -> list), so only the group is rewritten.
        const int b = qbound[q];
        const uint32_t o = node_ord(hp[b]);  // rank b-1 lives in hp[b]
        w_lo = w_hi = b - 1;
        while (w_lo > 0 && node_ord(hp[w_lo]) == o) w_lo--;
        while (w_hi + 1 < k && node_ord(hp[w_hi + 2]) == o) w_hi++;
    }
    for (int i = w_lo + lane; i <= w_hi; i += 32) {
        const unsigned long long node = hp[i + 1];
        uint32_t o = (uint32_t)(node >> 32);
        if (METRIC == METRIC_IP) o = ~o;
        out_dis[q * nlist + i] = ord2f(o);
        out_keys[q * nlist + i] = (int)(uint32_t)(node & 0xffffffffu);
    }
    if (lane == 0) {
        tie0[q] = 0x7fffffff;
        if (sorted_upto && !straddle) sorted_upto[q] = k;  // the replay ranks the whole row
    }
}

// entry[j], j < k: see heap_order_kernel.  Pure structure: simulate heap_pop on occupancy bits.
void heap_entry_table(int k, std::vector<int>& entry) {
    entry.assign(k + 1, 1);
    std::vector<char> real(k + 2, 0);
    for (int j = 0; j < k; j++) {
        // insertion j pops with v = h[k] (neutral for j == 0, real afterwards), then pushes at slot k
        int i = 1;
        if (j == 0) {
            entry[0] = 1;  // v is neutral: the literal walk only shuffles neutral values
        } else {
            while (true) {
                int i1 = i << 1, i2 = i1 + 1;
                if (i1 > k) break;  // leaf of the neutral region: v lands here
                bool left;
                if (i2 == k + 1) left = true;
                else if (real[i1] && real[i2]) break;          // first data-dependent comparison
                else left = !real[i1] && real[i2];             // cmp(h[i1], h[i2]): neutral beats real, tie -> right
                int c = left ? i1 : i2;
                if (real[c]) break;                             // single real child (slot k's stale copy)
                i = c;
            }
            entry[j] = i;
            real[i] = 1;  // the hole of this pop ends here: one more real node
        }
        real[k] = 1;      // push places d_j at slot k (and may sift up only through real parents)
    }
    // entry[k] = J: the longest prefix 1..J of insertions whose entry node is not on the path from
    // the root to slot k.  Until such a node is filled, slot k's parent is neutral (pushes stay at
    // slot k) and every pop works in a subtree that neither contains slot k nor a neutral node.
    int J = k - 1;
    int dk = 0;
    while ((k >> (dk + 1)) > 0) dk++;
    for (int j = 1; j < k; j++) {
        int n = entry[j], dn = 0;
        while ((n >> (dn + 1)) > 0) dn++;
        if ((k >> (dk - dn)) == n) {
            J = j - 1;
            break;
        }
    }
    entry[k] = std::max(J, 0);
}

// queries (from `list`, or all n when list == nullptr) whose first tie lies below `bound`
__global__ void collect_ties_kernel(const int* __restrict__ list, int n, const int* __restrict__ tie0, int bound,
                                    const int* __restrict__ qbound, const int* __restrict__ decided, int r0, int k,
                                    long nlist, const float* __restrict__ dis, int* __restrict__ fix_list,
                                    int* __restrict__ nfix, int* __restrict__ err) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    int q = list ? list[a] : a;
    if (decided && decided[q]) {
        // stop stage b known: the order inside [0, b) is irrelevant (every list in it is scanned, the check
        // no longer runs); what must follow the reference's heap is WHICH lists are in it, i.e. a tie
        // between rank b-1 and rank b -- looked at in the round that scans rank b-1
        const int b = qbound[q];
        if (tie0[q] != 0x7fffffff && b > r0 && b <= bound && b < k &&
            dis[(long)q * nlist + b - 1] == dis[(long)q * nlist + b]) {
            // the run of equal distances reaches back into ranks that were scanned in the provisional order
            if (err && r0 > 0 && dis[(long)q * nlist + r0 - 1] == dis[(long)q * nlist + b]) atomicOr(err, ERR_TIE_SPAN);
            fix_list[atomicAdd(nfix, 1)] = q | (1 << 30);
        }
        return;
    }
    int b = qbound ? min(bound, qbound[q]) : bound;  // a decided query never scans past its stop stage
    if (tie0[q] < b) fix_list[atomicAdd(nfix, 1)] = q;
}

void launch_fix_ties(int metric, const float* raw, long nlist, int k, const int* entry, const int* list, int n,
                     int* tie0, int bound, const int* qbound, int* fix_list, int* nfix, float* out_dis, int* out_keys,
                     cudaStream_t s, const int* decided, int r0, int* err, int* sorted_upto) {
    if (n == 0) return;
    CUDA_CHECK(cudaMemsetAsync(nfix, 0, sizeof(int), s));
    collect_ties_kernel<<<(n + 255) / 256, 256, 0, s>>>(list, n, tie0, bound, qbound, decided, r0, k, nlist, out_dis,
                                                        fix_list, nfix, err);
    size_t smem = (size_t)(k + 2) * 8;
    AUNCEL_CHECK(smem <= 220 * 1024, "nlist too large for the exact tie replay");
    auto kern = metric == METRIC_L2 ? heap_order_kernel<METRIC_L2> : heap_order_kernel<METRIC_IP>;
    if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n, 32, smem, s>>>(raw, nlist, k, entry, fix_list, nfix, out_dis, out_keys, tie0, qbound, sorted_upto);
    CUDA_CHECK(cudaGetLastError());
}

void launch_rank_rows(int metric, const float* dis, long nq, long nlist, float* out_dis, int* out_keys,
                      int* tie0, cudaStream_t s, const int* qlist, int* sorted_upto, const int* qbound, int need_upto) {
    if (nq == 0) return;
    int P = 1;
    while (P < nlist) P <<= 1;
    size_t smem = (size_t)P * sizeof(unsigned long long);
    AUNCEL_CHECK(smem <= 220 * 1024, "nlist too large for the in-smem centroid ranking");
    if (P >= 256) {  // register / shuffle / shared-memory hybrid
        auto kern = P <= 8192 ? rank_rows_reg_kernel<8> : rank_rows_reg_kernel<16>;
        const int threads = P <= 8192 ? P / 8 : P / 16;
        AUNCEL_CHECK(threads <= 1024, "nlist too large for the in-smem centroid ranking");
        if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)nq, threads, smem, s>>>(metric, dis, nlist, P, out_dis, out_keys, tie0, qlist, sorted_upto, qbound,
                                                 need_upto);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    AUNCEL_CHECK(qlist == nullptr && sorted_upto == nullptr, "partial ranking needs nlist >= 256");
    if (smem > 48 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(rank_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = std::max(32, std::min(1024, P / 2));
    rank_rows_kernel<<<(unsigned)nq, threads, smem, s>>>(metric, dis, nlist, P, out_dis, out_keys, tie0);
    CUDA_CHECK(cudaGetLastError());
}

// the RP_M best centroids per query (ties at the cut included); rows it cannot handle get sorted_upto = 0
void launch_rank_rows_partial(int metric, const float* dis, long nq, long nlist, float* out_dis, int* out_keys,
                              int* tie0, int* sorted_upto, cudaStream_t s) {
    if (nq == 0) return;
    rank_rows_partial_kernel<<<(unsigned)nq, RP_T, 0, s>>>(metric, dis, nlist, out_dis, out_keys, tie0, sorted_upto);
    CUDA_CHECK(cudaGetLastError());
}

int rank_rows_partial_width() { return RP_M; }

}  // namespace auncel
