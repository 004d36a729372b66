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
template <int METRIC>
__global__ void __launch_bounds__(32)
heap_order_kernel(const float* __restrict__ raw, long nlist, int k, const int* __restrict__ entry,
                  const int* __restrict__ fix_list, const int* __restrict__ nfix, float* __restrict__ out_dis,
                  int* __restrict__ out_keys, int* __restrict__ tie0) {
    if ((int)blockIdx.x >= *nfix) return;
    extern __shared__ __align__(16) unsigned long long hp[];  // hp[0] unused: 1-based heap, 16 B aligned pairs
    unsigned long long* h = hp;
    const long q = fix_list[blockIdx.x];
    const int lane = threadIdx.x;
    const float neut = METRIC == METRIC_L2 ? FLT_MAX : -FLT_MAX;
    uint32_t on = f2ord(neut);
    if (METRIC == METRIC_IP) on = ~on;
    for (int i = lane; i <= k; i += 32) hp[i] = ((unsigned long long)on << 32) | 0xffffffffu;  // heapify, id -1
    __syncwarp();
    if (lane == 0) {
        const float* row = raw + q * nlist;
        constexpr int PF = 8;  // software prefetch of the distance row
        float buf[PF];
        int ebuf[PF];
#pragma unroll
        for (int t = 0; t < PF; t++) {
            buf[t] = t < nlist ? __ldg(row + t) : 0.f;
            ebuf[t] = t < k ? __ldg(entry + t) : 1;
        }
        for (long j0 = 0; j0 < nlist; j0 += PF) {
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
                if (o < (uint32_t)(h[1] >> 32)) {  // dis < simi[0] (L2) / ip > simi[0] (IP), utils.cpp:441,479
                    heap_pop_dev(k, h, j < k ? ecur[t] : 1);
                    heap_push_dev(k, h, ((unsigned long long)o << 32) | (unsigned)j);
                }
            }
        }
        // heap_reorder, Heap.h:295-322 (every slot holds a real element because k <= nlist)
        for (int i = 0; i < k; i++) {
            const unsigned long long top = h[1];
            heap_pop_dev(k - i, h);
            h[k - i] = top;  // slot k-i is free after the pop: same as bh_val[k-ii-1] with ii == i
        }
    }
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
        const unsigned long long node = hp[i + 1];
        uint32_t o = (uint32_t)(node >> 32);
        if (METRIC == METRIC_IP) o = ~o;
        out_dis[q * nlist + i] = ord2f(o);
        out_keys[q * nlist + i] = (int)(uint32_t)(node & 0xffffffffu);
    }
    if (lane == 0) tie0[q] = 0x7fffffff;
}

// entry[j], j < k: see heap_order_kernel.  Pure structure: simulate heap_pop on occupancy bits.
void heap_entry_table(int k, std::vector<int>& entry) {
    entry.assign(k, 1);
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
}

// queries (from `list`, or all n when list == nullptr) whose first tie lies below `bound`
__global__ void collect_ties_kernel(const int* __restrict__ list, int n, const int* __restrict__ tie0, int bound,
                                    const int* __restrict__ qbound, int* __restrict__ fix_list,
                                    int* __restrict__ nfix) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    int q = list ? list[a] : a;
    int b = qbound ? min(bound, qbound[q]) : bound;  // a decided query never scans past its stop stage
    if (tie0[q] < b) fix_list[atomicAdd(nfix, 1)] = q;
}

void launch_fix_ties(int metric, const float* raw, long nlist, int k, const int* entry, const int* list, int n,
                     int* tie0, int bound, const int* qbound, int* fix_list, int* nfix, float* out_dis, int* out_keys,
                     cudaStream_t s) {
    if (n == 0) return;
    CUDA_CHECK(cudaMemsetAsync(nfix, 0, sizeof(int), s));
    collect_ties_kernel<<<(n + 255) / 256, 256, 0, s>>>(list, n, tie0, bound, qbound, fix_list, nfix);
    size_t smem = (size_t)(k + 2) * 8;
    AUNCEL_CHECK(smem <= 220 * 1024, "nlist too large for the exact tie replay");
    auto kern = metric == METRIC_L2 ? heap_order_kernel<METRIC_L2> : heap_order_kernel<METRIC_IP>;
    if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n, 32, smem, s>>>(raw, nlist, k, entry, fix_list, nfix, out_dis, out_keys, tie0);
    CUDA_CHECK(cudaGetLastError());
}

void launch_rank_rows(int metric, const float* dis, long nq, long nlist, float* out_dis, int* out_keys,
                      int* tie0, cudaStream_t s) {
    if (nq == 0) return;
    int P = 1;
    while (P < nlist) P <<= 1;
    size_t smem = (size_t)P * sizeof(unsigned long long);
    AUNCEL_CHECK(smem <= 220 * 1024, "nlist too large for the in-smem centroid ranking");
    if (smem > 48 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(rank_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = std::max(32, std::min(1024, P / 2));
    rank_rows_kernel<<<(unsigned)nq, threads, smem, s>>>(metric, dis, nlist, P, out_dis, out_keys, tie0);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
