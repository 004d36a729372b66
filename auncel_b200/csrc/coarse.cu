// Subsystem (1): coarse assignment.
//   dense_exact_kernel : query x centroid distances with the reference's exact arithmetic
//                        (fvec_L2sqr / fvec_inner_product SSE order, utils_simd.cpp:391-443 --
//                        the knn_*_sse path of utils.cpp:417-490 that IndexFlat::search takes
//                        for nx < 20 and that the parity harness forces for batches)
//   rank_rows_kernel   : full ranking of all nlist centroids per query (heap_reorder output of
//                        IndexFlat::search with k = nprobe = nlist, profile.cpp:218-222)
//   the same tile kernel also fills interdis_cem (IVF_pro.cpp:21-39) and does add()'s k=1 assign.
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
                 int* __restrict__ out_keys) {
    extern __shared__ unsigned long long skey[];
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
    }
}

void launch_rank_rows(int metric, const float* dis, long nq, long nlist, float* out_dis, int* out_keys,
                      cudaStream_t s) {
    if (nq == 0) return;
    int P = 1;
    while (P < nlist) P <<= 1;
    size_t smem = (size_t)P * sizeof(unsigned long long);
    AUNCEL_CHECK(smem <= 220 * 1024, "nlist too large for the in-smem centroid ranking");
    if (smem > 48 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(rank_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = std::max(32, std::min(1024, P / 2));
    rank_rows_kernel<<<(unsigned)nq, threads, smem, s>>>(metric, dis, nlist, P, out_dis, out_keys);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace auncel
