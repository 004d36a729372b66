// Index::train for the coarse quantizer: k-means with every data-sized step on the device.
//
// Follows Clustering::train (/root/reference/Auncel/Clustering.cpp:77-244) and
// km_update_centroids (utils.cpp:1078-1161) to the bit:
//   * seeded sub-sampling and initial centroids: rand_perm over std::mt19937 (utils.cpp:229-241,
//     RandomGenerator :110-133) -- a host-side permutation, the rows are gathered on the device;
//   * assignment = index.search(nx, x, 1): first strictly-best centroid in the reference's exact
//     arithmetic (dense_exact_kernel, OUT_BEST);
//   * update: a centroid is the float sum of its points IN INPUT ORDER divided by their count
//     (the reference's per-thread loop walks i = 0..n for its centroid range, utils.cpp:1096-1108, so
//     the order is the input order for any thread count).  Here: one warp per centroid walks the
//     assignment array 32 entries at a time (ballot), and adds the rows of its members in order, lanes
//     across dimensions -- a segmented sum without sorting;
//   * void clusters (utils.cpp:1121-1157): the choice of the cluster to split is a sequential draw on
//     the cluster sizes (k integers) and stays on the host; the copy + symmetric perturbation is
//     applied on the device;
//   * spherical k-means for inner product (IndexIVF.cpp:160-162, fvec_renorm_L2 utils.cpp:377-392).
// The training set and the centroids never leave the device between iterations.
#include <algorithm>
#include <cstring>
#include <random>
#include <vector>

#include "engine.h"

namespace auncel {

namespace {

__global__ void km_nonfinite_kernel(const float* __restrict__ x, size_t n, int* __restrict__ flag) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    bool bad = false;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) bad |= !isfinite(x[i]);
    if (bad) *flag = 1;
}

__global__ void km_gather_kernel(const float* __restrict__ x, int d, const int* __restrict__ rows, long m,
                                 float* __restrict__ out) {
    long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= m) return;
    const float* src = x + (size_t)rows[r] * d;
    for (int c = lane; c < d; c += 32) out[r * (size_t)d + c] = src[c];
}

__global__ void km_extract_kernel(const unsigned long long* __restrict__ best, long n, int* __restrict__ assign) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) assign[i] = (int)(best[i] & 0xffffffffull);
}

// One warp per centroid.  DCH = 32-float chunks of the row a lane keeps in registers per pass.
constexpr int KM_DCH = 8;  // 256 dimensions per pass over the assignment array

__global__ void __launch_bounds__(128)
km_accumulate_kernel(const float* __restrict__ x, long n, int d, const int* __restrict__ assign, long k,
                     float* __restrict__ cent, unsigned long long* __restrict__ hassign) {
    const long c = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= k) return;
    unsigned long long cnt = 0;
    for (int d0 = 0; d0 < d; d0 += 32 * KM_DCH) {
        float acc[KM_DCH];
#pragma unroll
        for (int t = 0; t < KM_DCH; t++) acc[t] = 0.f;
        cnt = 0;
        for (long i0 = 0; i0 < n; i0 += 32) {
            const long i = i0 + lane;
            unsigned m = __ballot_sync(0xffffffffu, i < n && assign[i] == (int)c);
            cnt += __popc(m);
            while (m) {  // members in input order
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const float* row = x + (size_t)(i0 + b) * d + d0;
#pragma unroll
                for (int t = 0; t < KM_DCH; t++) {
                    const int j = lane + 32 * t;
                    if (d0 + j < d) acc[t] = __fadd_rn(acc[t], row[j]);
                }
            }
        }
        const float ni = (float)cnt;  // utils.cpp:1114-1119
#pragma unroll
        for (int t = 0; t < KM_DCH; t++) {
            const int j = d0 + lane + 32 * t;
            if (j < d) cent[c * (size_t)d + j] = cnt ? __fdiv_rn(acc[t], ni) : 0.f;
        }
    }
    if (lane == 0) hassign[c] = cnt;
}

// utils.cpp:1134-1146: copy cj onto the void cluster ci, then perturb both symmetrically.  The
// operations are applied in order by one block (a cluster can be split more than once).
__global__ void km_split_kernel(float* __restrict__ cent, int d, const int2* __restrict__ ops, int nops) {
    const double EPS = 1 / 1024.;
    for (int o = 0; o < nops; o++) {
        const int ci = ops[o].x, cj = ops[o].y;
        for (int j = threadIdx.x; j < d; j += blockDim.x) {
            const float v = cent[(size_t)cj * d + j];
            const double up = 1 + EPS, dn = 1 - EPS;
            cent[(size_t)ci * d + j] = (float)__dmul_rn((double)v, (j % 2 == 0) ? up : dn);
            cent[(size_t)cj * d + j] = (float)__dmul_rn((double)v, (j % 2 == 0) ? dn : up);
        }
        __syncthreads();
    }
}

// fvec_renorm_L2 (utils.cpp:377-392) with fvec_norm_L2sqr's lane order (utils_simd.cpp:137-155):
// one thread per centroid for the norm (sequential over d/4 steps), then all threads scale.
__global__ void km_renorm_kernel(float* __restrict__ cent, int d, long k) {
    const long c = blockIdx.x;
    __shared__ float s_inv;
    __shared__ int s_scale;
    if (threadIdx.x == 0) {
        const float* xi = cent + c * (size_t)d;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        int j = 0;
        for (; j + 4 <= d; j += 4)
            for (int l = 0; l < 4; l++) s[l] = __fadd_rn(s[l], __fmul_rn(xi[j + l], xi[j + l]));
        for (int l = 0; l < 4; l++) {  // masked_read: the tail is zero-padded
            const float v = j + l < d ? xi[j + l] : 0.f;
            s[l] = __fadd_rn(s[l], __fmul_rn(v, v));
        }
        const float nr = __fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3]));
        s_scale = nr > 0;                                          // if (nr > 0), utils.cpp:384
        s_inv = s_scale ? (float)(1.0 / (double)sqrtf(nr)) : 1.f;  // const float inv_nr = 1.0 / sqrtf(nr)
    }
    __syncthreads();
    const float inv = s_inv;
    if (!s_scale) return;
    for (int j = threadIdx.x; j < d; j += blockDim.x) cent[c * (size_t)d + j] = __fmul_rn(cent[c * (size_t)d + j], inv);
}

void rand_perm(std::vector<int>& perm, size_t n, long seed) {  // utils.cpp:229-241
    perm.resize(n);
    for (size_t i = 0; i < n; i++) perm[i] = (int)i;
    std::mt19937 mt((unsigned int)seed);  // RandomGenerator, utils.cpp:110-111
    for (size_t i = 0; i + 1 < n; i++) {
        const int i2 = (int)(i + mt() % (int)(n - i));  // rand_int(max) = mt() % max, :123-126
        std::swap(perm[i], perm[i2]);
    }
}

}  // namespace

void train_kmeans(IvfIndex& ix, long nx, const float* x_in, int niter, bool tune) {
    const long k = ix.nlist;
    const int d = ix.d;
    AUNCEL_CHECK(nx >= k, "Number of training points should be at least as large as number of clusters");
    AUNCEL_CHECK(nx < (1L << 31), "too many training points");
    CUDA_CHECK(cudaSetDevice(ix.device));
    cudaStream_t s = ix.stream;
    const long max_pts = 256, seed = 1234;  // ClusteringParameters, Clustering.cpp:24-35

    // ---- training set -> device (sub-sampled first when it is larger than 256 points per centroid)
    DevBuf<float> x, cent;
    DevBuf<int> d_rows, d_assign, d_flag;
    DevBuf<unsigned long long> d_best, d_hassign;
    DevBuf<int2> d_ops;
    const float* x_src_host = x_in;
    std::vector<float> sub;
    if (nx > k * max_pts) {
        // the reference validates the whole input before sampling: exponent-all-ones test, 4-way OR
        const uint32_t* w = reinterpret_cast<const uint32_t*>(x_in);
        const size_t tot = (size_t)nx * d;
        bool bad = false;
        for (size_t i = 0; i < tot; i++) bad |= (w[i] & 0x7f800000u) == 0x7f800000u;
        AUNCEL_CHECK(!bad, "input contains NaN's or Inf's");
        std::vector<int> perm;
        rand_perm(perm, nx, seed);
        nx = k * max_pts;
        sub.resize((size_t)nx * d);
        for (long i = 0; i < nx; i++) memcpy(sub.data() + (size_t)i * d, x_in + (size_t)perm[i] * d, sizeof(float) * d);
        x_src_host = sub.data();
    }
    x.ensure((size_t)nx * d);
    CUDA_CHECK(cudaMemcpyAsync(x.p, x_src_host, (size_t)nx * d * sizeof(float), cudaMemcpyHostToDevice, s));
    // "input contains NaN's or Inf's" (Clustering.cpp:86-89 checks the whole input; so does this when no
    // sub-sampling happened, otherwise the sampled rows)
    d_flag.ensure(1);
    CUDA_CHECK(cudaMemsetAsync(d_flag.p, 0, sizeof(int), s));
    km_nonfinite_kernel<<<148 * 8, 256, 0, s>>>(x.p, (size_t)nx * d, d_flag.p);
    int h_flag = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h_flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    AUNCEL_CHECK(!h_flag, "input contains NaN's or Inf's");

    std::vector<float> h_cent((size_t)k * d);
    if (nx == k) {  // Clustering.cpp:113-123: just copy the training set
        memcpy(h_cent.data(), x_in, sizeof(float) * d * k);
        ix.set_centroids(h_cent.data(), tune);
        return;
    }

    // ---- initial centroids: rows perm[0..k) of the (sampled) training set
    std::vector<int> perm;
    rand_perm(perm, nx, seed + 1);
    d_rows.ensure(k);
    cent.ensure((size_t)k * d);
    CUDA_CHECK(cudaMemcpyAsync(d_rows.p, perm.data(), k * sizeof(int), cudaMemcpyHostToDevice, s));
    km_gather_kernel<<<(unsigned)((k * 32 + 255) / 256), 256, 0, s>>>(x.p, d, d_rows.p, k, cent.p);
    const bool spherical = ix.metric == METRIC_IP;  // IndexIVF.cpp:160-162
    if (spherical) km_renorm_kernel<<<(unsigned)k, 128, 0, s>>>(cent.p, d, k);

    d_best.ensure(nx);
    d_assign.ensure(nx);
    d_hassign.ensure(k);
    ix.centroids.ensure((size_t)k * ix.dpad);
    DevBuf<float> xpad;
    const float* xs = x.p;
    if (ix.dpad != d) {
        xpad.ensure((size_t)nx * ix.dpad);
        launch_pad_rows(x.p, nx, d, xpad.p, ix.dpad, s);
        xs = xpad.p;
    }
    std::vector<unsigned long long> hassign(k);
    std::vector<int2> ops;
    for (int it = 0; it < niter; it++) {
        // index.search(nx, x, 1, dis, assign), Clustering.cpp:192
        launch_pad_rows(cent.p, k, d, ix.centroids.p, ix.dpad, s);
        const long chunk = 65535L * 64;
        for (long j0 = 0; j0 < nx; j0 += chunk) {
            const long m = std::min(chunk, nx - j0);
            launch_coarse_distances(ix.metric, xs + (size_t)j0 * ix.dpad, m, ix.centroids.p, k, ix.dpad, nullptr,
                                    d_best.p + j0, s);
        }
        km_extract_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, s>>>(d_best.p, nx, d_assign.p);
        // km_update_centroids, utils.cpp:1078-1119
        km_accumulate_kernel<<<(unsigned)((k * 32 + 127) / 128), 128, 0, s>>>(x.p, nx, d, d_assign.p, k, cent.p,
                                                                             d_hassign.p);
        CUDA_CHECK(cudaMemcpyAsync(hassign.data(), d_hassign.p, k * sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        // void clusters (utils.cpp:1121-1157): which cluster gets split is decided on the sizes alone
        ops.clear();
        std::mt19937 mt(1234);
        for (long ci = 0; ci < k; ci++) {
            if (hassign[ci] != 0) continue;
            long cj;
            for (cj = 0; true; cj = (cj + 1) % k) {
                const float p = (float)(((double)hassign[cj] - 1.0) / (double)(float)(nx - k));
                const float r = mt() / float(mt.max());
                if (r < p) break;
            }
            ops.push_back(make_int2((int)ci, (int)cj));
            hassign[ci] = hassign[cj] / 2;
            hassign[cj] -= hassign[ci];
        }
        if (!ops.empty()) {
            d_ops.ensure(ops.size());
            CUDA_CHECK(cudaMemcpyAsync(d_ops.p, ops.data(), ops.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
            km_split_kernel<<<1, 256, 0, s>>>(cent.p, d, d_ops.p, (int)ops.size());
        }
        if (spherical) km_renorm_kernel<<<(unsigned)k, 128, 0, s>>>(cent.p, d, k);
        CUDA_CHECK(cudaGetLastError());
    }
    CUDA_CHECK(cudaMemcpyAsync(h_cent.data(), cent.p, (size_t)k * d * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    ix.set_centroids(h_cent.data(), tune);
}

}  // namespace auncel
