/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of Auncel's error-bounded
 * IVF-Flat query path.  Plain C, scalar, single-threaded, no intrinsics.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the
 * product (auncel_b200/) never does.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the
 * compiled, unmodified reference (oracle/_ref/libauncel_ref.so, built by
 * oracle/Makefile from /root/reference/Auncel) by tests/test_oracle_vs_ref.py (runs
 * where the reference tree exists) and against the committed fixtures in
 * tests/golden/ (generated from the reference by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/Auncel).  Build with -ffp-contract=off: the reference's default
 * build (-msse4, no FMA) rounds every multiply and add separately.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_IP 0 /* METRIC_INNER_PRODUCT, Index.h:49 */
#define ORC_L2 1 /* METRIC_L2,            Index.h:50 */

typedef long idx_t; /* Index.h:67 */

/* provided by oracle_sort.cpp: std::sort with the reference's comparators */
void orc_std_sort_pairs_desc_first(float* pairs, long n);
void orc_std_sort_floats(float* v, long n);

/* ------------------------------------------------------------------ distances */

/* utils_simd.cpp:391-416 (SSE build): four lane accumulators, lane l sums the
 * elements i == l (mod 4) in order, a zero-padded masked tail, then two hadds:
 * (s0+s1)+(s2+s3). */
float orc_fvec_L2sqr(const float* x, const float* y, long d) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    long i = 0;
    for (; i + 4 <= d; i += 4)
        for (int l = 0; l < 4; l++) {
            float t = x[i + l] - y[i + l];
            float m = t * t;
            s[l] = s[l] + m;
        }
    if (i < d) { /* masked_read: missing lanes are 0 -> (0-0)^2 added */
        for (int l = 0; l < 4; l++) {
            float xv = (i + l < d) ? x[i + l] : 0.f, yv = (i + l < d) ? y[i + l] : 0.f;
            float t = xv - yv;
            float m = t * t;
            s[l] = s[l] + m;
        }
    }
    float a = s[0] + s[1], b = s[2] + s[3];
    return a + b;
}

/* utils_simd.cpp:419-443 (SSE build). The tail block is executed unconditionally
 * (masked_read(0) gives zeros). */
float orc_fvec_inner_product(const float* x, const float* y, long d) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    long i = 0;
    for (; i + 4 <= d; i += 4)
        for (int l = 0; l < 4; l++) {
            float m = x[i + l] * y[i + l];
            s[l] = s[l] + m;
        }
    for (int l = 0; l < 4; l++) {
        float xv = (i + l < d) ? x[i + l] : 0.f, yv = (i + l < d) ? y[i + l] : 0.f;
        float m = xv * yv;
        s[l] = s[l] + m;
    }
    float a = s[0] + s[1], b = s[2] + s[3];
    return a + b;
}

/* ------------------------------------------------------------------ heaps */

/* Heap.h:51-78: CMin::cmp(a,b) = a<b (min-heap, used for IP), CMax::cmp = a>b
 * (max-heap, used for L2).  `mx` selects CMax. */
static inline int hcmp(int mx, float a, float b) { return mx ? (a > b) : (a < b); }
static inline float hneutral(int mx) { return mx ? FLT_MAX : -FLT_MAX; }

/* Heap.h:88-117 */
static void heap_pop(int mx, size_t k, float* bh_val, idx_t* bh_ids) {
    bh_val--;
    bh_ids--;
    float val = bh_val[k];
    size_t i = 1, i1, i2;
    while (1) {
        i1 = i << 1;
        i2 = i1 + 1;
        if (i1 > k) break;
        if (i2 == k + 1 || hcmp(mx, bh_val[i1], bh_val[i2])) {
            if (hcmp(mx, val, bh_val[i1])) break;
            bh_val[i] = bh_val[i1];
            bh_ids[i] = bh_ids[i1];
            i = i1;
        } else {
            if (hcmp(mx, val, bh_val[i2])) break;
            bh_val[i] = bh_val[i2];
            bh_ids[i] = bh_ids[i2];
            i = i2;
        }
    }
    bh_val[i] = bh_val[k];
    bh_ids[i] = bh_ids[k];
}

/* Heap.h:124-142 */
static void heap_push(int mx, size_t k, float* bh_val, idx_t* bh_ids, float val, idx_t ids) {
    bh_val--;
    bh_ids--;
    size_t i = k, i_father;
    while (i > 1) {
        i_father = i >> 1;
        if (!hcmp(mx, val, bh_val[i_father])) break;
        bh_val[i] = bh_val[i_father];
        bh_ids[i] = bh_ids[i_father];
        i = i_father;
    }
    bh_val[i] = val;
    bh_ids[i] = ids;
}

/* Heap.h:184-208 with k0 = 0 */
static void heap_heapify(int mx, size_t k, float* bh_val, idx_t* bh_ids) {
    for (size_t i = 0; i < k; i++) {
        bh_val[i] = hneutral(mx);
        bh_ids[i] = -1;
    }
}

/* Heap.h:295-322 */
static size_t heap_reorder(int mx, size_t k, float* bh_val, idx_t* bh_ids) {
    size_t i, ii;
    for (i = 0, ii = 0; i < k; i++) {
        float val = bh_val[0];
        idx_t id = bh_ids[0];
        heap_pop(mx, k - i, bh_val, bh_ids);
        bh_val[k - ii - 1] = val;
        bh_ids[k - ii - 1] = id;
        if (id != -1) ii++;
    }
    size_t nel = ii;
    memmove(bh_val, bh_val + k - ii, ii * sizeof(*bh_val));
    memmove(bh_ids, bh_ids + k - ii, ii * sizeof(*bh_ids));
    for (; ii < k; ii++) {
        bh_val[ii] = hneutral(mx);
        bh_ids[ii] = -1;
    }
    return nel;
}

/* ------------------------------------------------------------------ coarse */

/* knn_L2sqr_sse / knn_inner_product_sse (utils.cpp:417-490): exact per-pair
 * distances, heap of size k, strict comparison, heap_reorder => best first.
 * This is what IndexFlat::search (IndexFlat.cpp:42-56) runs for nx < 20 (or when
 * faiss::distance_compute_blas_threshold is raised above nx, utils.cpp:622). */
void orc_coarse(int metric, long n, const float* x, long ny, const float* y, int d, long k,
                float* dis, idx_t* keys) {
    int mx = metric == ORC_L2;
    for (long i = 0; i < n; i++) {
        float* simi = dis + i * k;
        idx_t* idxi = keys + i * k;
        heap_heapify(mx, k, simi, idxi);
        for (long j = 0; j < ny; j++) {
            float v = mx ? orc_fvec_L2sqr(x + i * d, y + j * d, d)
                         : orc_fvec_inner_product(x + i * d, y + j * d, d);
            if (mx ? (v < simi[0]) : (v > simi[0])) {
                heap_pop(mx, k, simi, idxi);
                heap_push(mx, k, simi, idxi, v, j);
            }
        }
        heap_reorder(mx, k, simi, idxi);
    }
}

/* ------------------------------------------------------------------ error_pro */

/* IVF_pro.cpp:21-39 + IndexIVF.cpp:97-109: packed strict upper triangle,
 * (i<j) at (2*num-1-i)*i/2 + j-1-i.  L2: squared distance. IP: acos(c_i . c_j);
 * the reference's "normalisation" loop rescales centroid 0 nlist times
 * (IndexIVF.cpp:102-107, `st` never advances) -- restated literally. */
void orc_interdis(int metric, long num, int d, const float* centroids, float* ret) {
    float* c = (float*)malloc(sizeof(float) * num * d);
    memcpy(c, centroids, sizeof(float) * num * d);
    if (metric == ORC_IP) {
        for (long i = 0; i < num; i++) {
            /* fvec_norm_L2sqr (utils_simd.cpp:137-155) == SSE inner product with itself */
            float nr = sqrtf(orc_fvec_inner_product(c, c, d));
            for (int j = 0; j < d; j++) c[j] /= nr;
        }
    }
    for (long i = 0; i < num; i++)
        for (long j = i + 1; j < num; j++) {
            float v = metric == ORC_IP ? orc_fvec_inner_product(c + i * d, c + j * d, d)
                                       : orc_fvec_L2sqr(c + i * d, c + j * d, d);
            if (metric == ORC_IP) v = acosf(v);
            ret[(2 * num - 1 - i) * i / 2 + j - 1 - i] = v;
        }
    free(c);
}

/* IVF_pro.cpp:151-160 */
void orc_construct_arcos(int len, float* out) {
    float sc = len / 2;
    for (int i = 0; i < len; i++) {
        float xv = (float)(i - sc) / sc;
        out[i] = acosf(xv);
    }
}

/* IVF_pro.cpp:179-184.  The reference throws outside [-1,1] and reads one past the
 * table for x == 1 (index 500 of 500).  Restated as: *err |= 1 on a domain error
 * (result of slot 0/size-1 by clamping), *err |= 2 when the index is clamped from
 * `size` to size-1. */
static float arcos_lookup(const float* tab, int size, float x, int* err) {
    if (!(x <= 1. && x >= -1.)) {
        *err |= 1;
        if (!(x == x)) return tab[size / 2];
        return x > 1.f ? tab[size - 1] : tab[0];
    }
    /* int index = x*arcos_size/2 + arcos_size/2;  (float * size_t -> float) */
    float f = x * (float)size;
    f = f / 2.f;
    f = f + (float)(size / 2);
    int index = (int)f;
    if (index >= size) {
        *err |= 2;
        index = size - 1;
    }
    return tab[index];
}

float orc_arcos(const float* tab, int size, float x, int* err) {
    return arcos_lookup(tab, size, x, err);
}

/* IVF_pro.cpp:41-51; pow(float,int) promotes to double in C++11; squares of floats
 * are exact in double, the two additions round in double, then one rounding to
 * float. *err |= 4 when the precondition a <= b fails (reference throws). */
float orc_cosine_theorem(float a, float b, float c, int* err) {
    if (!(a <= b)) *err |= 4;
    double t = (double)a * (double)a + (double)c * (double)c;
    t = t - (double)b * (double)b;
    float temp = (float)t;
    temp = temp / (2 * c);
    return c / 2 - temp;
}

/* IVF_pro.cpp:196-238 (set_online).  cd/ci: the query's ranked centroid distances
 * and ids (at least nlist/8+21 entries).  Outputs max_num = nlist/8+20 entries each;
 * dtb[max_num-1] is left 0 like the value-initialised vector in the reference. */
void orc_set_online(int metric, long nlist, const float* cd, const idx_t* ci,
                    const float* interdis, const float* arcos, int arcos_size, float* cenTocen,
                    float* dtb, int* err) {
    long max_num = nlist / 8 + 20;
    long cur_cen = ci[0];
    float* cend = (float*)calloc(max_num, sizeof(float));
    if (metric == ORC_IP)
        for (long i = 0; i < max_num; i++) cend[i] = arcos_lookup(arcos, arcos_size, cd[i], err);
    for (long k = 1; k <= max_num; k++) {
        long dst = ci[k];
        long i = cur_cen < dst ? cur_cen : dst;
        long j = cur_cen < dst ? dst : cur_cen;
        cenTocen[k - 1] = interdis[(2 * nlist - 1 - i) * i / 2 + j - 1 - i];
    }
    for (long k = 0; k < max_num; k++) dtb[k] = 0.f;
    for (long k = 0; k < max_num - 1; k++) {
        if (metric == ORC_L2)
            dtb[k] = orc_cosine_theorem(cd[0], cd[k + 1], cenTocen[k], err);
        else
            dtb[k] = orc_cosine_theorem(cend[0], cend[k + 1], cenTocen[k], err);
    }
    free(cend);
}

/* IVF_pro.cpp:162-177 */
float orc_sum_angle(float kdis, const float* dtb, long n, long start, const float* arcos,
                    int arcos_size, int* err) {
    float sum = 0;
    long end = start + n;
    for (long i = start; i < end; i++) {
        if (dtb[i] >= kdis) continue;
        float angle = arcos_lookup(arcos, arcos_size, dtb[i] / kdis, err);
        sum += angle;
    }
    return sum;
}

/* IVF_pro.cpp:84-107 (Trace::search); trace = n ascending (phi, U) buckets + stds */
float orc_trace_search(const float* phi, const float* U, const float* stds, long n, float k,
                       float std_m) {
    float sc = std_m;
    if (k <= phi[0]) return U[0] + sc * stds[0];
    if (k >= phi[n - 1]) {
        float ampli = k / phi[n - 1];
        return (U[n - 1] + sc * stds[n - 1]) * ampli;
    }
    size_t high = n - 1, low = 0, middle = 0;
    while (low <= high) {
        middle = (low + high) / 2;
        if (phi[middle] < k)
            low = middle + 1;
        else
            high = middle - 1;
    }
    if (phi[low] > k) low--;
    return U[low] + sc * stds[low];
}

/* IVF_pro.cpp:109-149 (Trace::SB).  pairs: n interleaved (phi, U) floats, unset
 * entries are (-1,-1).  std::sort with the reference's comparator decides the
 * permutation among equal phi (oracle_sort.cpp).  Returns the bucket count sz and
 * writes ascending phi_out/U_out/std_out (capacity >= (n+bs-1)/bs). */
long orc_trace_SB(float* pairs, long n, long bs, float* phi_out, float* U_out, float* std_out) {
    orc_std_sort_pairs_desc_first(pairs, n);
    long size = 0;
    for (long i = 0; i < n; i++) size += (pairs[2 * i] < 0 && pairs[2 * i + 1] < 0) ? 0 : 1;
    long sz = (size + bs - 1) / bs;
    for (long i = 0; i < sz; i++) {
        long left = i * bs, right = (i + 1) * bs;
        if (right > size) right = size;
        float ave1 = 0, ave2 = 0;
        for (long index = left; index < right; index++) {
            long j = index - left;
            ave1 = (float)j / (float)(j + 1) * ave1 + pairs[2 * index] / (j + 1);
            ave2 = (float)j / (float)(j + 1) * ave2 + pairs[2 * index + 1] / (j + 1);
        }
        double accum = 0.;
        for (long index = left; index < right; index++) {
            /* (d.second-ave2)*(d.second-ave2): float arithmetic, then += into double */
            float df = pairs[2 * index + 1] - ave2;
            float sq = df * df;
            accum += sq;
        }
        float sd = (float)sqrt(accum / bs);
        /* reversed to ascending order (IVF_pro.cpp:146-148) */
        phi_out[sz - 1 - i] = ave1;
        U_out[sz - 1 - i] = ave2;
        std_out[sz - 1 - i] = sd;
    }
    return sz;
}

/* IVF_pro.cpp:72-82 */
float orc_kscaling(float kdis, long in, const float* gt, long max_topk) {
    long index = 0;
    for (; index < max_topk; index++) {
        if (fabsf(gt[index] - kdis) / kdis < 1e-5 || fabsf(gt[index] - kdis) < 1e-5) break;
    }
    if (index >= max_topk) return -1;
    return (index + 1) / (float)(in + 1);
}

typedef struct {
    const float* arcos;
    int arcos_size;
    int n_traces;
    const long* trace_off; /* n_traces+1 */
    const float *phi, *U, *sigma;
    float std_m;
} orc_model;

static float model_U(const orc_model* m, long ind, float phi) {
    long o = m->trace_off[ind], n = m->trace_off[ind + 1] - o;
    return orc_trace_search(m->phi + o, m->U + o, m->sigma + o, n, phi, m->std_m);
}

/* IVF_pro.cpp:258-291 (cur_num) */
static size_t cur_num(const orc_model* m, const float* D, const float* dtb, size_t index,
                      size_t query_k, int* err) {
    size_t nprobe = (size_t)1 << index;
    size_t high = query_k - 1, low = 0, middle = 0;
    if (query_k * model_U(m, index,
                          orc_sum_angle(D[high], dtb, 15, nprobe - 1, m->arcos, m->arcos_size, err)) <=
        query_k * 1.005)
        return query_k;
    while (low <= high) {
        middle = (low + high) / 2;
        if (middle <= 0) return 0;
        if ((middle + 1) * model_U(m, index,
                                   orc_sum_angle(D[middle], dtb, 15, nprobe - 1, m->arcos,
                                                 m->arcos_size, err)) <=
            query_k) {
            low = middle + 1;
        } else {
            high = middle - 1;
        }
    }
    return low + 1;
}

size_t orc_cur_num(const float* arcos, int arcos_size, int n_traces, const long* trace_off,
                   const float* phi, const float* U, const float* sigma, float std_m,
                   const float* D, const float* dtb, long index, long query_k, int* err) {
    orc_model m = {arcos, arcos_size, n_traces, trace_off, phi, U, sigma, std_m};
    return cur_num(&m, D, dtb, index, query_k, err);
}

/* ------------------------------------------------------------------ the query loop */

/* IVFFlatScanner::scan_codes (IndexIVFFlat.cpp:117-137) */
static size_t scan_codes(int metric, int d, const float* xi, size_t list_size, const float* vecs,
                         const idx_t* ids, float* simi, idx_t* idxi, size_t k) {
    int mx = metric == ORC_L2;
    size_t nup = 0;
    for (size_t j = 0; j < list_size; j++) {
        const float* yj = vecs + (size_t)d * j;
        float dis = mx ? orc_fvec_L2sqr(xi, yj, d) : orc_fvec_inner_product(xi, yj, d);
        if (hcmp(mx, simi[0], dis)) {
            heap_pop(mx, k, simi, idxi);
            heap_push(mx, k, simi, idxi, dis, ids[j]);
            nup++;
        }
    }
    return nup;
}

/*
 * IndexIVF::search_preassigned (IndexIVF.cpp:382-736) for one thread.
 *   mode 0: plain fixed-nprobe search        (tune=false, training=false)
 *   mode 1: Auncel error-bounded search      (tune block, :551-638)
 *   mode 2: calibration                      (training block, :640-673)
 * Inverted lists are given CSR-style: list l = rows [list_off[l], list_off[l+1]) of
 * `codes` (row-major, d floats) and `ids`, in insertion order (InvertedLists.cpp:138-196).
 * keys/coarse_dis: n x nprobe from the coarse quantizer.  `offset` is the global id of
 * query 0 (the k>>32 packing, :389-392); require_acc / my_nprobe / t_recalls / gt_D /
 * train pairs are all indexed by the GLOBAL id i+offset like the reference.
 * train_pairs: n_traces arrays of train_num*(k/4) (phi,U) float pairs, concatenated.
 * dump_q >= 0: per-stage diagnostics for local query dump_q into dump[nprobe*4]:
 *   (pre_num, recall_after_plateau, ext, my_nprobe_after_stage); untouched stages keep -2.
 * Returns an error bitmask (0 = clean): see arcos_lookup / orc_cosine_theorem.
 */
/* error_pro::time_tune (IVF_pro.h:82-114): the latency-budget cut of Error_sys::time_search
 * (profile.cpp:229-244, IndexIVF.cpp:545-549).  The reference reads the wall clock
 * (IndexIVF::time(), :329-333: gettimeofday -> tv_sec + tv_usec * 1e-6); here the clock is a
 * model: every probe iteration that reaches the check costs us_per_list microseconds plus
 * ns_per_code nanoseconds per scanned code, and time() is evaluated on the resulting
 * (tv_sec, tv_usec) pair with the reference's arithmetic.  require_acc[id] is the budget in ms. */
static int g_time_tune = 0;
static long g_us_per_list = 0, g_ns_per_code = 0;
void orc_set_time_tune(int on, long us_per_list, long ns_per_code) {
    g_time_tune = on;
    g_us_per_list = us_per_list;
    g_ns_per_code = ns_per_code;
}
static double orc_vclock(unsigned long long total_ns) {
    unsigned long long us = total_ns / 1000ull;
    return (double)(us / 1000000ull) + (double)(us % 1000000ull) * 1e-6;
}

int orc_search_preassigned(int metric, int d, long nlist, const float* codes, const long* list_off,
                           const idx_t* ids, long n, const float* x, long k, long nprobe,
                           long max_codes, const idx_t* keys, const float* coarse_dis, int mode,
                           long offset, const float* interdis, const float* arcos, int arcos_size,
                           int n_traces, const long* trace_off, const float* tr_phi,
                           const float* tr_U, const float* tr_sigma, float multipler, float std_m,
                           long query_topk, const float* require_acc, const float* gt_D,
                           int profile, int overhead_profile, unsigned long* my_nprobe,
                           float* t_recalls, float* train_pairs, long train_num, float* D,
                           idx_t* I, long* stats, long dump_q, float* dump) {
    int err = 0;
    int mx = metric == ORC_L2;
    size_t nlistv = 0, ndis = 0, nheap = 0;
    long max_num = nlist / 8 + 20;
    orc_model m = {arcos, arcos_size, n_traces, trace_off, tr_phi, tr_U, tr_sigma, std_m};
    float* dtb = (float*)calloc(max_num, sizeof(float));
    float* c2c = (float*)calloc(max_num, sizeof(float));
    float* tmp_simi = (float*)malloc(sizeof(float) * k);
    int tune = mode == 1, training = mode == 2;

    for (long i = 0; i < n; i++) {
        long id_q = i + offset;
        const float* xi = x + i * d;
        float* simi = D + i * k;
        idx_t* idxi = I + i * k;
        heap_heapify(mx, k, simi, idxi); /* init_result, :421-427 */
        long nscan = 0;
        unsigned long long vt_ns = 0;
        double t0 = orc_vclock(0); /* :504-506 */
        size_t pre_num = 0, query_k = 0, stoped = 0;
        float true_KD_K = 0, pre_val = 0;
        if (tune) {
            query_k = query_topk;
            if (gt_D) true_KD_K = gt_D[id_q * k + query_k - 1]; /* :509 */
        }
        if (tune || training) /* :512-523 */
            orc_set_online(metric, nlist, coarse_dis + i * nprobe, keys + i * nprobe, interdis,
                           arcos, arcos_size, c2c, dtb, &err);

        for (long ik = 0; ik < nprobe; ik++) { /* :526 */
            /* scan_one_list, :439-475 */
            idx_t key = keys[i * nprobe + ik];
            vt_ns += (unsigned long long)g_us_per_list * 1000ull;
            if (key >= 0) {
                size_t ls = list_off[key + 1] - list_off[key];
                vt_ns += (unsigned long long)ls * (unsigned long long)g_ns_per_code;
                if (ls != 0) {
                    nlistv++;
                    nheap += scan_codes(metric, d, xi, ls, codes + (size_t)list_off[key] * d,
                                        ids + list_off[key], simi, idxi, k);
                    nscan += ls;
                }
            }
            if (max_codes && nscan >= max_codes) break; /* :541-543 */
            if (g_time_tune) { /* :545-549 */
                double now = orc_vclock(vt_ns);
                if ((now - t0) * 1000 >= require_acc[id_q] * 0.95 - (now - t0) * 1000 / (ik + 1)) break;
            }

            if (tune) { /* :551-638 */
                size_t stage = ik + 1;
                size_t ind = 0;
                size_t tmp_stage = (stage >= (size_t)nlist / 8 ? nlist / 8 - 1 : stage);
                while (tmp_stage > ((size_t)1 << ind)) ind++;
                memcpy(tmp_simi, simi, sizeof(float) * k);
                if (!mx)
                    for (long j = 0; j < k; j++)
                        tmp_simi[j] = arcos_lookup(arcos, arcos_size, tmp_simi[j], &err);
                orc_std_sort_floats(tmp_simi, k);
                pre_num = cur_num(&m, tmp_simi, dtb, ind, query_k, &err);
                float recall = pre_num / (float)query_k;
                size_t cnt = 0;
                float max_val = -1;
                size_t stops = (size_t)(require_acc[id_q] * 12);
                if (mx) {
                    for (long j = 0; j < k; j++) {
                        max_val = fmaxf(max_val, simi[j]);
                        if (simi[j] <= true_KD_K * 1.0005) cnt++;
                    }
                } else {
                    max_val = FLT_MAX;
                    for (long j = 0; j < k; j++) {
                        max_val = fminf(max_val, simi[j]);
                        if (simi[j] >= true_KD_K * 0.9995) cnt++;
                    }
                }
                if (stage > 1) {
                    if (max_val == pre_val)
                        stoped++;
                    else
                        stoped = 0;
                    if (stoped >= stops) recall = 1;
                    pre_val = max_val;
                } else {
                    pre_val = max_val;
                }
                float true_recall = cnt / (float)query_k;
                float require_recall = require_acc[id_q];
                int brk = 0;
                if (!overhead_profile) {
                    if (recall >= require_recall && my_nprobe[id_q] == 0) {
                        my_nprobe[id_q] = (unsigned long)(stage * multipler);
                        if (my_nprobe[id_q] >= (unsigned long)nlist) t_recalls[id_q] = 1.;
                    }
                    if (stage >= (size_t)nlist / 8 && my_nprobe[id_q] == 0) {
                        my_nprobe[id_q] = (unsigned long)(stage * multipler);
                        if (my_nprobe[id_q] >= (unsigned long)nlist) t_recalls[id_q] = 1.;
                    }
                    if (profile && my_nprobe[id_q] != 0 && my_nprobe[id_q] <= stage) {
                        t_recalls[id_q] = true_recall;
                        brk = 1;
                    }
                    if (!profile && my_nprobe[id_q] != 0 && my_nprobe[id_q] <= stage) brk = 1;
                } else {
                    if (stage >= (size_t)nlist / 8) brk = 1;
                }
                if (dump && i == dump_q) {
                    dump[ik * 4 + 0] = (float)pre_num;
                    dump[ik * 4 + 1] = recall;
                    dump[ik * 4 + 2] = max_val;
                    dump[ik * 4 + 3] = (float)my_nprobe[id_q];
                }
                if (brk) break;
            }
            if (training) { /* :640-673 */
                size_t stage = ik + 1;
                if (stage > (size_t)nlist / 8) break;
                if ((stage & (stage - 1)) != 0) continue;
                size_t ind = 0;
                while (stage != ((size_t)1 << ind)) ind++;
                memcpy(tmp_simi, simi, sizeof(float) * k);
                orc_std_sort_floats(tmp_simi, k);
                if (!mx) /* std::reverse */
                    for (long a = 0, b = k - 1; a < b; a++, b--) {
                        float t = tmp_simi[a];
                        tmp_simi[a] = tmp_simi[b];
                        tmp_simi[b] = t;
                    }
                long count = 0;
                float* tp = train_pairs + 2 * ((size_t)ind * train_num * (k / 4));
                for (long ij = 0; ij < k; ij++) {
                    float ks = orc_kscaling(tmp_simi[ij], ij, gt_D + id_q * k, k);
                    if (ks < 0) break;
                    float tval = tmp_simi[ij];
                    if (!mx) tval = arcos_lookup(arcos, arcos_size, tval, &err);
                    float sum_a = orc_sum_angle(tval, dtb, 15, stage - 1, arcos, arcos_size, &err);
                    size_t slot = (size_t)id_q * (k / 4) + count++;
                    tp[2 * slot] = sum_a;
                    tp[2 * slot + 1] = ks;
                    if (count >= k / 4) break;
                }
            }
        }
        ndis += nscan;
        heap_reorder(mx, k, simi, idxi); /* :677 */
    }
    if (stats) {
        stats[0] += nlistv;
        stats[1] += ndis;
        stats[2] += nheap;
    }
    free(dtb);
    free(c2c);
    free(tmp_simi);
    return err;
}

/* ------------------------------------------------------------------ range search */

/* IndexIVF::range_search_preassigned (IndexIVF.cpp:760-860, parallel_mode 0) with
 * IVFFlatScanner::scan_codes_range (IndexIVFFlat.cpp:139-155): every vector of the probed lists
 * with C::cmp(radius, dis) -- L2: dis < radius, IP: dis > radius -- in scan order (probe rank, then
 * in-list order).  Two passes like RangeSearchResult (AuxIndexStructures.h:31-50): out_D == NULL
 * only fills lims[n + 1]; stats[0] += lists visited, stats[1] += codes scanned. */
void orc_range_search(int metric, int d, const float* codes, const long* list_off, const idx_t* ids,
                      long n, const float* x, float radius, long nprobe, const idx_t* keys, long* lims,
                      float* out_D, idx_t* out_I, long* stats) {
    int mx = metric == ORC_L2;
    long pos = 0;
    for (long i = 0; i < n; i++) {
        lims[i] = pos;
        for (long ik = 0; ik < nprobe; ik++) {
            idx_t key = keys[i * nprobe + ik];
            if (key < 0) continue;
            long ls = list_off[key + 1] - list_off[key];
            if (ls == 0) continue;
            if (stats && !out_D) {
                stats[0]++;
                stats[1] += ls;
            }
            for (long j = 0; j < ls; j++) {
                const float* yj = codes + (size_t)(list_off[key] + j) * d;
                float dis = mx ? orc_fvec_L2sqr(x + i * d, yj, d) : orc_fvec_inner_product(x + i * d, yj, d);
                if (mx ? (radius > dis) : (radius < dis)) {
                    if (out_D) {
                        out_D[pos] = dis;
                        out_I[pos] = ids[list_off[key] + j];
                    }
                    pos++;
                }
            }
        }
    }
    lims[n] = pos;
}

/* ------------------------------------------------------------------ shards */

/* merge_tables (IndexShards.cpp:44-105): per query a heap over the heads of the
 * nshard sorted result rows; CMin<float,int> for L2, CMax for IP (:303-311);
 * labels < 0 end a shard's row. */
static void sheap_push(int l2, int k, float* v, int* s, float val, int sid) {
    v--;
    s--;
    int i = k, f;
    while (i > 1) {
        f = i >> 1;
        if (!(l2 ? (val < v[f]) : (val > v[f]))) break;
        v[i] = v[f];
        s[i] = s[f];
        i = f;
    }
    v[i] = val;
    s[i] = sid;
}
static void sheap_pop(int l2, int k, float* v, int* s) {
    v--;
    s--;
    float val = v[k];
    int i = 1, i1, i2;
#define SC(a, b) (l2 ? ((a) < (b)) : ((a) > (b)))
    while (1) {
        i1 = i << 1;
        i2 = i1 + 1;
        if (i1 > k) break;
        if (i2 == k + 1 || SC(v[i1], v[i2])) {
            if (SC(val, v[i1])) break;
            v[i] = v[i1];
            s[i] = s[i1];
            i = i1;
        } else {
            if (SC(val, v[i2])) break;
            v[i] = v[i2];
            s[i] = s[i2];
            i = i2;
        }
    }
#undef SC
    v[i] = v[k];
    s[i] = s[k];
}

void orc_merge_tables(int metric, long n, long k, long nshard, float* distances, idx_t* labels,
                      const float* all_distances, const idx_t* all_labels,
                      const long* translations) {
    if (k == 0) return;
    int l2 = metric == ORC_L2;
    long stride = n * k;
    int* pointer = (int*)malloc(sizeof(int) * nshard);
    int* shard_ids = (int*)malloc(sizeof(int) * nshard);
    float* heap_vals = (float*)malloc(sizeof(float) * nshard);
    for (long i = 0; i < n; i++) {
        const float* D_in = all_distances + i * k;
        const idx_t* I_in = all_labels + i * k;
        int heap_size = 0;
        for (long s = 0; s < nshard; s++) {
            pointer[s] = 0;
            if (I_in[stride * s] >= 0) sheap_push(l2, ++heap_size, heap_vals, shard_ids, D_in[stride * s], (int)s);
        }
        float* Dq = distances + i * k;
        idx_t* Iq = labels + i * k;
        for (long j = 0; j < k; j++) {
            if (heap_size == 0) {
                Iq[j] = -1;
                Dq[j] = l2 ? -FLT_MAX : FLT_MAX; /* C::neutral() of CMin / CMax, sic */
            } else {
                int s = shard_ids[0];
                int p = pointer[s];
                Dq[j] = heap_vals[0];
                Iq[j] = I_in[stride * s + p] + translations[s];
                sheap_pop(l2, heap_size--, heap_vals, shard_ids);
                p++;
                pointer[s] = p;
                if (p < k && I_in[stride * s + p] >= 0)
                    sheap_push(l2, ++heap_size, heap_vals, shard_ids, D_in[stride * s + p], s);
            }
        }
    }
    free(pointer);
    free(shard_ids);
    free(heap_vals);
}
