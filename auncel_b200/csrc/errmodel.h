// Auncel's error-estimation arithmetic (the phi-U map), written once for host and device.
//
// Every function states the reference lines it reproduces (paths relative to
// /root/reference/Auncel).  The whole library is compiled with -fmad=false and these
// helpers additionally use explicit single-rounding intrinsics on the device, because
// the termination decision truncates into a 500-entry LUT and compares against bucket
// edges: one ulp moves my_nprobe by a whole multiple (SURVEY.md §7 "hard parts" 1).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define AHD __host__ __device__ __forceinline__
#else
#define AHD inline
#endif

namespace auncel {

#if defined(__CUDA_ARCH__)
AHD float fadd(float a, float b) { return __fadd_rn(a, b); }
AHD float fsub(float a, float b) { return __fsub_rn(a, b); }
AHD float fmul(float a, float b) { return __fmul_rn(a, b); }
AHD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
AHD double dadd(double a, double b) { return __dadd_rn(a, b); }
AHD double dsub(double a, double b) { return __dsub_rn(a, b); }
AHD double dmul(double a, double b) { return __dmul_rn(a, b); }
#else
AHD float fadd(float a, float b) { volatile float r = a + b; return r; }
AHD float fsub(float a, float b) { volatile float r = a - b; return r; }
AHD float fmul(float a, float b) { volatile float r = a * b; return r; }
AHD float fdiv(float a, float b) { volatile float r = a / b; return r; }
AHD double dadd(double a, double b) { volatile double r = a + b; return r; }
AHD double dsub(double a, double b) { volatile double r = a - b; return r; }
AHD double dmul(double a, double b) { volatile double r = a * b; return r; }
#endif

enum : int { METRIC_IP = 0, METRIC_L2 = 1 };  // Index.h:48-51

// error bits reported instead of the reference's exceptions / out-of-bounds read
enum : int {
    ERR_ARCOS_DOMAIN = 1,    // IVF_pro.cpp:180 throws outside [-1,1]
    ERR_ARCOS_EDGE = 2,      // x == 1 indexes arcos_list[500] of 500 (IVF_pro.cpp:182)
    ERR_COSINE_PRECOND = 4,  // IVF_pro.cpp:42 throws when a > b
    ERR_TIE_SPAN = 8,        // (engine) a run of equal centroid distances spans a round boundary and the stop stage
};

// error_pro::arcos, IVF_pro.cpp:179-184:  int index = x*arcos_size/2 + arcos_size/2;
// (float*size_t -> float, /2 -> float, + size_t(250) -> float, truncation).  Outside the
// domain the reference throws; here the nearest table end is returned and *err is set.
AHD float arcos_lookup(const float* tab, int size, float x, int* err) {
    if (!(x <= 1.f && x >= -1.f)) {
        *err |= ERR_ARCOS_DOMAIN;
        if (!(x == x)) return tab[size / 2];
        return x > 1.f ? tab[size - 1] : tab[0];
    }
    float f = fmul(x, (float)size);
    f = fdiv(f, 2.f);
    f = fadd(f, (float)(size / 2));
    int index = (int)f;
    if (index >= size) {
        *err |= ERR_ARCOS_EDGE;
        index = size - 1;
    }
    return tab[index];
}

// cosine_theorem, IVF_pro.cpp:41-51.  pow(float,2) promotes to double (exact squares),
// two double additions, one rounding to float, then float / and -.
AHD float cosine_theorem(float a, float b, float c, int* err) {
    if (!(a <= b)) *err |= ERR_COSINE_PRECOND;
    double t = dadd(dmul((double)a, (double)a), dmul((double)c, (double)c));
    t = dsub(t, dmul((double)b, (double)b));
    float temp = (float)t;
    temp = fdiv(temp, fmul(2.f, c));
    return fsub(fdiv(c, 2.f), temp);
}

// packed strict-upper-triangle index of interdis_cem, IVF_pro.cpp:25,35,219
AHD size_t tri_index(size_t nlist, size_t a, size_t b) {
    size_t i = a < b ? a : b, j = a < b ? b : a;
    return (2 * nlist - 1 - i) * i / 2 + j - 1 - i;
}

// error_pro::sum_angle, IVF_pro.cpp:162-177 (n = 15 at every call site)
AHD float sum_angle(float kdis, const float* dtb, int n, int start, const float* tab, int size,
                    int* err) {
    float sum = 0.f;
    for (int i = start; i < start + n; i++) {
        float b = dtb[i];
        if (b >= kdis) continue;
        sum = fadd(sum, arcos_lookup(tab, size, fdiv(b, kdis), err));
    }
    return sum;
}

// Trace::search, IVF_pro.cpp:84-107.  n ascending (phi, U) buckets with sigma.
AHD float trace_search(const float* phi, const float* U, const float* sg, long n, float k,
                       float std_m) {
    if (k <= phi[0]) return fadd(U[0], fmul(std_m, sg[0]));
    if (k >= phi[n - 1]) {
        float ampli = fdiv(k, phi[n - 1]);
        return fmul(fadd(U[n - 1], fmul(std_m, sg[n - 1])), ampli);
    }
    long high = n - 1, low = 0;
    while (low <= high) {
        long middle = (low + high) / 2;
        if (phi[middle] < k)
            low = middle + 1;
        else
            high = middle - 1;
    }
    if (phi[low] > k) low--;
    return fadd(U[low], fmul(std_m, sg[low]));
}

struct ErrModelView {
    const float* arcos;  // arcos_size entries
    int arcos_size;
    int n_traces;
    const long* trace_off;  // n_traces + 1
    const float* phi;
    const float* U;
    const float* sigma;
    float std_m;
    float multipler;
};

AHD float model_U(const ErrModelView& m, int ind, float phi) {
    long o = m.trace_off[ind];
    return trace_search(m.phi + o, m.U + o, m.sigma + o, m.trace_off[ind + 1] - o, phi, m.std_m);
}

// error_pro::cur_num, IVF_pro.cpp:258-291.  D: ascending top-max_topk (IP: angles).
AHD unsigned cur_num(const ErrModelView& m, const float* D, const float* dtb, int index,
                     unsigned query_k, int* err) {
    int start = (1 << index) - 1;
    long high = (long)query_k - 1, low = 0;
    {
        float u = model_U(m, index, sum_angle(D[high], dtb, 15, start, m.arcos, m.arcos_size, err));
        // size_t*float -> float; size_t*1.005 -> double
        if ((double)fmul((float)query_k, u) <= dmul((double)query_k, 1.005)) return query_k;
    }
    while (low <= high) {
        long middle = (low + high) / 2;
        if (middle <= 0) return 0;
        float u = model_U(m, index, sum_angle(D[middle], dtb, 15, start, m.arcos, m.arcos_size, err));
        if (fmul((float)(middle + 1), u) <= (float)query_k)
            low = middle + 1;
        else
            high = middle - 1;
    }
    return (unsigned)(low + 1);
}

// stage -> trace index, IndexIVF.cpp:554-559
AHD int stage_to_ind(size_t stage, size_t nlist) {
    size_t tmp_stage = (stage >= nlist / 8 ? nlist / 8 - 1 : stage);
    int ind = 0;
    while (tmp_stage > ((size_t)1 << ind)) ind++;
    return ind;
}

// kscaling, IVF_pro.cpp:72-82 (fabs(float) stays float; the 1e-5 literals are double)
AHD float kscaling(float kdis, size_t in, const float* gt, size_t max_topk) {
    size_t index = 0;
    for (; index < max_topk; index++) {
        float df = fabsf(fsub(gt[index], kdis));
        if ((double)fdiv(df, kdis) < 1e-5 || (double)df < 1e-5) break;
    }
    if (index >= max_topk) return -1.f;
    return fdiv((float)(index + 1), (float)(in + 1));
}

// order-preserving float <-> uint32 (ascending)
AHD uint32_t f2ord(float f) {
#if defined(__CUDA_ARCH__)
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
AHD float ord2f(uint32_t o) {
    uint32_t u = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

}  // namespace auncel
