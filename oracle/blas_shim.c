/*
 * TEST INFRASTRUCTURE ONLY (oracle/): BLAS/LAPACK link shim for the compiled
 * reference (oracle/_ref).  Not part of the product.
 *
 * The reference's hot path needs exactly one BLAS routine, sgemm_
 * (/root/reference/Auncel/utils.cpp:521,574).  If an OpenBLAS shared object is
 * handed to ref_shim_set_blas() it is dlopen()ed and used; otherwise a plain
 * triple loop (column-major, Fortran calling convention) is used.  sgeqrf_ and
 * sorgqr_ are referenced by utils.cpp's matrix_qr, never called on this path.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

typedef int (*sgemm_fn)(const char*, const char*, const int*, const int*,
                        const int*, const float*, const float*, const int*,
                        const float*, const int*, const float*, float*,
                        const int*);

static sgemm_fn g_sgemm = NULL;
static void* g_blas = NULL;

int ref_shim_set_blas(const char* path) {
    if (!path || !*path) { g_sgemm = NULL; return 0; }
    g_blas = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!g_blas) return -1;
    g_sgemm = (sgemm_fn)dlsym(g_blas, "sgemm_");
    return g_sgemm ? 1 : -2;
}

int ref_shim_has_blas(void) { return g_sgemm != NULL; }

static int is_t(const char* s) { return s[0] == 'T' || s[0] == 't'; }

int sgemm_(const char* transa, const char* transb, const int* m, const int* n,
           const int* k, const float* alpha, const float* a, const int* lda,
           const float* b, const int* ldb, const float* beta, float* c,
           const int* ldc) {
    if (g_sgemm)
        return g_sgemm(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
    const int ta = is_t(transa), tb = is_t(transb);
    const int M = *m, N = *n, K = *k;
#pragma omp parallel for
    for (int j = 0; j < N; j++) {
        for (int i = 0; i < M; i++) {
            float acc = 0.f;
            for (int l = 0; l < K; l++) {
                float av = ta ? a[l + (size_t)i * *lda] : a[i + (size_t)l * *lda];
                float bv = tb ? b[j + (size_t)l * *ldb] : b[l + (size_t)j * *ldb];
                acc += av * bv;
            }
            float prev = (*beta == 0.f) ? 0.f : *beta * c[i + (size_t)j * *ldc];
            c[i + (size_t)j * *ldc] = *alpha * acc + prev;
        }
    }
    return 0;
}

int sgeqrf_(void) { fprintf(stderr, "oracle blas_shim: sgeqrf_ not available\n"); abort(); }
int sorgqr_(void) { fprintf(stderr, "oracle blas_shim: sorgqr_ not available\n"); abort(); }
