/*
 * TEST INFRASTRUCTURE ONLY (oracle/): a thin extern "C" driver around the
 * UNMODIFIED reference sources under /root/reference/Auncel, compiled where
 * they lie by oracle/Makefile into oracle/_ref/libauncel_ref.so.
 *
 * Purpose: (1) pin oracle/auncel_oracle.c (the CPU restatement) against the real
 * reference, (2) generate tests/golden/ fixtures (tests/golden/make_golden.py),
 * (3) serve as bench.py's `--impl reference` / cpu_baseline arm.
 * Nothing in the product (auncel_b200/) may load this library.
 *
 * The driver only calls the reference's public API:
 *   IndexFlat / IndexIVFFlat          Auncel/IndexFlat.h, Auncel/IndexIVFFlat.h
 *   Error_sys                         Auncel/profile.h:29-91
 *   error_pro / Trace                 Auncel/IVF_pro.h:44-175
 *   IndexShards / IndexReplicas       Auncel/IndexShards.h, IndexReplicas.h
 */
#include <omp.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "FaissException.h"
#include "IVF_pro.h"
#include "IndexFlat.h"
#include "IndexIVFFlat.h"
#include "IndexReplicas.h"
#include "IndexShards.h"
#include "AuxIndexStructures.h"
#include "profile.h"
#include "utils.h"

using faiss::Index;
typedef long idx_t;

namespace {

struct Ref {
    int d = 0;
    long nlist = 0;
    faiss::MetricType metric = faiss::METRIC_L2;
    faiss::IndexFlat* quantizer = nullptr;
    faiss::IndexIVFFlat* index = nullptr;
    faiss::Error_sys* es = nullptr;
    std::vector<float> acc;  // require_acc storage handed to Error_sys::set_queries
    std::vector<float> queries;
};

thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH                                   \
    }                                               \
    catch (const std::exception& e) {               \
        g_err = e.what();                           \
        return -1;                                  \
    }                                               \
    catch (...) {                                   \
        g_err = "unknown exception";                \
        return -1;                                  \
    }                                               \
    return 0;

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_blas_threshold(int v) { faiss::distance_compute_blas_threshold = v; }
int ref_get_blas_threshold() { return faiss::distance_compute_blas_threshold; }
void ref_omp_set_num_threads(int n) { omp_set_num_threads(n); }
int ref_omp_get_max_threads() { return omp_get_max_threads(); }

/* metric: 0 = inner product, 1 = L2  (Auncel/Index.h:48-51) */
void* ref_create(int d, long nlist, int metric) {
    Ref* r = new Ref;
    r->d = d;
    r->nlist = nlist;
    r->metric = metric == 0 ? faiss::METRIC_INNER_PRODUCT : faiss::METRIC_L2;
    if (metric == 0)
        r->quantizer = new faiss::IndexFlatIP(d);
    else
        r->quantizer = new faiss::IndexFlatL2(d);
    r->index = new faiss::IndexIVFFlat(r->quantizer, d, nlist, r->metric);
    return r;
}

void ref_free(void* h) {
    Ref* r = (Ref*)h;
    // Error_sys does not own the index; error_pro (ix->t) is leaked by the
    // reference as well (no owner deletes it) -- keep that behaviour.
    delete r->es;
    delete r->index;
    delete r->quantizer;
    delete r;
}

/* eval/bound.cpp:261-263 : set_tune_mode(); train(); set_tune_off(); */
int ref_train(void* h, long n, const float* x, int niter, int verbose) {
    Ref* r = (Ref*)h;
    REF_TRY
    if (niter > 0) r->index->cp.niter = niter;
    r->index->verbose = verbose != 0;
    r->index->set_tune_mode();
    r->index->train(n, x);
    r->index->set_tune_off();
    REF_CATCH
}

/* Import centroids without k-means, and fill interdis_cem exactly as
 * Level1Quantizer::train_q1 does (Auncel/IndexIVF.cpp:97-109). */
int ref_set_centroids(void* h, const float* centroids) {
    Ref* r = (Ref*)h;
    REF_TRY
    long nlist = r->nlist;
    int d = r->d;
    r->quantizer->reset();
    r->quantizer->add(nlist, centroids);
    r->quantizer->is_trained = true;
    r->index->is_trained = true;
    std::vector<float> c(centroids, centroids + (size_t)nlist * d);
    r->index->interdis_cem.resize((size_t)nlist * (nlist - 1) / 2);
    if (r->metric != faiss::METRIC_INNER_PRODUCT) {
        faiss::fvec_inter_vecs(r->index->interdis_cem.data(), c.data(), nlist, d);
    } else {
        for (long i = 0; i < nlist; i++) {  // sic: centroid 0 every time
            float* st = c.data();
            float norm = sqrtf(faiss::fvec_norm_L2sqr(st, d));
            for (int j = 0; j < d; j++) st[j] /= norm;
        }
        faiss::fvec_inter_vecs_IP(r->index->interdis_cem.data(), c.data(), nlist, d);
        for (size_t i = 0; i < r->index->interdis_cem.size(); i++)
            r->index->interdis_cem[i] = std::acos(r->index->interdis_cem[i]);
    }
    REF_CATCH
}

int ref_get_centroids(void* h, float* out) {
    Ref* r = (Ref*)h;
    memcpy(out, r->quantizer->xb.data(), sizeof(float) * r->nlist * r->d);
    return 0;
}

long ref_interdis_size(void* h) { return (long)((Ref*)h)->index->interdis_cem.size(); }

int ref_get_interdis(void* h, float* out) {
    Ref* r = (Ref*)h;
    memcpy(out, r->index->interdis_cem.data(), sizeof(float) * r->index->interdis_cem.size());
    return 0;
}

/* IndexIVFFlat::add_core (Auncel/IndexIVFFlat.cpp:41-80); list_no may be NULL */
int ref_add(void* h, long n, const float* x, const long* ids, const long* list_no) {
    Ref* r = (Ref*)h;
    REF_TRY
    r->index->add_core(n, x, ids, list_no);
    REF_CATCH
}

long ref_ntotal(void* h) { return ((Ref*)h)->index->ntotal; }

int ref_list_sizes(void* h, long* out) {
    Ref* r = (Ref*)h;
    for (long l = 0; l < r->nlist; l++) out[l] = (long)r->index->invlists->list_size(l);
    return 0;
}

int ref_list_ids(void* h, long list_no, long* out) {
    Ref* r = (Ref*)h;
    size_t n = r->index->invlists->list_size(list_no);
    memcpy(out, r->index->invlists->get_ids(list_no), n * sizeof(long));
    return 0;
}

/* quantizer->assign (Auncel/Index.cpp:42-47) */
int ref_assign(void* h, long n, const float* x, long* out) {
    Ref* r = (Ref*)h;
    REF_TRY
    r->quantizer->assign(n, x, out);
    REF_CATCH
}

/* quantizer->search (Auncel/IndexFlat.cpp:42-56) */
int ref_coarse(void* h, long n, const float* x, long nprobe, float* dis, long* keys) {
    Ref* r = (Ref*)h;
    REF_TRY
    r->quantizer->search(n, x, nprobe, dis, keys);
    REF_CATCH
}

static void ensure_t(Ref* r) {
    // Auncel/IndexIVF.cpp:529 dereferences t unconditionally.
    if (!r->index->t) r->index->init_tune(0, 4, nullptr, nullptr, nullptr, nullptr, nullptr);
}

/* plain IndexIVF::search (Auncel/IndexIVF.cpp:335-353) */
int ref_search_fixed(void* h, long n, const float* x, long k, long nprobe, long max_codes,
                     float* D, long* I) {
    Ref* r = (Ref*)h;
    REF_TRY
    ensure_t(r);
    r->index->nprobe = nprobe;
    r->index->max_codes = max_codes;
    r->index->search(n, x, k, D, I);
    r->index->max_codes = 0;
    REF_CATCH
}

/* ---- Error_sys (Auncel/profile.cpp) ---- */

int ref_es_create(void* h, long nq_total, long max_topk, const float* gtD, const long* gtI) {
    Ref* r = (Ref*)h;
    REF_TRY
    delete r->es;
    r->es = new faiss::Error_sys(r->index, nq_total, max_topk);
    r->es->set_gt(gtD, gtI);
    REF_CATCH
}

/* sys_train writes Validation_<d>_<np>.log into the cwd (profile.cpp:158-169) */
int ref_es_sys_train(void* h, long ts, const float* xq, const char* workdir) {
    Ref* r = (Ref*)h;
    REF_TRY
    char old[4096];
    if (!getcwd(old, sizeof(old))) old[0] = 0;
    if (workdir && *workdir && chdir(workdir) != 0) throw std::runtime_error("chdir failed");
    r->es->sys_train(ts, xq);
    if (old[0] && chdir(old) != 0) throw std::runtime_error("chdir back failed");
    REF_CATCH
}

long ref_n_traces(void* h) { return (long)((Ref*)h)->index->t->traces.size(); }
long ref_trace_size(void* h, long t) { return (long)((Ref*)h)->index->t->traces[t].trace.size(); }

int ref_get_trace(void* h, long t, float* phi, float* U, float* sigma) {
    Ref* r = (Ref*)h;
    const faiss::Trace& tr = r->index->t->traces[t];
    for (size_t i = 0; i < tr.trace.size(); i++) {
        phi[i] = tr.trace[i].first;
        U[i] = tr.trace[i].second;
        if (sigma && i < tr.stds.size()) sigma[i] = tr.stds[i];
    }
    return 0;
}

long ref_trace_nstd(void* h, long t) { return (long)((Ref*)h)->index->t->traces[t].stds.size(); }

/* replace the trained traces (used to make the two sides share one model) */
int ref_set_trace(void* h, long t, long n, const float* phi, const float* U, const float* sigma) {
    Ref* r = (Ref*)h;
    faiss::Trace& tr = r->index->t->traces[t];
    tr.trace.resize(n);
    tr.stds.resize(n);
    for (long i = 0; i < n; i++) {
        tr.trace[i] = std::make_pair(phi[i], U[i]);
        tr.stds[i] = sigma[i];
    }
    return 0;
}

int ref_get_arcos(void* h, float* out) {
    Ref* r = (Ref*)h;
    memcpy(out, r->index->t->arcos_list.data(), sizeof(float) * r->index->t->arcos_list.size());
    return 0;
}

/* set_topk + set_queries (profile.cpp:173-209); acc has `alloc` entries indexed by
 * the GLOBAL query id; xq is the base pointer of all queries (copied). */
int ref_es_set_queries(void* h, long query_topk, long num, const float* xq, long nq_total,
                       const float* acc, long alloc, float multipler, float std_m, int profile,
                       int overhead_profile) {
    Ref* r = (Ref*)h;
    REF_TRY
    r->queries.assign(xq, xq + (size_t)nq_total * r->d);
    r->acc.assign(acc, acc + alloc);
    r->es->set_topk(query_topk);
    r->es->set_queries(num, r->queries.data(), r->acc.data(), alloc);
    r->index->t->multipler = multipler;
    r->index->t->std_m = std_m;
    r->index->t->profile = profile != 0;
    r->index->t->overhead_profile = overhead_profile != 0;
    REF_CATCH
}

/* Deterministic clock for the latency-budget mode.  IndexIVF::time() (Auncel/IndexIVF.cpp:329-333) is
 * gettimeofday(); the library is linked with -Bsymbolic-functions so the reference's calls bind to
 * this definition.  Virtual mode: every call advances the clock by exactly one second (tv_usec = 0),
 * so (now - t0) is an exact small integer and the break rule of :545-549 becomes reproducible;
 * otherwise the real clock is returned (clock_gettime). */
static int g_virtual_clock = 0;
static long g_virtual_now = 1000;
int gettimeofday(struct timeval* tv, void* tz) noexcept {
    (void)tz;
    if (g_virtual_clock) {
        tv->tv_sec = ++g_virtual_now;
        tv->tv_usec = 0;
        return 0;
    }
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    tv->tv_sec = ts.tv_sec;
    tv->tv_usec = ts.tv_nsec / 1000;
    return 0;
}
void ref_set_virtual_clock(int on) { g_virtual_clock = on; }

/* Error_sys::time_search (profile.cpp:229-244): require_acc[id] is the budget in ms.  The
 * reference leaves error_pro::time_tune set afterwards (:242); `reset_flag` clears it again so
 * later searches of the harness are unaffected. */
int ref_es_time_search(void* h, float* D, long* I, long start, long search_size, int reset_flag) {
    Ref* r = (Ref*)h;
    REF_TRY
    r->es->time_search(D, (int64_t*)I, start, (size_t)search_size);
    if (reset_flag) r->index->t->time_tune = false;
    REF_CATCH
}

/* IndexIVF::range_search (IndexIVF.cpp:741-860).  Two calls like RangeSearchResult's own life
 * cycle: the first runs the search and returns lims (nq + 1), the second copies the buffers. */
static faiss::RangeSearchResult* g_range = nullptr;
int ref_range_search(void* h, long n, const float* x, float radius, long nprobe, long* lims) {
    Ref* r = (Ref*)h;
    REF_TRY
    delete g_range;
    g_range = new faiss::RangeSearchResult(n);
    r->index->nprobe = nprobe;
    r->index->range_search(n, x, radius, g_range);
    for (long i = 0; i <= n; i++) lims[i] = (long)g_range->lims[i];
    REF_CATCH
}
int ref_range_results(long total, float* D, long* I) {
    if (!g_range) return -1;
    memcpy(D, g_range->distances, sizeof(float) * total);
    memcpy(I, g_range->labels, sizeof(long) * total);
    delete g_range;
    g_range = nullptr;
    return 0;
}

/* Error_sys::search (profile.cpp:211-227) */
int ref_es_search(void* h, float* D, long* I, long start, long search_size) {
    Ref* r = (Ref*)h;
    REF_TRY
    r->es->search(D, (int64_t*)I, start, (size_t)search_size);
    REF_CATCH
}

int ref_get_my_nprobe(void* h, long start, long n, unsigned long* out) {
    Ref* r = (Ref*)h;
    for (long i = 0; i < n; i++) out[i] = r->index->t->my_nprobe[start + i];
    return 0;
}

int ref_clear_my_nprobe(void* h) {
    Ref* r = (Ref*)h;
    memset(r->index->t->my_nprobe, 0, r->index->t->alloc_s * sizeof(size_t));
    memset(r->index->t->t_recalls, 0, r->index->t->alloc_s * sizeof(float));
    return 0;
}

int ref_get_t_recalls(void* h, long start, long n, float* out) {
    Ref* r = (Ref*)h;
    for (long i = 0; i < n; i++) out[i] = r->index->t->t_recalls[start + i];
    return 0;
}

/* Timed CPU baseline.  The reference's own `#pragma omp for` over queries is
 * malformed (Auncel/IndexIVF.cpp:484-485) so the file only builds without
 * -fopenmp.  This restores the intended query-level parallelism from OUTSIDE the
 * unmodified code: `nthreads` std::threads each run the reference's 6-argument
 * IndexIVF::search on a contiguous query slice (the same static split an `omp
 * for` would do); per-query state (my_nprobe[id], t_recalls[id]) is indexed by
 * global id so slices do not interfere. */
int ref_es_search_threads_chunk(void* h, float* D, long* I, long start, long num, int nthreads, long chunk);
int ref_es_search_threads(void* h, float* D, long* I, long start, long num, int nthreads) {
    return ref_es_search_threads_chunk(h, D, I, start, num, nthreads, 0);
}

/* chunk == 0: static contiguous slices (what `omp for` does by default); chunk > 0: the threads draw
 * `chunk` queries at a time from a shared counter (`schedule(dynamic, chunk)`), so one query with a
 * very large my_nprobe does not set the time of the whole call. */
int ref_es_search_threads_chunk(void* h, float* D, long* I, long start, long num, int nthreads, long chunk) {
    Ref* r = (Ref*)h;
    REF_TRY
    faiss::IndexIVF* ix = r->index;
    long k = r->es->max_topk;
    ix->set_tune_mode();
    ix->nprobe = ix->nlist;
    std::vector<std::thread> th;
    std::vector<std::string> errs(nthreads);
    long per = chunk > 0 ? chunk : (num + nthreads - 1) / nthreads;
    std::atomic<long> next(0);
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([=, &errs, &next]() {
            omp_set_num_threads(1);
            for (;;) {
                long q0 = next.fetch_add(per), q1 = std::min(num, q0 + per);
                if (q0 >= q1) break;
                try {
                    ix->search(q1 - q0, r->queries.data() + (size_t)(start + q0) * r->d, k,
                               D + q0 * k, I + q0 * k, (size_t)(start + q0));
                } catch (const std::exception& e) {
                    errs[t] = e.what();
                    break;
                }
            }
        });
    }
    for (auto& t : th) t.join();
    ix->set_tune_off();
    for (auto& e : errs)
        if (!e.empty()) throw std::runtime_error(e);
    REF_CATCH
}

int ref_search_fixed_threads(void* h, long n, const float* x, long k, long nprobe, float* D,
                             long* I, int nthreads) {
    Ref* r = (Ref*)h;
    REF_TRY
    ensure_t(r);
    faiss::IndexIVF* ix = r->index;
    ix->nprobe = nprobe;
    std::vector<std::thread> th;
    long per = (n + nthreads - 1) / nthreads;
    int d = r->d;
    for (int t = 0; t < nthreads; t++) {
        long q0 = t * per, q1 = std::min(n, q0 + per);
        if (q0 >= q1) break;
        th.emplace_back([=]() {
            omp_set_num_threads(1);
            ((const faiss::IndexIVF*)ix)->search(q1 - q0, x + (size_t)q0 * d, k, D + q0 * k, I + q0 * k);
        });
    }
    for (auto& t : th) t.join();
    REF_CATCH
}

/* stats (Auncel/IndexIVF.h:361-374) */
void ref_stats_reset() { faiss::indexIVF_stats.reset(); }
void ref_stats_get(double* out6) {
    out6[0] = (double)faiss::indexIVF_stats.nq;
    out6[1] = (double)faiss::indexIVF_stats.nlist;
    out6[2] = (double)faiss::indexIVF_stats.ndis;
    out6[3] = (double)faiss::indexIVF_stats.nheap_updates;
    out6[4] = faiss::indexIVF_stats.quantization_time;
    out6[5] = faiss::indexIVF_stats.search_time;
}

/* ---- IndexShards / IndexReplicas over sub-indexes sharing centroids ---- */

/* subs: array of Ref handles. IndexShards(d, threaded, successive_ids=false) */
int ref_shards_search(void** subs, int nshard, long n, const float* x, long k, long nprobe,
                      float* D, long* I, int threaded) {
    REF_TRY
    Ref* r0 = (Ref*)subs[0];
    faiss::IndexShards sh(r0->d, threaded != 0, false);
    for (int s = 0; s < nshard; s++) {
        Ref* r = (Ref*)subs[s];
        ensure_t(r);
        r->index->nprobe = nprobe;
        sh.add_shard(r->index);
    }
    sh.search(n, x, k, D, I);
    REF_CATCH
}

int ref_replicas_search(void** subs, int nrep, long n, const float* x, long k, long nprobe,
                        float* D, long* I, int threaded) {
    REF_TRY
    Ref* r0 = (Ref*)subs[0];
    faiss::IndexReplicas rp(r0->d, threaded != 0);
    for (int s = 0; s < nrep; s++) {
        Ref* r = (Ref*)subs[s];
        ensure_t(r);
        r->index->nprobe = nprobe;
        rp.addIndex(r->index);
    }
    rp.search(n, x, k, D, I);
    REF_CATCH
}

/* IndexIVF::copy_subset_to (Auncel/IndexIVF.cpp:1055-1118) into an empty sub-index
 * that already holds the same centroids. */
int ref_copy_subset_to(void* h, void* other, int subset_type, long a1, long a2) {
    REF_TRY
    ((Ref*)h)->index->copy_subset_to(*((Ref*)other)->index, subset_type, a1, a2);
    REF_CATCH
}

/* low-level kernels, for pinning the restatement */
float ref_fvec_L2sqr(const float* x, const float* y, long d) { return faiss::fvec_L2sqr(x, y, d); }
float ref_fvec_inner_product(const float* x, const float* y, long d) {
    return faiss::fvec_inner_product(x, y, d);
}
float ref_cosine_theorem(float a, float b, float c) { return faiss::cosine_theorem(a, b, c); }
float ref_kscaling(float kdis, long in, const float* gt, long max_topk) {
    return faiss::kscaling(kdis, in, gt, max_topk);
}

}  // extern "C"
