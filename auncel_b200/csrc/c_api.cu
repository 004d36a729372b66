// extern "C" boundary (include/auncel_b200.h) + the host-side pieces of the path that the
// reference keeps on the host as well: k-means driver, calibration bookkeeping (Trace::SB).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <utility>

#include "../../include/auncel_b200.h"
#include "engine.h"

using namespace auncel;

struct AuncelIndex_H {
    IvfIndex ix;
    AuncelIndex_H(int d, long nlist, int metric, int device) : ix(d, nlist, metric, device) {}
    // host-API staging
    DevBuf<float> x, D, acc, gt, trec, snap, dtbo;
    DevBuf<long long> I;
    DevBuf<unsigned long long> np;
    // Index::search is const and callable from several threads in the reference (IndexReplicas /
    // IndexShards worker threads); calls on one handle share its stream and scratch, so they are
    // serialised here.
    std::recursive_mutex mu;
};
#define LOCK(idx) std::lock_guard<std::recursive_mutex> lock_((idx)->mu)

static thread_local std::string g_last_error;
void auncel_set_last_error(const std::string& m) { g_last_error = m; }          // shards.cu
auncel::IvfIndex* auncel_index_engine(AuncelIndex* idx) { return &idx->ix; }    // shards.cu

#define API_TRY try {
#define API_CATCH                                           \
    }                                                       \
    catch (const auncel::Error& e) {                        \
        g_last_error = e.what();                            \
        return e.code;                                      \
    }                                                       \
    catch (const std::exception& e) {                       \
        g_last_error = e.what();                            \
        return -4;                                          \
    }                                                       \
    catch (...) {                                           \
        g_last_error = "unknown exception";                 \
        return -1;                                          \
    }                                                       \
    return 0;

namespace {

void h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
}
void d2h(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
}

// ---- Trace::SB (IVF_pro.cpp:109-149): same std::sort call, same running means ----
void trace_SB(std::vector<std::pair<float, float>>& trace, std::vector<float>& stds, size_t bs) {
    std::sort(trace.begin(), trace.end(),
              [](std::pair<float, float>& left, std::pair<float, float>& right) { return left.first > right.first; });
    size_t size = 0;
    for (auto& dd : trace) size += (dd.first < 0 && dd.second < 0) ? 0 : 1;
    size_t sz = (size + bs - 1) / bs;
    std::vector<std::pair<float, float>> tmp(sz);
    stds.resize(sz);
    for (size_t i = 0; i < sz; i++) {
        size_t left = i * bs, right = std::min((i + 1) * bs, size);
        float ave1 = 0, ave2 = 0;
        for (size_t index = left; index < right; index++) {
            size_t j = index - left;
            ave1 = fadd(fmul(fdiv((float)j, (float)(j + 1)), ave1), fdiv(trace[index].first, (float)(j + 1)));
            ave2 = fadd(fmul(fdiv((float)j, (float)(j + 1)), ave2), fdiv(trace[index].second, (float)(j + 1)));
        }
        double accum = 0.;
        for (size_t index = left; index < right; index++) {
            float df = fsub(trace[index].second, ave2);
            accum += fmul(df, df);
        }
        stds[i] = (float)std::sqrt(accum / bs);
        tmp[i] = std::make_pair(ave1, ave2);
    }
    trace = tmp;
    std::reverse(trace.begin(), trace.end());
    std::reverse(stds.begin(), stds.end());
}

void calibrate(AuncelIndex_H* h, long n, const float* x, int K, const float* gt_D, float* D, long long* I) {
    IvfIndex& ix = h->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    AUNCEL_CHECK(K >= 4 && K <= MAX_K, "max_topk must be in [4, 128]");
    AUNCEL_CHECK(ix.nlist >= 16, "nlist too small for calibration");
    const int ntr = ix.expected_traces(), max_num = ix.max_num();
    const int d = ix.d;
    cudaStream_t s = ix.stream;
    h->x.ensure((size_t)n * d);
    h->D.ensure((size_t)n * K);
    h->I.ensure((size_t)n * K);
    h->snap.ensure((size_t)n * ntr * K);
    h->dtbo.ensure((size_t)n * max_num);
    h2d(h->x.p, x, (size_t)n * d * sizeof(float), s);
    {  // unreached stages keep neutral values -> kscaling finds no match
        std::vector<float> fill((size_t)n * ntr * K, ix.metric == METRIC_L2 ? FLT_MAX : -FLT_MAX);
        h2d(h->snap.p, fill.data(), fill.size() * sizeof(float), s);
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    QueryBatch qb;
    qb.n = n;
    qb.x = h->x.p;
    qb.k = K;
    qb.nprobe = (int)ix.nlist;
    qb.mode = 2;
    qb.D = h->D.p;
    qb.I = h->I.p;
    qb.snapshots = h->snap.p;
    qb.dtb_out = h->dtbo.p;
    ix.search(qb);
    std::vector<float> snap((size_t)n * ntr * K), dtb((size_t)n * max_num);
    d2h(snap.data(), h->snap.p, snap.size() * sizeof(float), s);
    d2h(dtb.data(), h->dtbo.p, dtb.size() * sizeof(float), s);
    if (D) d2h(D, h->D.p, (size_t)n * K * sizeof(float), s);
    if (I) d2h(I, h->I.p, (size_t)n * K * sizeof(long long), s);
    CUDA_CHECK(cudaStreamSynchronize(s));

    // training block of search_preassigned, IndexIVF.cpp:656-672, on the snapshots
    const size_t per = (size_t)(K / 4) * n;
    std::vector<std::vector<std::pair<float, float>>> traces(ntr);
    int err = 0;
    for (int t = 0; t < ntr; t++) {
        traces[t].assign(per, std::make_pair(-1.f, -1.f));  // IndexIVF.cpp:213-217
        const int stage = 1 << t;
        for (long q = 0; q < n; q++) {
            const float* sp = snap.data() + ((size_t)q * ntr + t) * K;
            const float* dq = dtb.data() + (size_t)q * max_num;
            int count = 0;
            for (int ij = 0; ij < K; ij++) {
                float ks = kscaling(sp[ij], ij, gt_D + (size_t)q * K, K);
                if (ks < 0) break;
                float tval = sp[ij];
                if (ix.metric == METRIC_IP) tval = arcos_lookup(ix.h_arcos.data(), (int)ix.h_arcos.size(), tval, &err);
                float sum_a = sum_angle(tval, dq, 15, stage - 1, ix.h_arcos.data(), (int)ix.h_arcos.size(), &err);
                traces[t][(size_t)q * (K / 4) + count++] = std::make_pair(sum_a, ks);
                if (count >= K / 4) break;
            }
        }
    }
    // error_pro::train, IVF_pro.cpp:186-194
    std::vector<long> off(ntr + 1, 0);
    std::vector<float> phi, U, sg;
    for (int t = 0; t < ntr; t++) {
        std::vector<float> stds;
        trace_SB(traces[t], stds, 250);
        AUNCEL_CHECK(!traces[t].empty(), "calibration produced an empty trace (no query matched its ground truth)");
        for (size_t i = 0; i < traces[t].size(); i++) {
            phi.push_back(traces[t][i].first);
            U.push_back(traces[t][i].second);
            sg.push_back(stds[i]);
        }
        off[t + 1] = (long)phi.size();
    }
    ix.set_error_model((int)ix.h_arcos.size(), ntr, off.data(), phi.data(), U.data(), sg.data(), ix.multipler,
                       ix.std_m);
    ix.stats.err_bits |= (uint64_t)err;
}

}  // namespace

extern "C" {

const char* auncel_get_last_error(void) { return g_last_error.c_str(); }

int auncel_index_new(AuncelIndex** out, int d, int64_t nlist, int metric, int device) {
    API_TRY
    AUNCEL_CHECK(out != nullptr, "null output handle");
    int ndev = 0;
    CUDA_CHECK(cudaGetDeviceCount(&ndev));
    AUNCEL_CHECK(device >= 0 && device < ndev, "no such CUDA device");
    *out = new AuncelIndex_H(d, (long)nlist, metric, device);
    API_CATCH
}

void auncel_index_free(AuncelIndex* idx) { delete idx; }
int auncel_index_d(const AuncelIndex* idx) { return idx->ix.d; }
int64_t auncel_index_nlist(const AuncelIndex* idx) { return idx->ix.nlist; }
int64_t auncel_index_ntotal(const AuncelIndex* idx) { return idx->ix.ntotal; }
int auncel_index_is_trained(const AuncelIndex* idx) { return idx->ix.trained ? 1 : 0; }

int auncel_index_wait_stream(AuncelIndex* idx, void* cuda_stream) {
    API_TRY
    LOCK(idx);
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    CUDA_CHECK(cudaEventRecord(ix.ev_in, (cudaStream_t)cuda_stream));
    CUDA_CHECK(cudaStreamWaitEvent(ix.stream, ix.ev_in, 0));
    API_CATCH
}

int auncel_index_set_centroids(AuncelIndex* idx, const float* centroids, int compute_interdis) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    AUNCEL_CHECK(centroids != nullptr, "null centroids");
    idx->ix.set_centroids(centroids, compute_interdis != 0);
    API_CATCH
}
int auncel_index_get_centroids(const AuncelIndex* idx, float* out) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    idx->ix.get_centroids(out);
    API_CATCH
}
int auncel_index_get_interdis(const AuncelIndex* idx, float* out) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    idx->ix.get_interdis(out);
    API_CATCH
}
int auncel_index_set_interdis(AuncelIndex* idx, const float* in) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    idx->ix.set_interdis(in);
    API_CATCH
}

int auncel_index_train(AuncelIndex* idx, int64_t n, const float* x, int niter, int tune) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    train_kmeans(idx->ix, (long)n, x, niter > 0 ? niter : 25, tune != 0);  // cp.niter = 25, IndexIVF.cpp:54
    API_CATCH
}

int auncel_index_add_device(AuncelIndex* idx, int64_t n, const float* x_dev, const int64_t* ids,
                            const int64_t* list_no) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    idx->ix.add_device((long)n, x_dev, (const long long*)ids, (const long long*)list_no);
    API_CATCH
}

int auncel_index_add(AuncelIndex* idx, int64_t n, const float* x, const int64_t* ids, const int64_t* list_no) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    const long bs = 1L << 22;  // bounded staging; order of appends is preserved
    for (long i0 = 0; i0 < n; i0 += bs) {
        long m = std::min(bs, (long)n - i0);
        idx->x.ensure((size_t)m * ix.d);
        h2d(idx->x.p, x + (size_t)i0 * ix.d, (size_t)m * ix.d * sizeof(float), ix.stream);
        ix.add_device(m, idx->x.p, ids ? (const long long*)ids + i0 : nullptr,
                      list_no ? (const long long*)list_no + i0 : nullptr);
    }
    API_CATCH
}

int auncel_index_assign(AuncelIndex* idx, int64_t n, const float* x, int64_t* list_no) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    idx->x.ensure((size_t)n * ix.d);
    h2d(idx->x.p, x, (size_t)n * ix.d * sizeof(float), ix.stream);
    ix.assign_device((long)n, idx->x.p, (long long*)list_no);
    API_CATCH
}

int auncel_index_reset(AuncelIndex* idx) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    idx->ix.reset();
    API_CATCH
}

int auncel_index_get_lists(const AuncelIndex* idx, float* codes, int64_t* ids) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    const IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    const long long nt = ix.h_list_off[ix.nlist];
    if (nt == 0) return 0;
    if (codes)
        CUDA_CHECK(cudaMemcpy2D(codes, ix.d * sizeof(float), ix.codes.p, ix.dpad * sizeof(float), ix.d * sizeof(float),
                                nt, cudaMemcpyDeviceToHost));
    if (ids) CUDA_CHECK(cudaMemcpy(ids, ix.ids.p, nt * sizeof(long long), cudaMemcpyDeviceToHost));
    API_CATCH
}

int auncel_index_get_params(const AuncelIndex* idx, float* multipler, float* std_m) {
    *multipler = idx->ix.multipler;
    *std_m = idx->ix.std_m;
    return 0;
}

int auncel_index_has_interdis(const AuncelIndex* idx) { return idx->ix.have_interdis ? 1 : 0; }

int auncel_index_list_sizes(const AuncelIndex* idx, int64_t* out) {
    for (long l = 0; l < idx->ix.nlist; l++) out[l] = idx->ix.h_list_off[l + 1] - idx->ix.h_list_off[l];
    return 0;
}

__global__ void widen_keys_kernel(const float* dis, const int* keys, long n, long nlist, long nprobe, float* odis,
                                  long long* okeys) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n * nprobe) return;
    long q = i / nprobe, p = i - q * nprobe;
    odis[i] = dis[q * nlist + p];
    okeys[i] = keys[q * nlist + p];
}

int auncel_index_coarse_search(AuncelIndex* idx, int64_t n, const float* x, int64_t nprobe, float* coarse_dis,
                               int64_t* keys) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    AUNCEL_CHECK(ix.trained, "index is not trained");
    AUNCEL_CHECK(nprobe >= 1 && nprobe <= ix.nlist, "nprobe out of range");
    if (n == 0) return 0;
    idx->x.ensure((size_t)n * ix.d);
    h2d(idx->x.p, x, (size_t)n * ix.d * sizeof(float), ix.stream);
    const float* xs = idx->x.p;
    if (ix.dpad != ix.d) {
        ix.q_x.ensure((size_t)n * ix.dpad);
        launch_pad_rows(idx->x.p, n, ix.d, ix.q_x.p, ix.dpad, ix.stream);
        xs = ix.q_x.p;
    }
    ix.coarse_rank((long)n, xs);
    if (ix.exact_ties)
        launch_fix_ties(ix.metric, ix.c_raw.p, ix.nlist, (int)nprobe, ix.entry_table((int)nprobe), nullptr, (int)n, ix.c_tie0.p, (int)nprobe,
                        nullptr, ix.fix_list.p, ix.ctl.p + 5, ix.c_dis.p, ix.c_keys.p, ix.stream);
    idx->D.ensure((size_t)n * nprobe);
    idx->I.ensure((size_t)n * nprobe);
    widen_keys_kernel<<<(unsigned)((n * nprobe + 255) / 256), 256, 0, ix.stream>>>(ix.c_dis.p, ix.c_keys.p, n, ix.nlist,
                                                                                 nprobe, idx->D.p, idx->I.p);
    d2h(coarse_dis, idx->D.p, (size_t)n * nprobe * sizeof(float), ix.stream);
    d2h(keys, idx->I.p, (size_t)n * nprobe * sizeof(long long), ix.stream);
    CUDA_CHECK(cudaStreamSynchronize(ix.stream));
    API_CATCH
}

int auncel_index_search_device(AuncelIndex* idx, int64_t n, const float* x_dev, int64_t k, int64_t nprobe,
                               int64_t max_codes, float* distances_dev, int64_t* labels_dev) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    QueryBatch qb;
    qb.n = (long)n;
    qb.x = x_dev;
    qb.k = (int)k;
    qb.nprobe = (int)std::min<int64_t>(nprobe, idx->ix.nlist);
    qb.max_codes = (long)max_codes;
    qb.mode = 0;
    qb.D = distances_dev;
    qb.I = (long long*)labels_dev;
    idx->ix.search(qb);
    API_CATCH
}

int auncel_index_search(AuncelIndex* idx, int64_t n, const float* x, int64_t k, int64_t nprobe, int64_t max_codes,
                        float* distances, int64_t* labels) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    AUNCEL_CHECK(k >= 1 && k <= MAX_K, "k must be in [1, 128]");
    if (n == 0) return 0;
    idx->x.ensure((size_t)n * ix.d);
    idx->D.ensure((size_t)n * k);
    idx->I.ensure((size_t)n * k);
    h2d(idx->x.p, x, (size_t)n * ix.d * sizeof(float), ix.stream);
    QueryBatch qb;
    qb.n = (long)n;
    qb.x = idx->x.p;
    qb.k = (int)k;
    qb.nprobe = (int)std::min<int64_t>(nprobe, ix.nlist);
    qb.max_codes = (long)max_codes;
    qb.D = idx->D.p;
    qb.I = idx->I.p;
    ix.search(qb);
    d2h(distances, idx->D.p, (size_t)n * k * sizeof(float), ix.stream);
    d2h(labels, idx->I.p, (size_t)n * k * sizeof(long long), ix.stream);
    CUDA_CHECK(cudaStreamSynchronize(ix.stream));
    API_CATCH
}

int auncel_index_set_time_model(AuncelIndex* idx, int64_t us_per_list, int64_t ns_per_code) {
    API_TRY
    AUNCEL_CHECK(us_per_list >= 0 && ns_per_code >= 0, "time model costs must be >= 0");
    idx->ix.time_us_per_list = us_per_list;
    idx->ix.time_ns_per_code = ns_per_code;
    API_CATCH
}

int auncel_index_search_timed_device(AuncelIndex* idx, int64_t n, const float* x_dev, int64_t k,
                                     const float* budget_ms_dev, float* distances_dev, int64_t* labels_dev) {
    API_TRY
    LOCK(idx);
    QueryBatch qb;
    qb.n = (long)n;
    qb.x = x_dev;
    qb.k = (int)k;
    qb.nprobe = (int)idx->ix.nlist;  // profile.cpp:237: ix->nprobe = ix->nlist
    qb.mode = 0;                     // time_search does not switch the tune block on (profile.cpp:229-244)
    qb.time_tune = 1;
    qb.require_acc = budget_ms_dev;
    qb.D = distances_dev;
    qb.I = (long long*)labels_dev;
    idx->ix.search(qb);
    API_CATCH
}

int auncel_index_search_timed(AuncelIndex* idx, int64_t n, const float* x, int64_t k, const float* budget_ms,
                              float* distances, int64_t* labels) {
    API_TRY
    LOCK(idx);
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    AUNCEL_CHECK(k >= 1 && k <= MAX_K, "k must be in [1, 128]");
    AUNCEL_CHECK(budget_ms != nullptr, "per-query budget missing");
    if (n == 0) return 0;
    cudaStream_t s = ix.stream;
    idx->x.ensure((size_t)n * ix.d);
    idx->D.ensure((size_t)n * k);
    idx->I.ensure((size_t)n * k);
    idx->acc.ensure(n);
    h2d(idx->x.p, x, (size_t)n * ix.d * sizeof(float), s);
    h2d(idx->acc.p, budget_ms, n * sizeof(float), s);
    int rc = auncel_index_search_timed_device(idx, n, idx->x.p, k, idx->acc.p, idx->D.p, (int64_t*)idx->I.p);
    if (rc != 0) return rc;
    d2h(distances, idx->D.p, (size_t)n * k * sizeof(float), s);
    d2h(labels, idx->I.p, (size_t)n * k * sizeof(long long), s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    API_CATCH
}

int auncel_index_range_search(AuncelIndex* idx, int64_t n, const float* x, float radius, int64_t nprobe,
                              int64_t* lims) {
    API_TRY
    LOCK(idx);
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    AUNCEL_CHECK(lims != nullptr, "lims must hold n + 1 entries");
    idx->x.ensure(std::max<size_t>((size_t)n * ix.d, 1));
    h2d(idx->x.p, x, (size_t)n * ix.d * sizeof(float), ix.stream);
    ix.range_search((long)n, idx->x.p, radius, (int)std::min<int64_t>(nprobe, ix.nlist), (long long*)lims);
    API_CATCH
}

int auncel_index_range_search_results(AuncelIndex* idx, float* distances, int64_t* labels) {
    API_TRY
    LOCK(idx);
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    if (ix.range_total == 0) return 0;
    if (distances) d2h(distances, ix.range_D.p, (size_t)ix.range_total * sizeof(float), ix.stream);
    if (labels) d2h(labels, ix.range_I.p, (size_t)ix.range_total * sizeof(long long), ix.stream);
    CUDA_CHECK(cudaStreamSynchronize(ix.stream));
    API_CATCH
}

int auncel_index_set_error_model(AuncelIndex* idx, int n_traces, const int64_t* trace_off, const float* phi,
                                 const float* U, const float* sigma, float multipler, float std_m) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    std::vector<long> off(trace_off, trace_off + n_traces + 1);
    idx->ix.set_error_model(500, n_traces, off.data(), phi, U, sigma, multipler, std_m);
    API_CATCH
}

int auncel_index_set_params(AuncelIndex* idx, float multipler, float std_m) {
    idx->ix.multipler = multipler;
    idx->ix.std_m = std_m;
    return 0;
}

int auncel_index_n_traces(const AuncelIndex* idx) { return idx->ix.n_traces; }
int64_t auncel_index_trace_size(const AuncelIndex* idx, int t) {
    if (t < 0 || t >= idx->ix.n_traces) return -1;
    return idx->ix.h_trace_off[t + 1] - idx->ix.h_trace_off[t];
}
int auncel_index_get_trace(const AuncelIndex* idx, int t, float* phi, float* U, float* sigma) {
    API_TRY
    AUNCEL_CHECK(t >= 0 && t < idx->ix.n_traces, "no such trace");
    long o = idx->ix.h_trace_off[t], m = idx->ix.h_trace_off[t + 1] - o;
    memcpy(phi, idx->ix.h_phi.data() + o, m * sizeof(float));
    memcpy(U, idx->ix.h_U.data() + o, m * sizeof(float));
    memcpy(sigma, idx->ix.h_sigma.data() + o, m * sizeof(float));
    API_CATCH
}

int auncel_index_calibrate(AuncelIndex* idx, int64_t n, const float* x, int64_t max_topk, const float* gt_D,
                           float* distances, int64_t* labels) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    AUNCEL_CHECK(n > 0 && x && gt_D, "calibration needs queries and their ground-truth distances");
    calibrate(idx, (long)n, x, (int)max_topk, gt_D, distances, (long long*)labels);
    API_CATCH
}

int auncel_index_search_bounded_device(AuncelIndex* idx, int64_t n, const float* x_dev, int64_t max_topk,
                                       int64_t query_topk, const float* require_acc_dev, const float* gt_kth_dev,
                                       uint64_t* my_nprobe_dev, float* t_recalls_dev, int flags,
                                       float* distances_dev, int64_t* labels_dev) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    QueryBatch qb;
    qb.n = (long)n;
    qb.x = x_dev;
    qb.k = (int)max_topk;
    qb.nprobe = (int)idx->ix.nlist;  // profile.cpp:218
    qb.mode = 1;
    qb.query_topk = (int)query_topk;
    qb.require_acc = require_acc_dev;
    qb.gt_kth = gt_kth_dev;
    qb.my_nprobe = (unsigned long long*)my_nprobe_dev;
    qb.t_recalls = t_recalls_dev;
    qb.profile = flags & 1;
    qb.overhead_profile = (flags >> 1) & 1;
    qb.time_tune = (flags >> 2) & 1;
    qb.D = distances_dev;
    qb.I = (long long*)labels_dev;
    idx->ix.search(qb);
    API_CATCH
}

int auncel_index_search_bounded(AuncelIndex* idx, int64_t n, const float* x, int64_t max_topk, int64_t query_topk,
                                const float* require_acc, const float* gt_kth, uint64_t* my_nprobe,
                                float* t_recalls, int flags, float* distances, int64_t* labels) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    IvfIndex& ix = idx->ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    AUNCEL_CHECK(max_topk >= 1 && max_topk <= MAX_K, "max_topk must be in [1, 128]");
    AUNCEL_CHECK(require_acc != nullptr && my_nprobe != nullptr, "require_acc / my_nprobe missing");
    if (n == 0) return 0;
    cudaStream_t s = ix.stream;
    const int K = (int)max_topk;
    idx->x.ensure((size_t)n * ix.d);
    idx->D.ensure((size_t)n * K);
    idx->I.ensure((size_t)n * K);
    idx->acc.ensure(n);
    idx->np.ensure(n);
    h2d(idx->x.p, x, (size_t)n * ix.d * sizeof(float), s);
    h2d(idx->acc.p, require_acc, n * sizeof(float), s);
    h2d(idx->np.p, my_nprobe, n * sizeof(uint64_t), s);
    if (gt_kth) {
        idx->gt.ensure(n);
        h2d(idx->gt.p, gt_kth, n * sizeof(float), s);
    }
    if (t_recalls) {
        idx->trec.ensure(n);
        h2d(idx->trec.p, t_recalls, n * sizeof(float), s);
    }
    int rc = auncel_index_search_bounded_device(idx, n, idx->x.p, max_topk, query_topk, idx->acc.p,
                                                gt_kth ? idx->gt.p : nullptr, (uint64_t*)idx->np.p,
                                                t_recalls ? idx->trec.p : nullptr, flags, idx->D.p,
                                                (int64_t*)idx->I.p);
    if (rc != 0) return rc;
    d2h(distances, idx->D.p, (size_t)n * K * sizeof(float), s);
    d2h(labels, idx->I.p, (size_t)n * K * sizeof(long long), s);
    d2h(my_nprobe, idx->np.p, n * sizeof(uint64_t), s);
    if (t_recalls) d2h(t_recalls, idx->trec.p, n * sizeof(float), s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    API_CATCH
}

int auncel_index_get_stats(const AuncelIndex* idx, double* out8) {
    const SearchStats& st = idx->ix.stats;
    out8[8] = st.scan_ms;
    out8[9] = (double)st.launches;
    out8[10] = (double)st.scan_launches;
    out8[11] = st.coarse_ms;
    out8[12] = (double)st.tc_rounds;
    out8[13] = (double)st.tc_candidates;
    out8[14] = (double)st.tc_fallbacks;
    out8[15] = st.tc_ms;
    out8[16] = (double)st.tc_ndis;
    out8[17] = st.simt_ms;
    out8[18] = (double)st.simt_ndis;
    out8[19] = (double)st.tc_uniq;
    out8[20] = (double)st.tc_staged;
    out8[21] = (double)st.simt_uniq;
    out8[22] = (double)st.simt_staged;
    out8[23] = (double)st.tc_audit_bad;
    out8[24] = (double)st.tc_audit_slots;
    out8[25] = (double)st.tc_audit_cands;
    out8[0] = (double)st.nq;
    out8[1] = (double)st.nlist;
    out8[2] = (double)st.ndis;
    out8[3] = st.search_ms;
    out8[4] = (double)st.rounds;
    out8[5] = (double)st.scan_tiles;
    out8[6] = (double)st.scan_pairs;
    out8[7] = (double)st.err_bits;
    return 0;
}

int auncel_index_get_round_stats(const AuncelIndex* idx, int max_rounds, double* out, int* n_rounds) {
    const auto& rs = idx->ix.round_stats;
    const int n = (int)std::min<size_t>(rs.size(), (size_t)std::max(max_rounds, 0));
    for (int r = 0; r < n; r++)
        for (int j = 0; j < 10; j++) out[r * 10 + j] = rs[r][j];
    if (n_rounds) *n_rounds = (int)rs.size();
    return 0;
}

int auncel_index_set_option(AuncelIndex* idx, const char* name, int value) {
    API_TRY
    std::string n(name ? name : "");
    if (n == "tensor_core_filter") idx->ix.tc_mode = value;      // 0 off, 1 auto, 2 whenever heaps are full
    else if (n == "partial_rank") idx->ix.partial_rank_mode = value;  // 0 never, 1 large batches, 2 whenever nlist >= 4096
    else if (n == "tc_kernel") idx->ix.tc_kernel = value;        // 0 / 1 shared-memory queries, 2 TMEM-resident queries, 3 CTA pairs (measured alternatives)
    else if (n == "tc_stream_min") idx->ix.tc_stream_min = value;  // d > 256: queries per list from which the query tile is streamed
    else if (n == "tc_audit") idx->ix.tc_audit = value;          // tests: exact rescan + comparison of every tensor-core round
    else if (n == "exact_ties") idx->ix.exact_ties = value != 0;  // replay the reference's heap order
    else AUNCEL_THROW(-2, "unknown option " + n);
    API_CATCH
}

int auncel_index_set_pool_budget(AuncelIndex* idx, size_t bytes) {
    idx->ix.pool_budget_bytes = std::max<size_t>(bytes, 1 << 20);
    return 0;
}

int auncel_merge_tables_device(int device, int metric, int64_t n, int64_t k, int64_t nshard,
                               const float* all_distances_dev, const int64_t* all_labels_dev,
                               const int64_t* translations_dev, float* distances_dev, int64_t* labels_dev,
                               void* cuda_stream) {
    API_TRY
    CUDA_CHECK(cudaSetDevice(device));
    launch_merge_tables(metric, n, k, nshard, all_distances_dev, (const long long*)all_labels_dev,
                        (const long long*)translations_dev, distances_dev, (long long*)labels_dev,
                        (cudaStream_t)cuda_stream);
    API_CATCH
}

int auncel_merge_tables(int metric, int64_t n, int64_t k, int64_t nshard, const float* all_distances,
                        const int64_t* all_labels, const int64_t* translations, float* distances,
                        int64_t* labels) {
    API_TRY
    if (n == 0 || k == 0) return 0;
    DevBuf<float> aD, oD;
    DevBuf<long long> aI, oI, tr;
    size_t tot = (size_t)nshard * n * k;
    aD.ensure(tot);
    aI.ensure(tot);
    oD.ensure((size_t)n * k);
    oI.ensure((size_t)n * k);
    CUDA_CHECK(cudaMemcpy(aD.p, all_distances, tot * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(aI.p, all_labels, tot * sizeof(long long), cudaMemcpyHostToDevice));
    if (translations) {
        tr.ensure(nshard);
        CUDA_CHECK(cudaMemcpy(tr.p, translations, nshard * sizeof(long long), cudaMemcpyHostToDevice));
    }
    launch_merge_tables(metric, n, k, nshard, aD.p, aI.p, translations ? tr.p : nullptr, oD.p, oI.p, nullptr);
    CUDA_CHECK(cudaMemcpy(distances, oD.p, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(labels, oI.p, (size_t)n * k * sizeof(long long), cudaMemcpyDeviceToHost));
    API_CATCH
}

int auncel_index_copy_subset_to(const AuncelIndex* idx, AuncelIndex* other, int subset_type, int64_t a1,
                                int64_t a2) {
    API_TRY
    LOCK(const_cast<AuncelIndex*>(idx));
    const IvfIndex& ix = idx->ix;
    IvfIndex& ox = other->ix;
    AUNCEL_CHECK(ix.nlist == ox.nlist && ix.d == ox.d, "incompatible indexes");
    AUNCEL_CHECK(subset_type == 1 || subset_type == 2, "subset type not implemented");
    CUDA_CHECK(cudaSetDevice(ix.device));
    const long nt = ix.h_list_off[ix.nlist];
    std::vector<long long> ids(nt);
    if (nt) CUDA_CHECK(cudaMemcpy(ids.data(), ix.ids.p, nt * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<long long> rows, sel_ids, sel_list;
    size_t accu_n = 0, accu_a1 = 0, accu_a2 = 0;
    for (long l = 0; l < ix.nlist; l++) {
        long long o = ix.h_list_off[l], n = ix.h_list_off[l + 1] - o;
        if (subset_type == 1) {
            for (long long i = 0; i < n; i++)
                if (ids[o + i] % a1 == a2) {
                    rows.push_back(o + i);
                    sel_ids.push_back(ids[o + i]);
                    sel_list.push_back(l);
                }
        } else {
            size_t next_accu_n = accu_n + n;
            size_t next_accu_a1 = next_accu_n * a1 / ix.ntotal;
            size_t i1 = next_accu_a1 - accu_a1;
            size_t next_accu_a2 = next_accu_n * a2 / ix.ntotal;
            size_t i2 = next_accu_a2 - accu_a2;
            for (size_t i = i1; i < i2; i++) {
                rows.push_back(o + i);
                sel_ids.push_back(ids[o + i]);
                sel_list.push_back(l);
            }
            accu_a1 = next_accu_a1;
            accu_a2 = next_accu_a2;
        }
        accu_n += n;
    }
    const long m = (long)rows.size();
    if (m == 0) return 0;
    std::vector<float> host((size_t)m * ix.d);
    const_cast<IvfIndex&>(ix).gather_rows(rows.data(), m, host.data());
    const long saved_total = ox.ntotal;
    int rc = auncel_index_add(other, m, host.data(), (const int64_t*)sel_ids.data(), (const int64_t*)sel_list.data());
    if (rc != 0) return rc;
    ox.ntotal = saved_total + m;
    API_CATCH
}

int auncel_heap_entry_table(int64_t k, int32_t* entry_out) {
    API_TRY
    AUNCEL_CHECK(k >= 1 && k <= (1 << 24) && entry_out != nullptr, "heap_entry_table: bad arguments");
    std::vector<int> e;
    heap_entry_table((int)k, e);
    std::copy(e.begin(), e.end(), entry_out);
    return 0;
    API_CATCH
}

}  // extern "C"
