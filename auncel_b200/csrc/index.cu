// IvfIndex: the device-resident IndexIVFFlat and the host side of the query path.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "merge.cuh"
#include "tcfilter.cuh"

namespace auncel {

IvfIndex::IvfIndex(int d_, long nlist_, int metric_, int device_)
    : d(d_), dpad((d_ + 3) / 4 * 4), nlist(nlist_), metric(metric_), device(device_) {
    AUNCEL_CHECK(d > 0 && nlist > 0, "d and nlist must be positive");
    AUNCEL_CHECK(metric == METRIC_L2 || metric == METRIC_IP, "metric must be 0 (IP) or 1 (L2)");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    num_sms = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreate(&ev0));
    CUDA_CHECK(cudaEventCreate(&ev1));
    CUDA_CHECK(cudaEventCreate(&ev2));
    CUDA_CHECK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    h_list_off.assign(nlist + 1, 0);
    list_off.ensure(nlist + 1);
    CUDA_CHECK(cudaMemset(list_off.p, 0, (nlist + 1) * sizeof(long long)));
    ctl.ensure(CTL_SIZE + 8);
    h_ctl.ensure(CTL_SIZE + 8);
    // arccos LUT: error_pro::construct_arcos, IVF_pro.cpp:151-160 (host libm, like the reference)
    int len = 500;
    h_arcos.resize(len);
    float sc = len / 2;
    for (int i = 0; i < len; i++) h_arcos[i] = std::acos(float(i - sc) / sc);
    d_arcos.ensure(len);
    CUDA_CHECK(cudaMemcpy(d_arcos.p, h_arcos.data(), len * sizeof(float), cudaMemcpyHostToDevice));
}

IvfIndex::~IvfIndex() {
    cudaSetDevice(device);
    if (stream) cudaStreamDestroy(stream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev2) cudaEventDestroy(ev2);
    if (ev_in) cudaEventDestroy(ev_in);
    for (auto e : scan_ev) cudaEventDestroy(e);
    for (auto e : tc_ev) cudaEventDestroy(e);
}

void IvfIndex::set_centroids(const float* c, bool compute_interdis) {
    CUDA_CHECK(cudaSetDevice(device));
    io_f.ensure((size_t)nlist * d);
    CUDA_CHECK(cudaMemcpyAsync(io_f.p, c, (size_t)nlist * d * sizeof(float), cudaMemcpyHostToDevice, stream));
    centroids.ensure((size_t)nlist * dpad);
    launch_pad_rows(io_f.p, nlist, d, centroids.p, dpad, stream);
    trained = true;
    have_interdis = false;
    if (compute_interdis) {
        // Level1Quantizer::train_q1's Auncel part, IndexIVF.cpp:97-109.
        size_t tri = (size_t)nlist * (nlist - 1) / 2;
        interdis.ensure(std::max<size_t>(tri, 1));
        if (metric == METRIC_L2) {
            launch_interdis(metric, centroids.p, nlist, dpad, interdis.p, stream);
        } else {
            // The reference rescales centroid 0 by its own norm nlist times (the loop never
            // advances `st`, :102-107) and then takes acos of the inner products with libm.
            // Same here: host rescale of row 0, device inner products, host std::acos.
            std::vector<float> c0(c, c + d);
            for (long i = 0; i < nlist; i++) {
                float s[4] = {0, 0, 0, 0};
                int j = 0;
                for (; j + 4 <= d; j += 4)
                    for (int l = 0; l < 4; l++) s[l] = fadd(s[l], fmul(c0[j + l], c0[j + l]));
                for (int l = 0; l < 4; l++) {
                    float v = j + l < d ? c0[j + l] : 0.f;
                    s[l] = fadd(s[l], fmul(v, v));
                }
                float norm = sqrtf(fadd(fadd(s[0], s[1]), fadd(s[2], s[3])));
                for (int jj = 0; jj < d; jj++) c0[jj] = fdiv(c0[jj], norm);
            }
            DevBuf<float> tmp;
            tmp.ensure((size_t)nlist * dpad);
            CUDA_CHECK(cudaMemcpyAsync(tmp.p, centroids.p, (size_t)nlist * dpad * sizeof(float),
                                       cudaMemcpyDeviceToDevice, stream));
            std::vector<float> row0(dpad, 0.f);
            std::copy(c0.begin(), c0.end(), row0.begin());
            CUDA_CHECK(cudaMemcpyAsync(tmp.p, row0.data(), dpad * sizeof(float), cudaMemcpyHostToDevice, stream));
            launch_interdis(metric, tmp.p, nlist, dpad, interdis.p, stream);
            std::vector<float> h(tri);
            CUDA_CHECK(cudaMemcpyAsync(h.data(), interdis.p, tri * sizeof(float), cudaMemcpyDeviceToHost, stream));
            CUDA_CHECK(cudaStreamSynchronize(stream));
            for (size_t i = 0; i < tri; i++) h[i] = std::acos(h[i]);
            CUDA_CHECK(cudaMemcpyAsync(interdis.p, h.data(), tri * sizeof(float), cudaMemcpyHostToDevice, stream));
        }
        have_interdis = true;
    }
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

void IvfIndex::get_centroids(float* out) const {
    CUDA_CHECK(cudaSetDevice(device));
    AUNCEL_CHECK(trained, "index has no centroids");
    CUDA_CHECK(cudaMemcpy2D(out, d * sizeof(float), centroids.p, dpad * sizeof(float), d * sizeof(float), nlist,
                            cudaMemcpyDeviceToHost));
}

void IvfIndex::get_interdis(float* out) const {
    CUDA_CHECK(cudaSetDevice(device));
    AUNCEL_CHECK(have_interdis, "interdis_cem was not computed");
    CUDA_CHECK(cudaMemcpy(out, interdis.p, (size_t)nlist * (nlist - 1) / 2 * sizeof(float), cudaMemcpyDeviceToHost));
}

void IvfIndex::set_interdis(const float* in) {
    CUDA_CHECK(cudaSetDevice(device));
    size_t tri = (size_t)nlist * (nlist - 1) / 2;
    interdis.ensure(std::max<size_t>(tri, 1));
    CUDA_CHECK(cudaMemcpy(interdis.p, in, tri * sizeof(float), cudaMemcpyHostToDevice));
    have_interdis = true;
}

void IvfIndex::reset() {
    ntotal = 0;
    std::fill(h_list_off.begin(), h_list_off.end(), 0);
    CUDA_CHECK(cudaSetDevice(device));
    CUDA_CHECK(cudaMemset(list_off.p, 0, (nlist + 1) * sizeof(long long)));
}

// quantizer->assign (Index.cpp:42-47): k = 1 search; first strictly-best centroid wins.
void IvfIndex::assign_device(long n, const float* x_dev, long long* list_no_host) {
    CUDA_CHECK(cudaSetDevice(device));
    AUNCEL_CHECK(trained, "index is not trained");
    const long chunk = 1 << 20;
    q_x.ensure((size_t)std::min(n, chunk) * dpad);
    assign_best.ensure(std::min(n, chunk));
    std::vector<unsigned long long> h(std::min(n, chunk));
    for (long i0 = 0; i0 < n; i0 += chunk) {
        long m = std::min(chunk, n - i0);
        const float* xs = x_dev + i0 * d;
        if (dpad != d) {
            launch_pad_rows(xs, m, d, q_x.p, dpad, stream);
            xs = q_x.p;
        }
        for (long j0 = 0; j0 < m; j0 += 65535L * 64) {
            long mm = std::min(65535L * 64, m - j0);
            launch_coarse_distances(metric, xs + j0 * dpad, mm, centroids.p, nlist, dpad, nullptr,
                                    assign_best.p + j0, stream);
        }
        CUDA_CHECK(cudaMemcpyAsync(h.data(), assign_best.p, m * sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        for (long i = 0; i < m; i++) list_no_host[i0 + i] = (long long)(h[i] & 0xffffffffull);
    }
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, int d, const long long* __restrict__ dst_row,
                                    long n, float* __restrict__ dst, int dpad) {
    // one warp per row
    long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= n) return;
    long long o = dst_row[r];
    if (o < 0) return;
    for (int c = lane; c < dpad; c += 32) dst[o * dpad + c] = c < d ? src[r * (long)d + c] : 0.f;
}

__global__ void scatter_ids_kernel(const long long* __restrict__ ids, const long long* __restrict__ dst_row, long n,
                                   long long* __restrict__ dst) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long o = dst_row[i];
    if (o >= 0) dst[o] = ids[i];
}

__global__ void move_lists_kernel(const float* __restrict__ old_codes, const long long* __restrict__ old_ids,
                                  const long long* __restrict__ old_off, const long long* __restrict__ new_off,
                                  long nlist, int dpad, float* __restrict__ codes, long long* __restrict__ ids) {
    // one CTA per list: copy the existing prefix of every list to its new place
    long l = blockIdx.x;
    long long o0 = old_off[l], cnt = old_off[l + 1] - o0, n0 = new_off[l];
    long long words = cnt * dpad;
    for (long long i = threadIdx.x; i < words; i += blockDim.x) codes[n0 * dpad + i] = old_codes[o0 * dpad + i];
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) ids[n0 + i] = old_ids[o0 + i];
}

// IndexIVFFlat::add_core (IndexIVFFlat.cpp:41-80): every vector is appended to the list of
// its nearest centroid, in input order (the in-list order matters at distance ties).  Lists
// live in one arena in list order, so an add rebuilds the arena: old prefix + new suffix.
void IvfIndex::add_device(long n, const float* x_dev, const long long* ids_host, const long long* list_no_host) {
    CUDA_CHECK(cudaSetDevice(device));
    AUNCEL_CHECK(trained, "index is not trained");
    if (n == 0) return;
    std::vector<long long> ln;
    if (!list_no_host) {
        ln.resize(n);
        assign_device(n, x_dev, ln.data());
        list_no_host = ln.data();
    }
    // stable counting sort on the host: destination row of every new vector
    std::vector<long long> add_cnt(nlist, 0);
    long nadd = 0;
    for (long i = 0; i < n; i++) {
        long long l = list_no_host[i];
        if (l < 0) continue;  // IndexIVFFlat.cpp:65-66
        AUNCEL_CHECK(l < nlist, "list number out of range");
        add_cnt[l]++;
        nadd++;
    }
    std::vector<long long> new_off(nlist + 1, 0);
    for (long l = 0; l < nlist; l++)
        new_off[l + 1] = new_off[l] + (h_list_off[l + 1] - h_list_off[l]) + add_cnt[l];
    std::vector<long long> cursor(nlist);
    for (long l = 0; l < nlist; l++) cursor[l] = new_off[l] + (h_list_off[l + 1] - h_list_off[l]);
    std::vector<long long> dst_row(n), idv(n);
    for (long i = 0; i < n; i++) {
        long long l = list_no_host[i];
        dst_row[i] = l < 0 ? -1 : cursor[l]++;
        idv[i] = ids_host ? ids_host[i] : ntotal + i;  // IndexIVFFlat.cpp:62
    }
    long long new_total = new_off[nlist];
    DevBuf<float> ncodes;
    DevBuf<long long> nids, d_new_off, d_dst, d_ids;
    ncodes.ensure(std::max<size_t>((size_t)new_total * dpad, 4));
    nids.ensure(std::max<size_t>(new_total, 1));
    d_new_off.ensure(nlist + 1);
    d_dst.ensure(n);
    d_ids.ensure(n);
    CUDA_CHECK(cudaMemcpyAsync(d_new_off.p, new_off.data(), (nlist + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaMemcpyAsync(d_dst.p, dst_row.data(), n * sizeof(long long), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaMemcpyAsync(d_ids.p, idv.data(), n * sizeof(long long), cudaMemcpyHostToDevice, stream));
    if (h_list_off[nlist] > 0)
        move_lists_kernel<<<(unsigned)nlist, 256, 0, stream>>>(codes.p, ids.p, list_off.p, d_new_off.p, nlist, dpad,
                                                             ncodes.p, nids.p);
    scatter_rows_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(x_dev, d, d_dst.p, n, ncodes.p, dpad);
    scatter_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_ids.p, d_dst.p, n, nids.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(list_off.p, d_new_off.p, (nlist + 1) * sizeof(long long), cudaMemcpyDeviceToDevice, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    std::swap(codes.p, ncodes.p);
    std::swap(codes.cap, ncodes.cap);
    std::swap(ids.p, nids.p);
    std::swap(ids.cap, nids.cap);
    h_list_off = new_off;
    make_codes_tensor_map(codes_tmap, codes.p, new_total, dpad);
    make_codes_tensor_map(codes_tmap64, codes.p, new_total, dpad, 64);
    launch_row_norms(codes.p, new_total, dpad, vnorm.ensure(std::max<size_t>(new_total, 1)), stream);
    launch_list_norm_max(vnorm.p, list_off.p, nlist, list_nmax.ensure(nlist), stream);
    CUDA_CHECK(cudaStreamSynchronize(stream));
    ntotal += n;  // the reference counts skipped (-1) vectors too, IndexIVFFlat.cpp:79
}

__global__ void gather_rows_kernel(const float* __restrict__ codes, int dpad, int d,
                                   const long long* __restrict__ rows, long m, float* __restrict__ out) {
    long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= m) return;
    const float* src = codes + rows[r] * dpad;
    for (int c = lane; c < d; c += 32) out[r * (long)d + c] = src[c];
}

// rows of the arena (arena row numbers, host) -> compact m x d host matrix
void IvfIndex::gather_rows(const long long* rows_host, long m, float* out_host) {
    CUDA_CHECK(cudaSetDevice(device));
    if (m == 0) return;
    DevBuf<long long> r;
    DevBuf<float> o;
    const long chunk = 1L << 22;
    r.ensure(std::min(m, chunk));
    o.ensure((size_t)std::min(m, chunk) * d);
    for (long i0 = 0; i0 < m; i0 += chunk) {
        long mm = std::min(chunk, m - i0);
        CUDA_CHECK(cudaMemcpyAsync(r.p, rows_host + i0, mm * sizeof(long long), cudaMemcpyHostToDevice, stream));
        gather_rows_kernel<<<(unsigned)((mm * 32 + 255) / 256), 256, 0, stream>>>(codes.p, dpad, d, r.p, mm, o.p);
        CUDA_CHECK(cudaMemcpyAsync(out_host + (size_t)i0 * d, o.p, (size_t)mm * d * sizeof(float),
                                   cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }
}

void IvfIndex::set_error_model(int arcos_size, int ntr, const long* trace_off, const float* phi,
                               const float* U, const float* sigma, float mult, float sm) {
    CUDA_CHECK(cudaSetDevice(device));
    AUNCEL_CHECK(arcos_size == (int)h_arcos.size(), "arccos table size must be 500 (IVF_pro.h:86)");
    AUNCEL_CHECK(ntr == expected_traces(), "need one trace per power of two <= nlist/8 (IndexIVF.cpp:209-221)");
    for (int t = 0; t < ntr; t++)
        AUNCEL_CHECK(trace_off[t + 1] > trace_off[t], "every trace needs at least one (phi,U) bucket");
    n_traces = ntr;
    h_trace_off.assign(trace_off, trace_off + ntr + 1);
    long tot = trace_off[ntr];
    h_phi.assign(phi, phi + tot);
    h_U.assign(U, U + tot);
    h_sigma.assign(sigma, sigma + tot);
    d_trace_off.ensure(ntr + 1);
    d_phi.ensure(tot);
    d_U.ensure(tot);
    d_sigma.ensure(tot);
    CUDA_CHECK(cudaMemcpy(d_trace_off.p, h_trace_off.data(), (ntr + 1) * sizeof(long), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d_phi.p, h_phi.data(), tot * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d_U.p, h_U.data(), tot * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d_sigma.p, h_sigma.data(), tot * sizeof(float), cudaMemcpyHostToDevice));
    multipler = mult;
    std_m = sm;
}

const int* IvfIndex::entry_table(int k) {
    if (k != heap_entry_k) {
        std::vector<int> e;
        heap_entry_table(k, e);
        heap_entry.ensure(e.size());  // k entries + the length of the parallel prefix
        CUDA_CHECK(cudaMemcpyAsync(heap_entry.p, e.data(), e.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        heap_entry_k = k;
    }
    return heap_entry.p;
}

ErrModelView IvfIndex::model_view() const {
    ErrModelView m;
    m.arcos = d_arcos.p;
    m.arcos_size = (int)h_arcos.size();
    m.n_traces = n_traces;
    m.trace_off = d_trace_off.p;
    m.phi = d_phi.p;
    m.U = d_U.p;
    m.sigma = d_sigma.p;
    m.std_m = std_m;
    m.multipler = multipler;
    return m;
}

// IndexFlat::search with k = nlist (IndexFlat.cpp:42-56): all centroids, best first.
void IvfIndex::coarse_rank(long n, const float* xs /* n x dpad, device */, bool allow_partial) {
    c_dis.ensure((size_t)n * nlist);
    c_keys.ensure((size_t)n * nlist);
    c_tie0.ensure(n);
    c_sorted.ensure(n);
    fix_list.ensure(n);
    // Auncel mode ranks all nlist centroids per query but probes a few hundred: large batches rank the best
    // rank_rows_partial_width() up front and complete a row only when a round is about to read past that
    partial_rank = allow_partial && partial_rank_mode != 0 && nlist >= 4L * rank_rows_partial_width() &&
                   (n >= 2048 || partial_rank_mode == 2);
    DevBuf<float>& raw = c_raw;  // kept for the whole search: the tie replay reads it
    const long chunk = 65535L * 64;
    raw.ensure((size_t)n * nlist);
    for (long i0 = 0; i0 < n; i0 += chunk) {
        long m = std::min(chunk, n - i0);
        launch_coarse_distances(metric, xs + i0 * dpad, m, centroids.p, nlist, dpad, raw.p + i0 * nlist, nullptr, stream);
        if (partial_rank) {
            launch_rank_rows_partial(metric, raw.p + i0 * nlist, m, nlist, c_dis.p + i0 * nlist, c_keys.p + i0 * nlist,
                                     c_tie0.p + i0, c_sorted.p + i0, stream);
            // rows the partial kernel gave up on (a long run of equal distances at the cut): sorted_upto == 0
            launch_rank_rows(metric, raw.p + i0 * nlist, m, nlist, c_dis.p + i0 * nlist, c_keys.p + i0 * nlist,
                             c_tie0.p + i0, stream, nullptr, c_sorted.p + i0, nullptr, 1);
        } else {
            launch_rank_rows(metric, raw.p + i0 * nlist, m, nlist, c_dis.p + i0 * nlist, c_keys.p + i0 * nlist,
                             c_tie0.p + i0, stream);
        }
    }
}

// IndexIVF::search -> search_preassigned (IndexIVF.cpp:335-736) for a batch of device-resident
// queries.  Rounds: [r0, r0+w) probe ranks per active query -> plan (group pairs by list) ->
// scan (fused selection) -> merge + termination check -> compact the active list.
void IvfIndex::search(const QueryBatch& qb) {
    CUDA_CHECK(cudaSetDevice(device));
    AUNCEL_CHECK(trained, "index is not trained");
    AUNCEL_CHECK(qb.k >= 1 && qb.k <= MAX_K, "k must be in [1, 128]");
    AUNCEL_CHECK(qb.nprobe >= 1, "nprobe must be >= 1");
    const long n = qb.n;
    if (n == 0) return;
    AUNCEL_CHECK(n < (1L << 31), "too many queries in one call");
    const int K = qb.k;
    const int nprobe = (int)std::min<long>(qb.nprobe, nlist);
    if (qb.mode != 0) {
        AUNCEL_CHECK(have_interdis, "tuned search needs interdis_cem (train/set_centroids in tune mode)");
        AUNCEL_CHECK(nprobe == nlist, "Auncel search ranks all centroids: nprobe must equal nlist (profile.cpp:218)");
        AUNCEL_CHECK(nlist >= 16, "nlist too small for the error model");
    }
    if (qb.mode == 1) {
        AUNCEL_CHECK(n_traces > 0, "Search tune start can't start without IVF_pro init and training");
        AUNCEL_CHECK(qb.query_topk >= 1 && qb.query_topk <= K, "query_topk must be in [1, max_topk]");
        AUNCEL_CHECK(qb.require_acc != nullptr, "require_acc missing");
    }
    stats = SearchStats();
    if (tc_audit) CUDA_CHECK(cudaMemsetAsync(audit_ctr.ensure(8), 0, 8 * sizeof(unsigned long long), stream));
    round_log.clear();
    debug_rounds = getenv("AUNCEL_DEBUG_ROUNDS") != nullptr;
    CUDA_CHECK(cudaEventRecord(ev0, stream));

    // ---- stage queries (pad rows), coarse ranking
    const float* xs = qb.x;
    if (dpad != d) {
        q_x.ensure((size_t)n * dpad);
        launch_pad_rows(qb.x, n, d, q_x.p, dpad, stream);
        xs = q_x.p;
    }
    // (a plain search reads only its nprobe best centroids: the partial ranking covers them)
    // (Auncel mode: set_online and the first tie replay read ranks 0 .. max_num = nlist / 8 + 20, which a partial ranking
    // covers up to nlist ~ 8000; beyond that every row is ranked completely)
    coarse_rank(n, xs, qb.max_codes == 0 && !qb.time_tune &&
                           (qb.mode != 0 ? max_num() + 1 < rank_rows_partial_width() : nprobe + 1 < rank_rows_partial_width()));
    if (tc_mode) launch_row_norms(xs, n, dpad, qnorm.ensure(n), stream);
    CUDA_CHECK(cudaEventRecord(ev2, stream));
    CUDA_CHECK(cudaMemsetAsync(ctl.p, 0, (CTL_SIZE + 8) * sizeof(int), stream));
    uint64_t launches = 2 * ((n + 65535L * 64 - 1) / (65535L * 64)) + (dpad != d ? 1 : 0) + 2 /*init, finalize*/ +
                        (qb.mode != 0 ? 1 : 0);

    // ---- per-query state
    state.ensure(QState::bytes(n, K));
    RoundParams rp;
    rp.codes = codes.p;
    rp.list_off = list_off.p;
    rp.ids = ids.p;
    rp.dpad = dpad;
    rp.nlist = nlist;
    rp.metric = metric;
    rp.xq = xs;
    rp.ckeys = c_keys.p;
    rp.cdis = c_dis.p;
    rp.n = n;
    rp.K = K;
    rp.st.carve(state.p, n, K);
    rp.ctl = ctl.p;
    // shards with single-index semantics: error-bounded and calibration searches only (a plain search
    // merges result tables instead, IndexShards.cpp:44-105)
    const bool sharded = shard_x != nullptr && shard_x->world > 1 && qb.mode != 0;
    rp.shard_rank = sharded ? shard_x->rank : -1;
    if (sharded) {
        AUNCEL_CHECK(qb.max_codes == 0 && !qb.time_tune, "sharded rounds: max_codes / time_tune cut on the local list sizes");
        AUNCEL_CHECK(h_list_off[nlist] > 0, "sharded rounds: empty shard");
        AUNCEL_CHECK(!tc_audit, "sharded rounds: tc_audit is a single-GPU check");
        long long longest = 0;
        for (long l = 0; l < nlist; l++) longest = std::max(longest, h_list_off[l + 1] - h_list_off[l]);
        AUNCEL_CHECK(longest < (1ll << SHARD_CODE_SHIFT) && shard_x->world <= (1 << (32 - SHARD_CODE_SHIFT)),
                     "sharded rounds: list too long / too many shards for the 32-bit candidate code");
        shard_x->entries_sent = shard_x->entries_recv = shard_x->bytes_recv = shard_x->exchanges = 0;
    }
    rp.round_work = round_work.ensure(4);
    h_round_work.ensure(4);
    std::vector<uint64_t> round_uniq, round_staged;
    std::vector<int> tc_round_of;     // per round: index of its tensor-core event pair, or -1
    std::vector<uint64_t> round_ndis;
    std::vector<std::array<double, 3>> round_meta;  // r0, w, active queries
    round_stats.clear();
    rp.list_cnt = list_cnt.ensure(nlist);
    rp.list_pair_off = list_pair_off.ensure(nlist + 1);
    rp.list_tile_off = list_tile_off.ensure(nlist + 1);
    rp.list_cursor = list_cursor.ensure(nlist);
    active.ensure(n);
    active2.ensure(n);

    TuneParams tp;
    tp.mode = qb.mode;
    tp.query_topk = qb.query_topk;
    tp.profile = qb.profile;
    tp.overhead_profile = qb.overhead_profile;
    tp.nprobe = nprobe;
    tp.max_codes = qb.max_codes;
    tp.time_tune = qb.time_tune;
    tp.us_per_list = time_us_per_list;
    tp.ns_per_code = time_ns_per_code;
    if (qb.time_tune) AUNCEL_CHECK(qb.require_acc != nullptr, "time_tune needs the per-query budget (require_acc, ms)");
    tp.model = model_view();
    tp.require_acc = qb.require_acc;
    tp.gt_kth = qb.gt_kth;
    tp.t_recalls = qb.t_recalls;
    tp.dtb = nullptr;
    tp.max_num = max_num();
    tp.snapshots = qb.snapshots;
    tp.n_traces = expected_traces();
    // Small batches: one replay wave fits every tie-affected query, so fix all ranks up front
    // instead of paying one wave per round.  Large batches: only what set_online reads
    // (ranks 0..max_num) now, the rest lazily before each round.
    const bool ties_all_upfront = exact_ties && n >= 16 && n <= 1024 && !partial_rank;  // (a partial ranking only knows the ties of its prefix)
    if (exact_ties && (qb.mode != 0 || ties_all_upfront))
        launch_fix_ties(metric, c_raw.p, nlist, nprobe, entry_table(nprobe), nullptr, (int)n, c_tie0.p,
                        ties_all_upfront ? nprobe : tp.max_num + 1, nullptr, fix_list.p, ctl.p + CTL_NFIX, c_dis.p,
                        c_keys.p, stream, nullptr, 0, nullptr, partial_rank ? c_sorted.p : nullptr);
    if (qb.mode != 0) {
        float* dtb_p = qb.dtb_out ? qb.dtb_out : dtb.ensure((size_t)n * tp.max_num);
        launch_set_online(metric, nlist, n, c_dis.p, c_keys.p, interdis.p, d_arcos.p, (int)h_arcos.size(), dtb_p,
                          tp.max_num, ctl.p, stream);
        tp.dtb = dtb_p;
    }
    launch_init_state(rp, tp, qb.mode == 1 ? qb.my_nprobe : nullptr, active.p, stream);

    // ---- rounds
    int n_active = h_list_off[nlist] > 0 ? (int)n : 0;  // an empty index has nothing to scan
    int min_rcnt = 0, not_full = (int)n, rem_sum = 0;
    bool ties_done = false;
    const int tc_min_r0 = getenv("AUNCEL_TC_MIN_R0") ? atoi(getenv("AUNCEL_TC_MIN_R0")) : 1;
    const int wide_slot_r0 = getenv("AUNCEL_WIDE_R0") ? atoi(getenv("AUNCEL_WIDE_R0")) : 8;  // rounds starting below this rank get wide slots
    int r0 = 0;
    int* act_cur = active.p;
    int* act_nxt = active2.p;
    const int max_stage = qb.mode == 2 ? (int)std::min<long>(nprobe, nlist / 8 + 1) : nprobe;
    const size_t pool_entries = std::max<size_t>(pool_budget_bytes / 8, (size_t)K);
    float scan_ms_total = 0.f;
    while (n_active > 0 && r0 < max_stage) {
        // window: fixed/calibration scans as wide as the pool allows; the error-bounded search grows
        // its window with the rank already reached (see below), so undecided queries never speculate
        // past what the reference itself would scan.
        long w_cap = (long)(pool_budget_bytes / ((size_t)n_active * ((size_t)K * 8 + (size_t)dpad * 4)));
        static const long w_hard = getenv("AUNCEL_WMAX") ? atol(getenv("AUNCEL_WMAX")) : 4096;
        w_cap = std::max(1L, std::min<long>(w_cap, w_hard));
        long w = max_stage - r0;
        if (qb.mode == 1 && !qb.overhead_profile) {
            // few queries: start with a wider window -- speculative lists cost little HBM time,
            // every extra round costs a fixed launch/sync latency
            // (measured: 32 lists for up to 16 queries, 64 for 32..128 -- batch 64: 3.86 instead of 4.32 ms per call)
            static const long w0_env = getenv("AUNCEL_W0") ? atol(getenv("AUNCEL_W0")) : 0;
            const long w0_cap = w0_env > 0 ? w0_env : (n >= 32 && n <= 128 ? 64 : 32);
            const long w0 = std::max(1L, std::min(w0_cap, (w0_cap * 128) / n));
            // A query still undecided after r0 lists will stop no earlier than multipler * r0
            // (my_nprobe = stage * multipler, IndexIVF.cpp:615-626), so growing the window by up to
            // that factor scans nothing that would not be scanned anyway -- and every round saved is
            // one pass over the arena saved.
            double g = std::min<double>(std::max<double>(multipler, 2.0), 8.0);
            if (const char* e = getenv("AUNCEL_GROWTH")) g = std::max(2.0, atof(e));
            w = std::min<long>(w, std::max<long>(w0, (long)(r0 * (g - 1.0))));
            // the round after the first list still runs on the exact FP32 scan (loose thresholds would
            // overflow the tensor-core filter's per-pair slots): keep it short, the filter takes over next
            // (measured: 3 lists at d = 128, no cap at d = 96 where the exact scan is cheaper per pair)
            // Large batches: the round after the first list already runs on the tensor-core filter (wide
            // slots).  Such a round streams every list once whatever the number of pairs per list (up to a
            // tile of queries), so it covers more ranks than the growth rule alone would give: ranks an
            // early-deciding query does not need cost filter flops only, and one pass over the arena is saved.
            static const long w1_env = getenv("AUNCEL_W1") ? atol(getenv("AUNCEL_W1")) : -1;
            const long w1 = w1_env >= 0 ? w1_env : 31;
            if (w1 > 0 && stats.rounds == 1 && n >= 2048 && tc_mode == 1) w = std::min<long>(max_stage - r0, std::max<long>(w, w1));
        } else if (qb.mode == 0 && n >= 2048 && tc_mode == 1 && max_stage > 8) {
            // plain search, large batch: the first list exactly (it fills the heaps), then the tensor-core
            // filter -- one round of 31 ranks with wide slots, then everything that is left (every query scans
            // all nprobe lists anyway; a filter round costs one pass over the lists whatever its width)
            w = stats.rounds == 0 ? 1 : stats.rounds == 1 ? std::min<long>(w, 31) : w;
        } else if ((long)n_active * max_stage >= 4096 && max_stage > 8) {
            // plain / calibration search: a few narrow rounds first, so that the bulk of the
            // lists is scanned against a tight threshold (cheap selection)
            w = std::min<long>(w, std::max(2, 2 * r0));
        }
        w = std::min(w, w_cap);
        // segments: split lists when there are too few (list, query-tile) units to fill the GPU
        long est_tiles = std::min<long>((long)n_active * w, nlist) + (long)n_active * w / SCAN_QT;  // lower bound
        long S = (2L * num_sms + est_tiles - 1) / est_tiles;
        S = std::max(1L, std::min<long>(S, 32));
        // narrow tiles (8 queries, rows split over warps) when lists are probed by few queries
        const double avg_q = (double)n_active * w / (double)std::min<long>(nlist, (long)n_active * w);
        static const double nsub_t4 = getenv("AUNCEL_NSUB_T4") ? atof(getenv("AUNCEL_NSUB_T4")) : 10.0;
        static const double nsub_t2 = getenv("AUNCEL_NSUB_T2") ? atof(getenv("AUNCEL_NSUB_T2")) : 24.0;
        int nsub = avg_q <= nsub_t4 ? 4 : avg_q <= nsub_t2 ? 2 : 1;
        // a handful of queries: the per-query merge of S*nsub partial results per list is the latency,
        // not the scan -- do not split rows over warps as well
        if (n_active <= 64) nsub = 1;
        while (nsub > 1 && (size_t)n_active * w * nsub * K > pool_entries) nsub >>= 1;
        while (S > 1 && (size_t)n_active * w * S * nsub * K > pool_entries) S--;
        static const int nc4 = getenv("AUNCEL_NC4") ? atoi(getenv("AUNCEL_NC4")) : 1;
        rp.nc = (nsub == 4 && nc4) ? 4 : 8;
        rp.qt = 4 * rp.nc / nsub;
        rp.nsub = nsub;
        rp.unsorted = 0;
        rp.merged = 0;
        rp.defer_sort = (n_active >= 1024 && !sharded) ? 1 : 0;  // few queries: one merge warp per query would sort serially
        rp.filtered = 0;
        rp.pair_flag = nullptr;
        rp.redo_ord = nullptr;
        rp.redo_d = nullptr;
        rp.redo_off = nullptr;
        rp.redo_cnt = nullptr;
        // tensor-core filter round: (nearly) every remaining query already holds K results, so only
        // a handful of vectors per list can still enter -- filter with TF32 MMAs, rerank exactly.
        // The few queries whose heaps are not full yet let everything through and are redone by
        // the exact scan (per-pair overflow), so they only cost time, never correctness.
        const bool mostly_full = stats.rounds > 0 && (long)not_full * 50 <= (long)n_active;
        bool use_tc = tc_mode == 2 ? (stats.rounds > 0 && min_rcnt >= K)
                                   : (tc_mode == 1 && mostly_full && r0 >= tc_min_r0 && avg_q >= 4.0 && (long)n_active * w >= 2048);
        if (use_tc && getenv("AUNCEL_NO_TC")) use_tc = false;
        // which filter kernel: tcfilter.cu (queries in shared memory), or on request tcfilter2.cu (queries
        // resident in TMEM, all shared memory streams lists) -- measured slower, see its header
        static const int tck_env = getenv("AUNCEL_TC_KERNEL") ? atoi(getenv("AUNCEL_TC_KERNEL")) : 0;
        const int tck = tck_env ? tck_env : tc_kernel;
        const bool tc_v2 = tck == 2 && tc2_tile_queries(dpad) > 0;
        const bool tc_v3 = tck == 3 && tc3_tile_queries(dpad) >= 32 && num_sms >= 2;  // CTA pairs (tcfilter3.cu)
        // d > 256: the resident query tile shrinks (32 queries at d = 960) and every list is re-streamed once per
        // tile; when lists are probed by many queries the tile is streamed through the ring instead (256 queries
        // per list pass, the tile re-read from L2 per 128-row block).  Pairs of this round: at most what the
        // active queries have left up to their bounds (counted by compact_active).
        const double est_pairs = std::min<double>((double)n_active * w, stats.rounds > 0 ? (double)rem_sum : 1e30);
        const bool stream_b = !tc_v2 && !tc_v3 && tc_stream_queries(dpad) &&
                              est_pairs / (double)std::min<long>(nlist, (long)n_active * w) >= (double)tc_stream_min;
        const int Ntc = tc_v2 ? tc2_tile_queries(dpad) : tc_v3 ? tc3_tile_queries(dpad) : tc_tile_queries(dpad, stream_b);
        if (use_tc) {
            S = 1;
            rp.qt = Ntc;
            rp.nsub = nsub = 1;
            rp.unsorted = 1;
        }
        rp.active = act_cur;
        rp.n_active = n_active;
        rp.r0 = r0;
        rp.w = (int)w;
        rp.S = (int)S;
        size_t slots = (size_t)n_active * w * S * nsub;
        // slot capacity: K everywhere, except in the first tensor-core round(s) whose threshold comes from
        // a single list -- there a pair can have a few hundred survivors, which merge_check reduces to
        // the K best (sort in shared memory, capacity 2 * KP) instead of an exact redo of the pair
        int KP = 16;
        while (KP < K) KP <<= 1;
        // (merge_check can order 2 KP entries itself -- the audit path --, slot_sort_kernel 512)
        const int cap = (use_tc && r0 < wide_slot_r0) ? std::max(K, tc_audit ? 2 * KP : std::min(4 * KP, 512)) : K;
        rp.cap = cap;
        pool.ensure(slots * cap * 8);
        rp.cand_d = reinterpret_cast<float*>(pool.p);
        rp.cand_off = reinterpret_cast<unsigned*>(pool.p + slots * cap * 4);
        rp.slot_cnt = slot_cnt.ensure(slots);
        rp.pairs = pairs.ensure((size_t)n_active * w);
        rp.xq_sorted = q_sorted.ensure(((size_t)n_active * w + 256) * dpad);
        alignas(64) unsigned char qmap[128];
        make_queries_tensor_map(qmap, rp.xq_sorted, (long long)n_active * w + 256, dpad);

        if (scan_ev.size() < 2 * (stats.rounds + 1)) {
            cudaEvent_t a, b;
            CUDA_CHECK(cudaEventCreate(&a));
            CUDA_CHECK(cudaEventCreate(&b));
            scan_ev.push_back(a);
            scan_ev.push_back(b);
        }
        if (partial_rank && r0 + (int)w + 1 > rank_rows_partial_width())
            // this round reads ranks the partial ranking did not produce: complete the rows that need them
            launch_rank_rows(metric, c_raw.p, n_active, nlist, c_dis.p, c_keys.p, c_tie0.p, stream, act_cur, c_sorted.p,
                             rp.st.bound, r0 + (int)w + 1);
        if (exact_ties && !ties_all_upfront && !ties_done) {
            // ranks [r0, r0+w) are about to be scanned: their order must be the reference's.  Once
            // the remaining queries fit one replay wave, fix all of their ranks and stop checking.
            // Error-bounded search: a query whose stop stage is known only needs the tie AT that stage.
            const bool all_now = n_active <= 1000 && qb.mode != 1;
            launch_fix_ties(metric, c_raw.p, nlist, nprobe, entry_table(nprobe), act_cur, n_active, c_tie0.p,
                            all_now ? nprobe : r0 + (int)w, rp.st.bound, fix_list.p, ctl.p + CTL_NFIX, c_dis.p,
                            c_keys.p, stream, qb.mode == 1 ? rp.st.decided : nullptr, r0, ctl.p + CTL_ERR,
                            partial_rank ? c_sorted.p : nullptr);
            ties_done = all_now;
        }
        CUDA_CHECK(cudaMemsetAsync(rp.round_work, 0, 4 * sizeof(unsigned long long), stream));
        launch_plan(rp, stream);
        CUDA_CHECK(cudaMemcpyAsync(h_round_work.p, rp.round_work, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaEventRecord(scan_ev[2 * stats.rounds], stream));
        bool scanned = false;
        int tc_idx = -1;
        size_t nredo_round = 0;
        if (use_tc) {
            TcArgs ta;
            ta.vnorm = vnorm.p;
            ta.qnorm = qnorm.p;
            ta.list_nmax = list_nmax.p;
            ta.c1 = 2.f * (1.02f / 512.f + (float)dpad / 2097152.f);
            ta.c2 = 1.f / 1048576.f;
            ta.c3 = 1.f / 16384.f;
            ta.cand_cap = (int)std::min<size_t>((size_t)64 << 20, std::max<size_t>((size_t)n_active * w * 32, 1 << 20));
            ta.cand = tc_cand.ensure(ta.cand_cap);
            ta.N = Ntc;
            ta.dry = 0;
            ta.stream_b = stream_b ? 1 : 0;
            alignas(64) unsigned char bmap[128];
            make_queries_tensor_map_tc(bmap, rp.xq_sorted, (long long)n_active * w + 256, dpad, tc_v2 ? 64 : tc_v3 ? Ntc / 2 : Ntc);
            CUDA_CHECK(cudaMemsetAsync(ctl.p + CTL_NCAND, 0, 2 * sizeof(int), stream));  // NCAND, OVERFLOW
            rp.pair_flag = pair_flag.ensure((size_t)n_active * w);
            CUDA_CHECK(cudaMemsetAsync(rp.pair_flag, 0, (size_t)n_active * w * sizeof(int), stream));
            tc_idx = (int)stats.tc_rounds + (int)stats.tc_fallbacks;
            if (tc_ev.size() < 2 * (size_t)(tc_idx + 1)) {
                cudaEvent_t a, b;
                CUDA_CHECK(cudaEventCreate(&a));
                CUDA_CHECK(cudaEventCreate(&b));
                tc_ev.push_back(a);
                tc_ev.push_back(b);
            }
            CUDA_CHECK(cudaEventRecord(tc_ev[2 * tc_idx], stream));
            if (tc_v2)
                launch_tc_filter2(rp, ta, codes_tmap64, bmap, num_sms, stream);
            else if (tc_v3) {
                // experiment (AUNCEL_TC_DRY=1|2): the same launch first without its epilogue / without the appends,
                // on the round's real inputs -- its duration shows in an ncu launch list
                static const int dry_env = getenv("AUNCEL_TC_DRY") ? atoi(getenv("AUNCEL_TC_DRY")) : 0;
                if (dry_env) {
                    TcArgs td = ta;
                    td.dry = dry_env;
                    launch_tc_filter3(rp, td, codes_tmap, bmap, num_sms, stream);
                    CUDA_CHECK(cudaMemsetAsync(ctl.p + CTL_TILE_COUNTER, 0, sizeof(int), stream));
                }
                launch_tc_filter3(rp, ta, codes_tmap, bmap, num_sms, stream);
            }
            else
                launch_tc_filter(rp, ta, codes_tmap, bmap, num_sms, stream);
            CUDA_CHECK(cudaEventRecord(tc_ev[2 * tc_idx + 1], stream));
            launch_rerank(rp, ta, num_sms, stream);
            CUDA_CHECK(cudaMemcpyAsync(h_ctl.p, ctl.p, CTL_SIZE * sizeof(int), cudaMemcpyDeviceToHost, stream));
            if (!tc_audit) launch_slot_sort(rp, num_sms, stream);  // (the audit compares the raw survivor sets)
            CUDA_CHECK(cudaStreamSynchronize(stream));
            launches += 2;
            stats.tc_candidates += (uint64_t)h_ctl.p[CTL_NCAND];
            if (h_ctl.p[CTL_OVERFLOW] < 0) {
                // the survivor list itself overflowed: redo the whole round with the exact scan
                stats.tc_fallbacks++;
                tc_idx = -1;
                rp.qt = SCAN_QT;
                rp.unsorted = 0;
                rp.filtered = 0;
                CUDA_CHECK(cudaMemsetAsync(rp.round_work, 0, 4 * sizeof(unsigned long long), stream));
                launch_plan(rp, stream);
                CUDA_CHECK(cudaMemcpyAsync(h_round_work.p, rp.round_work, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
            } else {
                stats.tc_rounds++;
                scanned = true;
                if (tc_audit) {
                    // exact rescan of the whole round into a second pool, slot-by-slot comparison
                    RoundParams ra = rp;
                    ra.cap = K;
                    ra.qt = SCAN_QT;
                    ra.unsorted = 0;
                    ra.defer_sort = 0;
                    ra.pair_flag = nullptr;
                    const size_t aslots = (size_t)n_active * w;
                    audit_pool.ensure(aslots * K * 8);
                    ra.cand_d = reinterpret_cast<float*>(audit_pool.p);
                    ra.cand_off = reinterpret_cast<unsigned*>(audit_pool.p + aslots * K * 4);
                    ra.slot_cnt = audit_cnt.ensure(aslots);
                    ra.round_work = audit_ctr.ensure(8) + 4;
                    launch_plan(ra, stream);
                    launch_scan(ra, codes_tmap, qmap, num_sms, stream);
                    launch_tc_audit(rp, ra.cand_d, ra.cand_off, ra.slot_cnt, audit_ctr.p, stream);
                }
                if (h_ctl.p[CTL_OVERFLOW] > 0) {
                    // some (query, list) pairs had more than K survivors (loose or missing tau):
                    // the exact scan redoes just those pairs and rewrites their slots
                    // h_ctl[CTL_OVERFLOW] counts overflowing survivors, an upper bound of the flagged pairs
                    const size_t nredo = std::min<size_t>((size_t)h_ctl.p[CTL_OVERFLOW], (size_t)n_active * w);
                    redo_pool.ensure(nredo * 4 * K * 8);
                    RoundParams rr = rp;
                    rr.filtered = 1;
                    rr.nc = nc4 ? 4 : 8;
                    rr.qt = rr.nc;  // narrow tiles: 4 * nc / nsub queries, the 128 rows of a block split over 4 warps
                    rr.nsub = 4;
                    rr.S = 1;
                    rr.cap = K;
                    rr.defer_sort = 0;
                    rr.cand_d = reinterpret_cast<float*>(redo_pool.p);
                    rr.cand_off = reinterpret_cast<unsigned*>(redo_pool.p + nredo * 4 * K * 4);
                    rr.slot_cnt = redo_cnt.ensure(nredo * 4);
                    rr.redo_ord = redo_ord.ensure((size_t)n_active * w);
                    CUDA_CHECK(cudaMemsetAsync(rr.slot_cnt, 0, nredo * 4 * sizeof(int), stream));
                    launch_plan(rr, stream);
                    launch_scan(rr, codes_tmap, qmap, num_sms, stream);
                    rp.redo_ord = rr.redo_ord;
                    rp.redo_d = rr.cand_d;
                    rp.redo_off = rr.cand_off;
                    rp.redo_cnt = rr.slot_cnt;
                    nredo_round = nredo;
                    launches += 5;
                }
            }
        }
        if (!scanned) {
            launch_scan(rp, codes_tmap, qmap, num_sms, stream);
            // few queries: many partial results per pair (segments x row subsets) -- merge them in parallel
            // before the per-query sequential pass
            if (S * nsub > 1 && !rp.defer_sort && (long)n_active * w <= 16384) {
                launch_stage_merge(rp, num_sms, stream);
                rp.merged = 1;
            }
        }
        CUDA_CHECK(cudaEventRecord(scan_ev[2 * stats.rounds + 1], stream));
        if (sharded) exchange_candidates(rp, nredo_round);  // rp now describes the union of all shards' candidates
        launch_merge_check(rp, tp, stream);
        launches += 7 + (exact_ties && !ties_all_upfront ? 2 : 0);  // [collect_ties, heap_order,] plan x3, gather, scan, merge_check, compact_active
        launch_compact_active(rp, r0 + (int)w, act_nxt, h_ctl.p, stream);
        CUDA_CHECK(cudaStreamSynchronize(stream));
        tc_round_of.push_back(tc_idx);
        round_meta.push_back({(double)r0, (double)w, (double)rp.n_active});
        round_ndis.push_back(h_round_work.p[0]);
        round_uniq.push_back(h_round_work.p[1]);
        round_staged.push_back(h_round_work.p[2]);
        n_active = h_ctl.p[CTL_N_ACTIVE];
        min_rcnt = h_ctl.p[CTL_MIN_RCNT];
        not_full = h_ctl.p[CTL_NOT_FULL];
        rem_sum = h_ctl.p[CTL_REM_SUM];
        if (debug_rounds)
            round_log.push_back({r0, (int)w, rp.unsorted ? -1 : (int)(S * nsub), rp.n_active, h_ctl.p[CTL_TOTAL_TILES],
                                 h_ctl.p[CTL_TOTAL_PAIRS]});
        stats.rounds++;
        stats.scan_tiles += (uint64_t)h_ctl.p[CTL_TOTAL_TILES];
        stats.scan_pairs += (uint64_t)h_ctl.p[CTL_TOTAL_PAIRS];
        std::swap(act_cur, act_nxt);
        r0 += (int)w;
    }

    // ---- results
    DevBuf<unsigned long long>& st_dev = io_u;
    st_dev.ensure(4);
    CUDA_CHECK(cudaMemsetAsync(st_dev.p, 0, 4 * sizeof(unsigned long long), stream));
    launch_finalize(rp, tp, qb.D, qb.I, qb.my_nprobe, st_dev.p, stream);
    if (sharded) shard_x->all_reduce_max(qb.I, (size_t)n * K, stream);  // every label from the shard that holds the vector
    unsigned long long h_st[2];
    CUDA_CHECK(cudaMemcpyAsync(h_st, st_dev.p, sizeof(h_st), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaMemcpyAsync(h_ctl.p, ctl.p, CTL_SIZE * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaEventRecord(ev1, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    stats.nq = n;
    stats.nlist = h_st[0];
    stats.ndis = h_st[1];
    stats.search_ms = ms;
    for (uint64_t r = 0; r < stats.rounds; r++) {
        float t = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&t, scan_ev[2 * r], scan_ev[2 * r + 1]));
        scan_ms_total += t;
        float tt = 0.f;
        if (tc_round_of[r] >= 0)
            CUDA_CHECK(cudaEventElapsedTime(&tt, tc_ev[2 * tc_round_of[r]], tc_ev[2 * tc_round_of[r] + 1]));
        round_stats.push_back({round_meta[r][0], round_meta[r][1], round_meta[r][2], tc_round_of[r] >= 0 ? 1.0 : 0.0,
                               (double)round_ndis[r], (double)round_uniq[r], (double)round_staged[r], (double)t, (double)tt, 0.0});
        if (tc_round_of[r] >= 0) {
            stats.tc_ms += tt;
            stats.tc_ndis += round_ndis[r];
            stats.tc_uniq += round_uniq[r];
            stats.tc_staged += round_staged[r];
        } else {
            stats.simt_ms += t;
            stats.simt_ndis += round_ndis[r];
            stats.simt_uniq += round_uniq[r];
            stats.simt_staged += round_staged[r];
        }
        if (debug_rounds && r < round_log.size())
            fprintf(stderr, "[auncel] round %2d r0=%4d w=%4d S=%2d active=%6d tiles=%7d pairs=%8d scan=%8.3f ms\n", (int)r,
                    round_log[r][0], round_log[r][1], round_log[r][2], round_log[r][3], round_log[r][4],
                    round_log[r][5], t);
    }
    stats.scan_ms = scan_ms_total;
    stats.scan_launches = stats.rounds;
    stats.launches = launches;
    float cms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&cms, ev0, ev2));
    stats.coarse_ms = cms;
    stats.err_bits = (uint64_t)h_ctl.p[CTL_ERR];
    if (tc_audit) {
        unsigned long long h_a[3];
        CUDA_CHECK(cudaMemcpy(h_a, audit_ctr.p, sizeof(h_a), cudaMemcpyDeviceToHost));
        stats.tc_audit_bad = h_a[0];
        stats.tc_audit_slots = h_a[1];
        stats.tc_audit_cands = h_a[2];
    }
}

}  // namespace auncel
