// Host-side C++ mirror of the reference's Index API for the error-bounded IVF-Flat path,
// header-only, above the C ABI (include/auncel_b200.h).  Same names, argument meaning and
// error behaviour as /root/reference/Auncel so that the reference's drivers
// (eval/bound.cpp, eval/effect_error.cpp, dist/worker.cpp) read the same against it:
//
//   faiss::Index            Index.h:66-210          faiss::IndexFlatL2 / IndexFlatIP  IndexFlat.h
//   faiss::IndexIVF / IndexIVFFlat   IndexIVF.h:97-308, IndexIVFFlat.h:24-59
//   faiss::error_pro        IVF_pro.h:77-175 (the fields drivers touch)
//   faiss::Error_sys        profile.h:29-91
//   faiss::IndexShards / IndexReplicas   IndexShards.h, IndexReplicas.h
//
// Everything numeric happens behind the C ABI on the GPU; this file is bookkeeping.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../auncel_b200.h"

namespace faiss {

/// FaissException.h
class FaissException : public std::exception {
   public:
    explicit FaissException(const std::string& m) : msg(m) {}
    const char* what() const noexcept override { return msg.c_str(); }
    std::string msg;
};

#define AUNCEL_FAISS_THROW_IF_NOT_MSG(X, MSG) \
    do { if (!(X)) throw ::faiss::FaissException(std::string("Error: '" #X "' failed: ") + (MSG)); } while (0)

inline void auncel_check(int rc) {
    if (rc != 0) throw FaissException(auncel_get_last_error());
}

enum IndexType { IVF = 0, NSW = 1, OTHER = 2 };           // Index.h:42-46
enum MetricType { METRIC_INNER_PRODUCT = 0, METRIC_L2 = 1 };  // Index.h:48-51

struct Index {  // Index.h:66-210
    using idx_t = long;
    using component_t = float;
    using distance_t = float;
    bool tune = false;
    IndexType type = OTHER;
    int d;
    idx_t ntotal = 0;
    bool verbose = false;
    bool is_trained = true;
    MetricType metric_type;

    explicit Index(idx_t d = 0, MetricType metric = METRIC_L2) : d((int)d), metric_type(metric) {}
    virtual ~Index() {}
    virtual void set_tune_mode() { tune = true; }
    virtual void set_tune_off() { tune = false; }
    virtual void train(idx_t /*n*/, const float* /*x*/) {}
    virtual void add(idx_t n, const float* x) = 0;
    virtual void add_with_ids(idx_t, const float*, const long*) { throw FaissException("add_with_ids not implemented for this type of index"); }
    virtual void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const = 0;
    virtual void reset() = 0;
    void assign(idx_t n, const float* x, idx_t* labels, idx_t k = 1) {  // Index.cpp:42-47
        std::vector<float> dis(n * k);
        search(n, x, k, dis.data(), labels);
    }
};

/// The coarse quantizer object (IndexFlat.h): stores the centroids; search ranks them on the GPU.
struct IndexFlat : Index {
    std::vector<float> xb;
    int device = 0;
    mutable AuncelIndex* h = nullptr;
    explicit IndexFlat(idx_t d, MetricType metric = METRIC_L2) : Index(d, metric) {}
    ~IndexFlat() override { if (h) auncel_index_free(h); }
    void add(idx_t n, const float* x) override {
        xb.insert(xb.end(), x, x + n * d);
        ntotal += n;
        drop();
    }
    void reset() override { xb.clear(); ntotal = 0; drop(); }
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(ntotal > 0, "empty flat index");
        if (!h) {
            auncel_check(auncel_index_new(&h, d, ntotal, (int)metric_type, device));
            auncel_check(auncel_index_set_centroids(h, xb.data(), 0));
        }
        idx_t kk = std::min<idx_t>(k, ntotal);
        std::vector<float> dd((size_t)n * kk);
        std::vector<int64_t> ll((size_t)n * kk);
        auncel_check(auncel_index_coarse_search(h, n, x, kk, dd.data(), ll.data()));
        for (idx_t i = 0; i < n; i++)
            for (idx_t j = 0; j < k; j++) {  // pad like heap_reorder, Heap.h:316-319
                distances[i * k + j] = j < kk ? dd[i * kk + j] : (metric_type == METRIC_L2 ? 3.402823466e+38f : -3.402823466e+38f);
                labels[i * k + j] = j < kk ? (idx_t)ll[i * kk + j] : -1;
            }
    }
   private:
    void drop() { if (h) { auncel_index_free(h); h = nullptr; } }
};
struct IndexFlatL2 : IndexFlat { explicit IndexFlatL2(idx_t d) : IndexFlat(d, METRIC_L2) {} };
struct IndexFlatIP : IndexFlat { explicit IndexFlatIP(idx_t d) : IndexFlat(d, METRIC_INNER_PRODUCT) {} };

/// IVF_pro.h:77-175 -- the members the drivers read and write
class error_pro {
   public:
    float std_m = 1.0f, multipler = 1.0f;
    size_t arcos_size = 500;
    const float* require_acc = nullptr;
    bool profile = false, overhead_profile = false, time_tune = false;
    size_t alloc_s = 0;
    float* t_recalls = nullptr;
    size_t* my_nprobe = nullptr;
    size_t query_topk = (size_t)-1;
    size_t nlist = 0, max_topk = 0, d = 0, train_num = 0;
    const float* train_D = nullptr;
    /// IVF_pro.cpp:240-256: line `id` of ../hyperparameter.txt holds (multipler, std_m)
    void setparam(int id, const char* fn = "../hyperparameter.txt") {
        std::ifstream infile(fn);
        AUNCEL_FAISS_THROW_IF_NOT_MSG(infile.good(), "cannot open hyperparameter file");
        for (int i = 0; i < 12; i++) {
            float a, b;
            infile >> a >> b;
            if (i == id - 1) { multipler = a; std_m = b; }
        }
        profile = false;
    }
    ~error_pro() { delete[] my_nprobe; delete[] t_recalls; }
};

/// AuxIndexStructures.h:31-50
struct RangeSearchResult {
    using idx_t = long;
    size_t nq;
    size_t* lims;       ///< size nq + 1: result of query i is [lims[i], lims[i+1])
    idx_t* labels = nullptr;
    float* distances = nullptr;  ///< not sorted
    size_t buffer_size = 1024 * 256;
    explicit RangeSearchResult(idx_t nq, bool alloc_lims = true) : nq((size_t)nq) {
        lims = alloc_lims ? new size_t[nq + 1]() : nullptr;
    }
    /// called when lims holds the offsets (AuxIndexStructures.cpp:39-45)
    virtual void do_allocation() {
        size_t ofs = lims[nq];
        labels = new idx_t[ofs];
        distances = new float[ofs];
    }
    virtual ~RangeSearchResult() {
        delete[] labels;
        delete[] distances;
        delete[] lims;
    }
};

struct IndexIVF : Index {  // IndexIVF.h:97-308
    Index* quantizer;
    size_t nlist;
    bool own_fields = false;
    bool training = false;
    error_pro* t = nullptr;
    size_t nprobe = 1, max_codes = 0;
    int device = 0;
    int niter = 25;  // cp.niter, IndexIVF.cpp:54
    AuncelIndex* h = nullptr;

    IndexIVF(Index* quantizer, size_t d, size_t nlist, MetricType metric, int device = 0)
        : Index(d, metric), quantizer(quantizer), nlist(nlist), device(device) {
        AUNCEL_FAISS_THROW_IF_NOT_MSG((int)d == quantizer->d, "quantizer dimension mismatch");  // IndexIVF.cpp:155
        type = IVF;
        is_trained = quantizer->is_trained && quantizer->ntotal == (idx_t)nlist;
        auncel_check(auncel_index_new(&h, (int)d, (int64_t)nlist, (int)metric, device));
        if (is_trained) import_quantizer(false);
    }
    ~IndexIVF() override {
        if (h) auncel_index_free(h);
        if (own_fields) delete quantizer;
        delete t;
    }
    void set_tune_mode() override { tune = true; quantizer->tune = true; }   // IndexIVF.cpp:179-182
    void set_tune_off() override { tune = false; quantizer->tune = false; }
    void set_train_mode() { training = true; quantizer->tune = true; }
    void set_train_off() { training = false; quantizer->tune = false; }

    /// IndexIVF::train -> Level1Quantizer::train_q1 (IndexIVF.cpp:71-137,995-1008)
    void train(idx_t n, const float* x) override {
        IndexFlat* fq = dynamic_cast<IndexFlat*>(quantizer);
        AUNCEL_FAISS_THROW_IF_NOT_MSG(fq != nullptr, "the quantizer must be an IndexFlat");
        if (quantizer->is_trained && quantizer->ntotal == (idx_t)nlist) {
            import_quantizer(quantizer->tune);
        } else {
            auncel_check(auncel_index_train(h, n, x, niter, quantizer->tune ? 1 : 0));
            std::vector<float> c(nlist * d);
            auncel_check(auncel_index_get_centroids(h, c.data()));
            fq->reset();
            fq->add(nlist, c.data());
            quantizer->is_trained = true;
        }
        is_trained = true;
    }
    void add(idx_t n, const float* x) override { add_with_ids(n, x, nullptr); }
    void reset() override { auncel_check(auncel_index_reset(h)); ntotal = 0; }

    /// 5-argument search, IndexIVF.cpp:335-353
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override {
        std::vector<int64_t> ll((size_t)n * k);
        auncel_check(auncel_index_search(h, n, x, k, (int64_t)nprobe, (int64_t)max_codes, distances, ll.data()));
        for (size_t i = 0; i < ll.size(); i++) labels[i] = (idx_t)ll[i];
    }
    /// 6-argument search with the global id of query 0, IndexIVF.cpp:355-378: the Auncel path
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels, size_t offset) const {
        if (!tune) { search(n, x, k, distances, labels); return; }
        AUNCEL_FAISS_THROW_IF_NOT_MSG(t != nullptr, "Search tune start can't start without IVF_pro init and training");
        AUNCEL_FAISS_THROW_IF_NOT_MSG(t->my_nprobe != nullptr && t->require_acc != nullptr, "set_queries was not called");
        AUNCEL_FAISS_THROW_IF_NOT_MSG(offset + n <= t->alloc_s, "query ids exceed the allocated range");
        auncel_check(auncel_index_set_params(h, t->multipler, t->std_m));
        std::vector<int64_t> ll((size_t)n * k);
        std::vector<uint64_t> np(t->my_nprobe + offset, t->my_nprobe + offset + n);
        std::vector<float> kth;
        if (t->train_D) {
            kth.resize(n);
            for (idx_t i = 0; i < n; i++) kth[i] = t->train_D[(offset + i) * k + t->query_topk - 1];  // IndexIVF.cpp:509
        }
        int flags = (t->profile ? 1 : 0) | (t->overhead_profile ? 2 : 0) | (t->time_tune ? 4 : 0);
        auncel_check(auncel_index_search_bounded(h, n, x, k, (int64_t)t->query_topk, t->require_acc + offset,
                                                 kth.empty() ? nullptr : kth.data(), np.data(),
                                                 t->t_recalls ? t->t_recalls + offset : nullptr, flags, distances,
                                                 ll.data()));
        for (idx_t i = 0; i < n; i++) t->my_nprobe[offset + i] = (size_t)np[i];
        for (size_t i = 0; i < ll.size(); i++) labels[i] = (idx_t)ll[i];
    }

    /// IndexIVF::range_search, IndexIVF.cpp:741-860
    void range_search(idx_t nx, const float* x, float radius, RangeSearchResult* result) const {
        std::vector<int64_t> lims((size_t)nx + 1);
        auncel_check(auncel_index_range_search(h, nx, x, radius, (int64_t)nprobe, lims.data()));
        for (idx_t i = 0; i <= nx; i++) result->lims[i] = (size_t)lims[i];
        result->do_allocation();
        static_assert(sizeof(idx_t) == sizeof(int64_t), "idx_t must be 64-bit");
        auncel_check(auncel_index_range_search_results(h, result->distances, (int64_t*)result->labels));
    }
    /// the search of Error_sys::time_search (profile.cpp:229-244): latency budget per query in
    /// t->require_acc (ms), tune block off, error_pro::time_tune cut (IndexIVF.cpp:545-549)
    void search_timed(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels, size_t offset) const {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(t != nullptr && t->require_acc != nullptr, "set_queries was not called");
        std::vector<int64_t> ll((size_t)n * k);
        auncel_check(auncel_index_search_timed(h, n, x, k, t->require_acc + offset, distances, ll.data()));
        for (size_t i = 0; i < ll.size(); i++) labels[i] = (idx_t)ll[i];
    }

   protected:
    void import_quantizer(bool with_interdis) {
        IndexFlat* fq = dynamic_cast<IndexFlat*>(quantizer);
        AUNCEL_FAISS_THROW_IF_NOT_MSG(fq != nullptr, "the quantizer must be an IndexFlat");
        auncel_check(auncel_index_set_centroids(h, fq->xb.data(), with_interdis ? 1 : 0));
    }
};

struct IndexIVFFlat : IndexIVF {  // IndexIVFFlat.h:24-59
    IndexIVFFlat(Index* quantizer, size_t d, size_t nlist_, MetricType metric = METRIC_L2, int device = 0)
        : IndexIVF(quantizer, d, nlist_, metric, device) {}
    void add_with_ids(idx_t n, const float* x, const long* xids) override { add_core(n, x, xids, nullptr); }
    /// IndexIVFFlat.cpp:41-80
    virtual void add_core(idx_t n, const float* x, const long* xids, const long* precomputed_idx) {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(is_trained, "is_trained");
        static_assert(sizeof(long) == sizeof(int64_t), "idx_t must be 64-bit");
        auncel_check(auncel_index_add(h, n, x, (const int64_t*)xids, (const int64_t*)precomputed_idx));
        ntotal += n;
    }
};

/// Error_sys, profile.h:29-91 / profile.cpp
class Error_sys {
   public:
    const float* queries = nullptr;
    size_t num = 0;
    const float* require_acc = nullptr;
    bool is_trained = false;
    std::string key = "Base";
    size_t train_num, max_topk;
    IndexIVF* index = nullptr;
    std::vector<float> train_D;
    std::vector<Index::idx_t> train_I;

    Error_sys(Index* in, size_t nq, size_t topk) : train_num(nq), max_topk(topk) {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(nq % 10 == 0, "Train num must be evenly divided by ten");  // profile.cpp:31-32
        if (IndexIVF* ix = dynamic_cast<IndexIVF*>(in)) {
            index = ix;
            key = "IVF";
        }
    }
    void set_gt(const float* gt_D_in, const Index::idx_t* gt_I_in) {  // profile.cpp:44-54
        AUNCEL_FAISS_THROW_IF_NOT_MSG(gt_D_in != nullptr && gt_I_in != nullptr,
                                      "the ground truth must not be null ptr when setting up");
        train_D.assign(gt_D_in, gt_D_in + train_num * max_topk);
        train_I.assign(gt_I_in, gt_I_in + train_num * max_topk);
    }
    /// profile.cpp:88-171: init_tune + calibration search + error_pro::train
    void sys_train(size_t nq, const float* xq) {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(nq <= train_num,
                                      "Error sys training does not have the same nb of queries compared with creation");
        AUNCEL_FAISS_THROW_IF_NOT_MSG(train_I.size() == train_num * max_topk, "ground truth not initialized");
        AUNCEL_FAISS_THROW_IF_NOT_MSG(index != nullptr, "Error_sys needs an IndexIVF");
        delete index->t;
        index->t = new error_pro;  // IndexIVF::init_tune, IndexIVF.cpp:203-244
        index->t->nlist = index->nlist;
        index->t->max_topk = max_topk;
        index->t->d = index->d;
        index->t->train_num = nq;
        index->t->train_D = train_D.data();
        auncel_check(auncel_index_calibrate(index->h, (int64_t)nq, xq, (int64_t)max_topk, train_D.data(), nullptr, nullptr));
        is_trained = true;
    }
    void set_queries(size_t n, const float* q, const float* acc, size_t allo_size) {  // profile.cpp:173-202
        num = n;
        queries = q;
        require_acc = acc;
        error_pro* t = index->t;
        AUNCEL_FAISS_THROW_IF_NOT_MSG(t != nullptr, "your must init tune for index first");
        t->alloc_s = allo_size;
        delete[] t->my_nprobe;
        t->my_nprobe = new size_t[allo_size]();
        delete[] t->t_recalls;
        t->t_recalls = new float[allo_size]();
        t->require_acc = acc;
    }
    void set_topk(size_t new_topk) { index->t->query_topk = new_topk; }  // profile.cpp:204-209
    /// profile.cpp:211-227
    void search(float* D, int64_t* I, size_t start, size_t search_size = (size_t)-1) {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(is_trained == true, "Error sys must be trained before searching");
        AUNCEL_FAISS_THROW_IF_NOT_MSG(num <= train_num, "Error sys search num must be lower than all qeuries num");
        index->set_tune_mode();
        index->nprobe = index->nlist;
        size_t n = search_size == (size_t)-1 ? num : search_size;
        index->search((Index::idx_t)n, queries + start * index->d, (Index::idx_t)max_topk, D, (Index::idx_t*)I, start);
        index->set_tune_off();
    }
    /// profile.cpp:229-244.  Like the reference (:242) the flag error_pro::time_tune stays set.
    void time_search(float* D, int64_t* I, size_t start, size_t search_size = (size_t)-1) {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(is_trained == true, "Error sys must be trained before searching");
        AUNCEL_FAISS_THROW_IF_NOT_MSG(num <= train_num, "Error sys search num must be lower than all qeuries num");
        index->t->time_tune = true;
        index->nprobe = index->nlist;
        size_t n = search_size == (size_t)-1 ? num : search_size;
        index->search_timed((Index::idx_t)n, queries + start * index->d, (Index::idx_t)max_topk, D, (Index::idx_t*)I, start);
        index->t->time_tune = true;
    }
};

/// ThreadedIndex / IndexShards / IndexReplicas (ThreadedIndex-inl.h:119-190, IndexShards.cpp, IndexReplicas.cpp):
/// one host thread per sub-index (one GPU each); per-shard exceptions are collected and rethrown.
struct ThreadedIndexBase : Index {
    std::vector<Index*> indices;
    bool threaded;
    ThreadedIndexBase(int d, bool threaded) : Index(d), threaded(threaded) {}
    void addIndex(Index* index) {
        if (indices.empty() && d == 0) d = index->d;
        AUNCEL_FAISS_THROW_IF_NOT_MSG(index->d == d, "addIndex: dimension mismatch for newly added index");
        if (indices.empty()) metric_type = index->metric_type;
        indices.push_back(index);
    }
    int count() const { return (int)indices.size(); }
    Index* at(int i) const { return indices[i]; }
    template <class F>
    void runOnIndex(F f) const {
        std::vector<std::string> errs(indices.size());
        auto body = [&](int i) {
            try { f(i, indices[i]); } catch (const std::exception& e) { errs[i] = e.what(); if (errs[i].empty()) errs[i] = "?"; }
        };
        if (threaded) {
            std::vector<std::thread> th;
            for (int i = 0; i < count(); i++) th.emplace_back(body, i);
            for (auto& t : th) t.join();
        } else {
            for (int i = 0; i < count(); i++) body(i);
        }
        std::string all;
        for (int i = 0; i < count(); i++)
            if (!errs[i].empty()) all += "Exception thrown from index " + std::to_string(i) + ": " + errs[i] + "\n";
        if (!all.empty()) throw FaissException(all);
    }
    void reset() override { runOnIndex([](int, Index* ix) { ix->reset(); }); ntotal = 0; }
};

struct IndexShards : ThreadedIndexBase {
    bool successive_ids;
    explicit IndexShards(int d, bool threaded = false, bool successive_ids = true)
        : ThreadedIndexBase(d, threaded), successive_ids(successive_ids) {}
    void add_shard(Index* index) { addIndex(index); }
    void train(idx_t n, const float* x) override { runOnIndex([=](int, Index* ix) { ix->train(n, x); }); }
    void add(idx_t n, const float* x) override { add_with_ids(n, x, nullptr); }
    /// IndexShards.cpp:196-259: contiguous split i0 = no*n/nshard
    void add_with_ids(idx_t n, const float* x, const long* xids) override {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(!(successive_ids && xids), "It makes no sense to pass in ids and request them to be shifted");
        if (successive_ids)
            AUNCEL_FAISS_THROW_IF_NOT_MSG(ntotal == 0, "when adding to IndexShards with sucessive_ids, only add() in a single pass is supported");
        idx_t nshard = count();
        std::vector<long> aids;
        const long* ids = xids;
        if (!ids && !successive_ids) {
            aids.resize(n);
            for (idx_t i = 0; i < n; i++) aids[i] = ntotal + i;
            ids = aids.data();
        }
        int dd = d;
        runOnIndex([=](int no, Index* ix) {
            idx_t i0 = (idx_t)no * n / nshard, i1 = ((idx_t)no + 1) * n / nshard;
            if (ids) ix->add_with_ids(i1 - i0, x + i0 * dd, ids + i0);
            else ix->add(i1 - i0, x + i0 * dd);
        });
        ntotal += n;
    }
    /// IndexShards.cpp:261-311: every shard answers all queries, then merge_tables
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override {
        long nshard = count();
        std::vector<float> all_d((size_t)nshard * k * n);
        std::vector<idx_t> all_l((size_t)nshard * k * n);
        runOnIndex([&](int no, Index* ix) { ix->search(n, x, k, all_d.data() + (size_t)no * k * n, all_l.data() + (size_t)no * k * n); });
        std::vector<int64_t> tr(nshard, 0);
        if (successive_ids)
            for (int s = 0; s + 1 < nshard; s++) tr[s + 1] = tr[s] + at(s)->ntotal;
        std::vector<int64_t> out((size_t)n * k);
        auncel_check(auncel_merge_tables((int)metric_type, n, k, nshard, all_d.data(), (const int64_t*)all_l.data(), tr.data(),
                                         distances, out.data()));
        for (size_t i = 0; i < out.size(); i++) labels[i] = (idx_t)out[i];
    }
};

struct IndexReplicas : ThreadedIndexBase {
    explicit IndexReplicas(int d, bool threaded = true) : ThreadedIndexBase(d, threaded) {}
    void addReplica(Index* index) { addIndex(index); }
    void train(idx_t n, const float* x) override { runOnIndex([=](int, Index* ix) { ix->train(n, x); }); }
    void add(idx_t n, const float* x) override { runOnIndex([=](int, Index* ix) { ix->add(n, x); }); ntotal += n; }
    /// IndexReplicas.cpp:79-118: queries split in contiguous chunks of ceil(n / count)
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(count() > 0, "no replicas in index");
        if (n == 0) return;
        idx_t per = (n + count() - 1) / count();
        int dd = d;
        runOnIndex([=](int i, Index* ix) {
            idx_t base = (idx_t)i * per;
            if (base < n) ix->search(std::min(per, n - base), x + base * dd, k, distances + base * k, labels + base * k);
        });
    }
};

/// ParameterSpace::set_index_parameters (AutoTune.cpp:455-563) for the IVF parameters of this path:
/// a comma-separated "nprobe=16,max_codes=0" string.
inline void set_index_parameters(Index* index, const char* description) {
    IndexIVF* ix = dynamic_cast<IndexIVF*>(index);
    AUNCEL_FAISS_THROW_IF_NOT_MSG(ix != nullptr, "set_index_parameters: not an IndexIVF");
    std::string s(description);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t end = s.find(',', pos);
        if (end == std::string::npos) end = s.size();
        std::string tok = s.substr(pos, end - pos);
        size_t eq = tok.find('=');
        AUNCEL_FAISS_THROW_IF_NOT_MSG(eq != std::string::npos, "could not parse parameter " + tok);
        std::string name = tok.substr(0, eq);
        double val = std::atof(tok.c_str() + eq + 1);
        if (name == "nprobe") ix->nprobe = (size_t)val;
        else if (name == "max_codes") ix->max_codes = std::isfinite(val) ? (size_t)val : 0;
        else throw FaissException("ParameterSpace::set_index_parameter: could not set parameter " + name);
        pos = end + 1;
    }
}

/// index_factory (AutoTune.cpp:741-852) for the one description this path uses: "IVF<nlist>,Flat"
/// (eval/bound.cpp:220).  The returned IndexIVFFlat owns its IndexFlat quantizer.
inline Index* index_factory(int d, const char* description, MetricType metric = METRIC_L2) {
    int nlist = 0, consumed = 0;
    if (std::sscanf(description, "IVF%d,Flat%n", &nlist, &consumed) != 1 || description[consumed] != '\0' || nlist <= 0)
        throw FaissException(std::string("index_factory: could not parse the description ") + description);
    Index* quantizer = metric == METRIC_L2 ? (Index*)new IndexFlatL2(d) : (Index*)new IndexFlatIP(d);
    IndexIVFFlat* ix = new IndexIVFFlat(quantizer, d, nlist, metric);
    ix->own_fields = true;
    return ix;
}

/// write_index / read_index (index_io.cpp:383-460, 890+) for IndexIVFFlat: the reference's "IwFl" layout --
/// index header (:196-203), nlist, nprobe, the quantizer as "IxF2"/"IxFI" (:384-390), direct map,
/// ArrayInvertedLists "ilar" with a "full" or "sprs" size table (:280-330) -- so files are exchangeable
/// with the reference.  (auncel_b200/index_io.py additionally appends the Auncel state.)
namespace io_detail {
inline uint32_t fourcc(const char* s) { return (uint32_t)(unsigned char)s[0] | ((uint32_t)(unsigned char)s[1] << 8) | ((uint32_t)(unsigned char)s[2] << 16) | ((uint32_t)(unsigned char)s[3] << 24); }
template <class T> inline void put(FILE* f, const T& v) { if (std::fwrite(&v, sizeof(T), 1, f) != 1) throw FaissException("write error"); }
template <class T> inline void get(FILE* f, T& v) { if (std::fread(&v, sizeof(T), 1, f) != 1) throw FaissException("read error"); }
inline void header(FILE* f, int d, long ntotal, bool trained, int metric) {
    put(f, d); put(f, ntotal); long dummy = 1 << 20; put(f, dummy); put(f, dummy); put(f, trained); put(f, metric);
}
inline void rheader(FILE* f, int& d, long& ntotal, bool& trained, int& metric) {
    long dummy; get(f, d); get(f, ntotal); get(f, dummy); get(f, dummy); get(f, trained); get(f, metric);
}
}  // namespace io_detail

inline void write_index(const Index* idx, const char* fname) {
    using namespace io_detail;
    const IndexIVFFlat* ix = dynamic_cast<const IndexIVFFlat*>(idx);
    AUNCEL_FAISS_THROW_IF_NOT_MSG(ix != nullptr, "write_index: only IndexIVFFlat is supported on this path");
    const IndexFlat* fq = dynamic_cast<const IndexFlat*>(ix->quantizer);
    AUNCEL_FAISS_THROW_IF_NOT_MSG(fq != nullptr, "the quantizer must be an IndexFlat");
    FILE* f = std::fopen(fname, "wb");
    if (!f) throw FaissException(std::string("could not open ") + fname + " for writing");
    put(f, fourcc("IwFl"));
    header(f, ix->d, ix->ntotal, ix->is_trained, (int)ix->metric_type);
    put(f, (size_t)ix->nlist); put(f, (size_t)ix->nprobe);
    put(f, fourcc(ix->metric_type == METRIC_L2 ? "IxF2" : "IxFI"));
    header(f, fq->d, fq->ntotal, fq->is_trained, (int)fq->metric_type);
    put(f, (size_t)fq->xb.size());
    if (!fq->xb.empty() && std::fwrite(fq->xb.data(), sizeof(float), fq->xb.size(), f) != fq->xb.size()) throw FaissException("write error");
    put(f, false); put(f, (size_t)0);  // maintain_direct_map, direct_map
    std::vector<int64_t> sizes(ix->nlist);
    auncel_check(auncel_index_list_sizes(ix->h, sizes.data()));
    size_t tot = 0, nonempty = 0;
    for (auto s : sizes) { tot += (size_t)s; nonempty += s > 0; }
    std::vector<float> codes(tot * ix->d);
    std::vector<int64_t> ids(tot);
    auncel_check(auncel_index_get_lists(ix->h, codes.data(), ids.data()));
    put(f, fourcc("ilar")); put(f, (size_t)ix->nlist); put(f, (size_t)(sizeof(float) * ix->d));
    if (nonempty > ix->nlist / 2) {  // index_io.cpp:293-313
        put(f, fourcc("full")); put(f, (size_t)ix->nlist);
        for (auto s : sizes) put(f, (size_t)s);
    } else {
        put(f, fourcc("sprs")); put(f, (size_t)(2 * nonempty));
        for (size_t l = 0; l < ix->nlist; l++) if (sizes[l] > 0) { put(f, l); put(f, (size_t)sizes[l]); }
    }
    size_t off = 0;
    for (size_t l = 0; l < ix->nlist; l++) {
        size_t n = (size_t)sizes[l];
        if (n == 0) continue;
        if (std::fwrite(codes.data() + off * ix->d, sizeof(float), n * ix->d, f) != n * ix->d) throw FaissException("write error");
        if (std::fwrite(ids.data() + off, sizeof(int64_t), n, f) != n) throw FaissException("write error");
        off += n;
    }
    std::fclose(f);
}

inline Index* read_index(const char* fname, int device = 0) {
    using namespace io_detail;
    FILE* f = std::fopen(fname, "rb");
    if (!f) throw FaissException(std::string("could not open ") + fname + " for reading");
    uint32_t h; get(f, h);
    AUNCEL_FAISS_THROW_IF_NOT_MSG(h == fourcc("IwFl"), "read_index: only IndexIVFFlat files are supported on this path");
    int d, metric, dq, mq; long ntotal, nq; bool trained, tq;
    rheader(f, d, ntotal, trained, metric);
    size_t nlist, nprobe; get(f, nlist); get(f, nprobe);
    uint32_t hq; get(f, hq);
    AUNCEL_FAISS_THROW_IF_NOT_MSG(hq == fourcc("IxF2") || hq == fourcc("IxFI"), "quantizer is not an IndexFlat");
    rheader(f, dq, nq, tq, mq);
    size_t nx; get(f, nx);
    std::vector<float> xb(nx);
    if (nx && std::fread(xb.data(), sizeof(float), nx, f) != nx) throw FaissException("read error");
    bool mdm; get(f, mdm);
    size_t ndm; get(f, ndm);
    std::fseek(f, (long)(ndm * sizeof(long)), SEEK_CUR);
    IndexFlat* q = metric == (int)METRIC_L2 ? (IndexFlat*)new IndexFlatL2(d) : (IndexFlat*)new IndexFlatIP(d);
    if (nq > 0) q->add(nq, xb.data());
    IndexIVFFlat* ix = new IndexIVFFlat(q, d, nlist, (MetricType)metric, device);
    ix->own_fields = true;
    ix->nprobe = nprobe;
    uint32_t hl; get(f, hl);
    if (hl == fourcc("ilar")) {
        size_t nl, cs; get(f, nl); get(f, cs);
        AUNCEL_FAISS_THROW_IF_NOT_MSG(nl == nlist && cs == sizeof(float) * d, "inverted lists do not match the index");
        uint32_t lt; get(f, lt);
        size_t nt; get(f, nt);
        std::vector<size_t> tab(nt);
        if (nt && std::fread(tab.data(), sizeof(size_t), nt, f) != nt) throw FaissException("read error");
        std::vector<size_t> sizes(nlist, 0);
        if (lt == fourcc("full")) sizes = tab;
        else if (lt == fourcc("sprs")) for (size_t i = 0; i + 1 < nt; i += 2) sizes[tab[i]] = tab[i + 1];
        else throw FaissException("unknown list type");
        for (size_t l = 0; l < nlist; l++) {
            size_t n = sizes[l];
            if (n == 0) continue;
            std::vector<float> codes(n * d);
            std::vector<long> ids(n), ln(n, (long)l);
            if (std::fread(codes.data(), sizeof(float), n * d, f) != n * d) throw FaissException("read error");
            if (std::fread(ids.data(), sizeof(long), n, f) != n) throw FaissException("read error");
            ix->add_core((Index::idx_t)n, codes.data(), ids.data(), ln.data());
        }
        ix->ntotal = ntotal;
    } else {
        AUNCEL_FAISS_THROW_IF_NOT_MSG(hl == fourcc("il00"), "unsupported inverted lists");
    }
    std::fclose(f);
    return ix;
}

}  // namespace faiss
