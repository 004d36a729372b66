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
        int flags = (t->profile ? 1 : 0) | (t->overhead_profile ? 2 : 0);
        auncel_check(auncel_index_search_bounded(h, n, x, k, (int64_t)t->query_topk, t->require_acc + offset,
                                                 kth.empty() ? nullptr : kth.data(), np.data(),
                                                 t->t_recalls ? t->t_recalls + offset : nullptr, flags, distances,
                                                 ll.data()));
        for (idx_t i = 0; i < n; i++) t->my_nprobe[offset + i] = (size_t)np[i];
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

}  // namespace faiss
