// A driver in the shape of the reference's eval/bound.cpp (Auncel/eval/bound.cpp:216-430),
// written against include/auncel/faiss_api.h: build "IVF<nlist>,Flat", calibrate the error
// model, answer queries one by one under an error bound, check the bound the way the
// reference does (inter_sec, bound.cpp:117-128,404-414); then the IndexShards / IndexReplicas
// invariants of tests/test_merge.cpp and tests/test_threaded_index.cpp.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <sys/time.h>

#include "auncel/faiss_api.h"

static double elapsed() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

static size_t inter_sec(const float* gt, size_t topk, const float* D) {  // bound.cpp:117-128, type 0
    size_t res = 0;
    float t_val = gt[topk - 1];
    for (size_t i = 0; i < topk; i++)
        if (D[i] <= t_val + 1e-6) res++;
    return res;
}

int main(int argc, char** argv) {
    const int d = 32;
    const size_t nlist = 1024, nb = 200000, ts = 500, ses = 500, k = 40, input_k = 10;
    const float error_bound = 0.1f;
    int fails = 0;
    std::mt19937 rng(123);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> centers(256 * d);
    for (auto& v : centers) v = nd(rng);
    auto gen = [&](size_t n) {
        std::vector<float> x(n * d);
        for (size_t i = 0; i < n; i++) {
            size_t c = rng() % 256;
            for (int j = 0; j < d; j++) x[i * d + j] = centers[c * d + j] + 0.6f * nd(rng);
        }
        return x;
    };
    std::vector<float> xb = gen(nb), xq = gen(ts + ses);
    double t0 = elapsed();
    try {
        std::unique_ptr<faiss::Index> index_p(faiss::index_factory(d, "IVF1024,Flat"));  // bound.cpp:220
        faiss::IndexIVFFlat& index = *dynamic_cast<faiss::IndexIVFFlat*>(index_p.get());
        faiss::IndexFlat& quantizer = *dynamic_cast<faiss::IndexFlat*>(index.quantizer);
        index.niter = 8;
        index.set_tune_mode();  // bound.cpp:261-263
        index.train(nb, xb.data());
        index.set_tune_off();
        index.add(nb, xb.data());
        printf("[%.3f s] trained + added, ntotal=%ld\n", elapsed() - t0, index.ntotal);

        // ground truth = exhaustive search
        size_t nq = ts + ses;
        std::vector<float> gt_v(nq * k);
        std::vector<faiss::Index::idx_t> gt(nq * k);
        index.nprobe = nlist;
        index.search(nq, xq.data(), k, gt_v.data(), gt.data());
        printf("[%.3f s] ground truth done\n", elapsed() - t0);

        faiss::Error_sys err_sys(&index, nq, k);  // bound.cpp:356-361
        err_sys.set_gt(gt_v.data(), gt.data());
        err_sys.sys_train(ts, xq.data());
        err_sys.set_topk(input_k);
        std::vector<float> acc(ts + ses, 1.f - error_bound), D(ses * k);
        std::vector<int64_t> I(ses * k);
        err_sys.set_queries(ses, xq.data(), acc.data(), ts + ses);
        index.t->multipler = 6.0f;  // what setparam(figureid) reads from hyperparameter.txt
        index.t->std_m = 2.0f;
        double t1 = elapsed();
        for (size_t i = ts; i < ts + ses; i++)  // bound.cpp:390-396: one query per call
            err_sys.search(D.data() + k * (i - ts), I.data() + k * (i - ts), i, 1);
        double lat = (elapsed() - t1) / ses;
        float minf = 1.f;
        for (size_t i = ts; i < ts + ses; i++)
            minf = std::min(minf, inter_sec(&gt_v[i * k], input_k, D.data() + (i - ts) * k) / float(input_k));
        double mean_np = 0;
        for (size_t i = ts; i < ts + ses; i++) mean_np += index.t->my_nprobe[i];
        printf("latency mode: %.1f us/query, mean my_nprobe %.1f, Error Bound : %f -> %s\n", lat * 1e6, mean_np / ses, minf,
               minf >= 1 - error_bound ? "Error bound is guaranteed" : "NOT guaranteed");
        if (minf < 1 - error_bound) fails++;

        // batch mode gives the same answers (effect_error.cpp:294)
        std::vector<float> D2(ses * k);
        std::vector<int64_t> I2(ses * k);
        err_sys.set_queries(ses, xq.data(), acc.data(), ts + ses);
        err_sys.search(D2.data(), I2.data(), ts);
        bool same = D2 == D;
        printf("batch == one-by-one: %s\n", same ? "yes" : "NO");
        fails += !same;

        // error behaviour (FAISS_THROW_* -> FaissException)
        try {
            faiss::Error_sys bad(&index, 15, k);
            fails++;
        } catch (const faiss::FaissException& e) {
            printf("expected exception: %s\n", e.what());
        }

        // shards: 3 sub-indexes sharing the centroids == the single index (tests/test_merge.cpp:94-152)
        faiss::set_index_parameters(&index, "nprobe=8,max_codes=0");  // AutoTune.cpp:455-563
        std::vector<float> Dref(ses * 10);
        std::vector<faiss::Index::idx_t> Iref(ses * 10);
        index.search(ses, xq.data() + ts * d, 10, Dref.data(), Iref.data());
        std::vector<faiss::IndexFlatL2*> qs;
        std::vector<faiss::IndexIVFFlat*> subs;
        faiss::IndexShards shards(d, true, false);
        for (int s = 0; s < 3; s++) {
            qs.push_back(new faiss::IndexFlatL2(d));
            qs.back()->add(nlist, quantizer.xb.data());
            subs.push_back(new faiss::IndexIVFFlat(qs.back(), d, nlist));
            subs.back()->nprobe = 8;
            shards.add_shard(subs.back());
        }
        shards.add(nb, xb.data());
        std::vector<float> Ds(ses * 10);
        std::vector<faiss::Index::idx_t> Is(ses * 10);
        shards.search(ses, xq.data() + ts * d, 10, Ds.data(), Is.data());
        size_t ndiff = 0;
        for (size_t i = 0; i < Is.size(); i++) ndiff += Is[i] != Iref[i];
        printf("shards vs single index: ndiff=%zu, distances %s\n", ndiff, Ds == Dref ? "equal" : "DIFFER");
        fails += !(Ds == Dref);
        // replicas: query split, no merge (IndexReplicas.cpp:79-118)
        faiss::IndexReplicas reps(d, true);
        reps.addReplica(&index);
        reps.addReplica(&index);
        std::vector<float> Dr(ses * 10);
        std::vector<faiss::Index::idx_t> Ir(ses * 10);
        reps.search(ses, xq.data() + ts * d, 10, Dr.data(), Ir.data());
        printf("replicas vs single index: %s\n", (Dr == Dref && Ir == Iref) ? "equal" : "DIFFER");
        fails += !(Dr == Dref && Ir == Iref);
        for (auto p : subs) delete p;
        for (auto p : qs) delete p;
    } catch (const std::exception& e) {
        printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    printf(fails ? "DEMO FAILED (%d)\n" : "DEMO OK\n", fails);
    return fails ? 1 : 0;
}
