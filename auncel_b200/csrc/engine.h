// Internal C++ engine behind the C ABI (include/auncel_b200.h).
// One IvfIndex = one device-resident IndexIVFFlat (centroids + inverted lists) together
// with Auncel's error model and the scratch arenas the query path needs.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <functional>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "errmodel.h"

namespace auncel {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define AUNCEL_THROW(code, msg) throw ::auncel::Error(code, std::string(msg) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")")
#define AUNCEL_CHECK(cond, msg) do { if (!(cond)) AUNCEL_THROW(-2, std::string("Error: '" #cond "' failed: ") + msg); } while (0)
#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) AUNCEL_THROW(-4, std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #x); } while (0)

// grow-only device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    T* ensure(size_t n) {
        if (n > cap) {
            release();
            size_t want = n + n / 8 + 16;
            CUDA_CHECK(cudaMalloc(&p, want * sizeof(T)));
            cap = want;
        }
        return p;
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    T* ensure(size_t n) {
        if (n > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            CUDA_CHECK(cudaMallocHost(&p, (n + 16) * sizeof(T)));
            cap = n + 16;
        }
        return p;
    }
};

constexpr int MAX_K = 128;          // widest result heap the fused selection supports
constexpr int SCAN_QT = 32;         // queries per scan tile
constexpr int SCAN_VT = 128;        // list vectors per pipeline block
constexpr int SCAN_DK = 32;         // floats per k-chunk
constexpr int SCAN_THREADS = 256;

struct SearchStats {  // IndexIVFStats, IndexIVF.h:361-374
    uint64_t nq = 0, nlist = 0, ndis = 0, nheap_updates = 0;
    double quantization_ms = 0, search_ms = 0;
    uint64_t rounds = 0, scan_tiles = 0, scan_pairs = 0, launches = 0, scan_launches = 0;
    uint64_t tc_rounds = 0, tc_candidates = 0, tc_fallbacks = 0;
    double tc_ms = 0;        // device time of the tensor-core filter kernels
    uint64_t tc_ndis = 0;    // distance evaluations (sum of |list| over pairs) they covered
    uint64_t simt_ndis = 0;  // ... covered by exact-scan rounds
    uint64_t tc_uniq = 0, tc_staged = 0, simt_uniq = 0, simt_staged = 0;  // vectors: distinct lists touched / staged per tile
    double simt_ms = 0;      // device time of the scan phase of those rounds
    double coarse_ms = 0;
    double scan_ms = 0;      // device time of the scan kernels of the last search
    uint64_t err_bits = 0;   // ERR_* bits raised by the last search
    // option "tc_audit": slots of tensor-core rounds compared with an exact rescan of the round
    uint64_t tc_audit_bad = 0, tc_audit_slots = 0, tc_audit_cands = 0;
};

// per-query arguments of a search call; all pointers are DEVICE pointers
struct QueryBatch {
    long n = 0;
    const float* x = nullptr;  // n x d
    int k = 0;                 // result width (max_topk in tune mode)
    int nprobe = 0;
    long max_codes = 0;
    int mode = 0;              // 0 fixed, 1 tune (error bounded), 2 training (calibration)
    int query_topk = 0;
    const float* require_acc = nullptr;  // n (time_tune: the latency budget in ms, IndexIVF.cpp:547)
    int time_tune = 0;                   // error_pro::time_tune: latency-budget cut after every list (:545-549)
    const float* gt_kth = nullptr;       // n, GT distance at rank query_topk-1 (profile)
    unsigned long long* my_nprobe = nullptr;  // n, in/out
    float* t_recalls = nullptr;          // n, in/out
    int profile = 0, overhead_profile = 0;
    float* D = nullptr;       // n x k
    long long* I = nullptr;   // n x k
    float* snapshots = nullptr;  // training: n x n_traces x k sorted distances
    float* dtb_out = nullptr;    // optional n x max_num (training needs it on the host)
};

struct RoundParams;

// Candidate exchange between the shards of one index (shard_rounds.cu).  The collectives are injected by
// the owner of the communicator (shards.cu: NCCL), the engine itself links no communication library.
struct ShardExchange {
    int rank = 0, world = 1;
    // every rank contributes `bytes` bytes from `send`; `recv` receives world x bytes, in rank order
    std::function<void(const void* send, void* recv, size_t bytes, cudaStream_t s)> all_gather;
    // element-wise maximum over ranks, in place
    std::function<void(long long* buf, size_t count, cudaStream_t s)> all_reduce_max;
    DevBuf<unsigned long long> ctr;
    DevBuf<int> cnt2, ovf_ord, fill, inv;
    DevBuf<unsigned char> pool, send, recv, ovf_pool;
    std::vector<unsigned long long> h_counts;
    // since the last reset (one search): candidates this rank sent / all ranks sent, bytes received, rounds
    uint64_t entries_sent = 0, entries_recv = 0, bytes_recv = 0, exchanges = 0;
};

struct IvfIndex {
    int d = 0, dpad = 0;
    long nlist = 0;
    int metric = METRIC_L2;
    int device = 0;
    long ntotal = 0;
    bool trained = false;

    // device-resident index
    DevBuf<float> centroids;   // nlist x dpad
    DevBuf<float> interdis;    // nlist*(nlist-1)/2, packed triangle (Auncel's interdis_cem)
    bool have_interdis = false;
    DevBuf<float> codes;       // ntotal x dpad, lists concatenated in list order
    DevBuf<long long> ids;     // ntotal
    DevBuf<long long> list_off;  // nlist + 1 (in vectors)
    std::vector<long long> h_list_off;
    alignas(64) unsigned char codes_tmap[128];  // CUtensorMap over `codes` (rebuilt by add)
    alignas(64) unsigned char codes_tmap64[128];  // the same arena with 64-row boxes (tcfilter2.cu)

    // error model (device copies + host mirror)
    std::vector<float> h_arcos;
    std::vector<long> h_trace_off;
    std::vector<float> h_phi, h_U, h_sigma;
    DevBuf<float> d_arcos, d_phi, d_U, d_sigma;
    DevBuf<long> d_trace_off;
    float multipler = 1.f, std_m = 1.f;
    int n_traces = 0;

    // scratch (grow-only)
    DevBuf<float> q_x;         // staged queries (n x dpad)
    DevBuf<float> c_raw;       // unranked coarse distances (one chunk)
    DevBuf<float> c_dis;       // n x nlist coarse distances, ranked
    DevBuf<int> c_keys;        // n x nlist ranked centroid ids
    DevBuf<int> c_tie0;        // n: first rank with an equal-distance neighbour (INT_MAX: none / replayed)
    DevBuf<int> c_sorted;      // n: ranks of the row that are in place (partial ranking), nlist = all
    bool partial_rank = false; // last coarse_rank ranked only the best centroids
    int partial_rank_mode = 1; // 0 never, 1 large batches, 2 whenever nlist allows (tests)
    DevBuf<int> fix_list;      // scratch for the tie replay
    DevBuf<int> heap_entry;    // heap_entry_table(heap_entry_k), device copy
    int heap_entry_k = -1;
    const int* entry_table(int k);
    bool exact_ties = true;
    DevBuf<float> dtb;         // n x max_num
    DevBuf<unsigned char> state;  // per-query running state, see search.cu
    DevBuf<unsigned char> pool;
    DevBuf<int> slot_cnt;
    DevBuf<int> active, active2;
    DevBuf<int> list_cnt, list_pair_off, list_tile_off, list_cursor;
    DevBuf<unsigned long long> pairs;
    DevBuf<float> q_sorted;
    DevBuf<float> vnorm, qnorm;              // squared norms of arena rows / of the batch's queries
    DevBuf<float> list_nmax;                 // largest squared norm in every inverted list
    DevBuf<unsigned long long> tc_cand;      // survivors of the tensor-core filter
    DevBuf<int> pair_flag;                   // per slot: overflowed in a tensor-core round
    DevBuf<unsigned char> redo_pool;         // exact redo of those pairs: compact pool, 4 sub-slots per pair
    DevBuf<int> redo_cnt, redo_ord;
    int tc_mode = 1;                         // 0 off, 1 automatic, 2 whenever every active heap is full
    int tc_audit = 0;                        // tests: redo every tensor-core round exactly and compare the slots
    int tc_stream_min = 96;                  // d > 256: stream the query tile when lists are probed by at least this many queries
    int tc_kernel = 0;                       // 0 / 1 tcfilter.cu, 2 tcfilter2.cu (TMEM-resident queries; d <= 256), 3 tcfilter3.cu (CTA pairs)
    DevBuf<unsigned char> audit_pool;
    DevBuf<int> audit_cnt;
    DevBuf<unsigned long long> audit_ctr;
    DevBuf<int> ctl;           // small control block (counters)
    PinnedBuf<int> h_ctl;
    DevBuf<float> io_f;        // host-API staging
    DevBuf<long long> io_l;
    DevBuf<unsigned long long> io_u;
    DevBuf<unsigned long long> assign_best;

    // range search (range.cu): results of the last call stay here until they are fetched
    DevBuf<unsigned long long> range_off, range_lims;
    DevBuf<float> range_D;
    DevBuf<long long> range_I;
    long long range_total = 0;
    // clock model of the latency-budget mode (IndexIVF::time(), IndexIVF.cpp:329-333): cost of one
    // probe iteration / of one scanned code.  Default: a B200 streaming a list at HBM speed.
    long long time_us_per_list = 2, time_ns_per_code = 0;

    size_t pool_budget_bytes = (size_t)16 << 30;  // candidate pools are sparse; 180 GB of HBM make a wide window cheap
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    cudaEvent_t ev_in = nullptr;  // auncel_index_wait_stream: orders this stream behind a caller stream
    std::vector<cudaEvent_t> scan_ev;  // pairs of events around every scan launch
    std::vector<cudaEvent_t> tc_ev;    // pairs of events around every tensor-core filter launch
    DevBuf<unsigned long long> round_work;
    PinnedBuf<unsigned long long> h_round_work;
    SearchStats stats;
    bool debug_rounds = false;
    std::vector<std::array<int, 6>> round_log;
    // per round of the last search: r0, w, active queries, 1 = tensor-core filter round, distance evaluations,
    // vectors of the distinct lists touched, vectors staged, scan-phase ms, tc_filter_kernel ms, (reserved)
    std::vector<std::array<double, 10>> round_stats;
    int num_sms = 148;

    IvfIndex(int d, long nlist, int metric, int device);
    ~IvfIndex();

    // centroids are HOST pointers here
    void set_centroids(const float* c, bool compute_interdis);
    void get_centroids(float* out) const;
    void get_interdis(float* out) const;
    void set_interdis(const float* in);

    // x: device pointer n x d (row stride d); list_no/ids: host (nullable)
    void add_device(long n, const float* x_dev, const long long* ids_host, const long long* list_no_host);
    void assign_device(long n, const float* x_dev, long long* list_no_host);
    void reset();
    void gather_rows(const long long* rows_host, long m, float* out_host);

    void set_error_model(int arcos_size, int n_traces, const long* trace_off, const float* phi,
                         const float* U, const float* sigma, float multipler, float std_m);

    // coarse: ranks all nlist centroids for n staged queries (x_dev n x d) into c_dis/c_keys
    void coarse_rank(long n, const float* x_dev, bool allow_partial = false);
    void search(const QueryBatch& qb);
    // error-bounded / calibration search over the shards of one index with single-index semantics: when
    // set, every round of search() exchanges its candidates with the other ranks (shard_rounds.cu)
    ShardExchange* shard_x = nullptr;
    void exchange_candidates(RoundParams& rp, size_t nredo);
    // IndexIVF::range_search: lims_host gets n + 1 offsets; distances / labels stay in range_D / range_I
    void range_search(long n, const float* x_dev, float radius, int nprobe, long long* lims_host);
    int max_num() const { return (int)(nlist / 8 + 20); }
    int expected_traces() const { int t = 0; for (long p = 1; p <= nlist / 8; p <<= 1) t++; return t; }
    ErrModelView model_view() const;
};

// ---- kernels' host launchers (one per .cu) ----
void launch_pad_rows(const float* src, long n, int d, float* dst, int dpad, cudaStream_t s);
void launch_coarse_distances(int metric, const float* xq, long nq, const float* cent, long nlist,
                             int dpad, float* out_dis /*nq x nlist or null*/,
                             unsigned long long* out_best /*nq or null*/, cudaStream_t s);
void launch_rank_rows(int metric, const float* dis, long nq, long nlist, float* out_dis, int* out_keys,
                      int* tie0, cudaStream_t s, const int* qlist = nullptr, int* sorted_upto = nullptr,
                      const int* qbound = nullptr, int need_upto = 0);
void launch_rank_rows_partial(int metric, const float* dis, long nq, long nlist, float* out_dis, int* out_keys,
                              int* tie0, int* sorted_upto, cudaStream_t s);
int rank_rows_partial_width();
// replay the reference's size-k heap for the queries of `list` (all n if null) that have equal
// coarse distances below rank `bound`; rewrites their rows of out_dis/out_keys in heap order
void heap_entry_table(int k, std::vector<int>& entry);
void launch_fix_ties(int metric, const float* raw, long nlist, int k, const int* entry, const int* list, int n,
                     int* tie0, int bound, const int* qbound, int* fix_list, int* nfix, float* out_dis, int* out_keys, cudaStream_t s,
                     const int* decided = nullptr, int r0 = 0, int* err = nullptr, int* sorted_upto = nullptr);
void launch_interdis(int metric, const float* cent, long nlist, int dpad, float* out, cudaStream_t s);
void launch_merge_tables(int metric, long n, long k, long nshard, const float* all_D,
                         const long long* all_I, const long long* translations, float* D,
                         long long* I, cudaStream_t s);

void launch_merge_tables_strided(int metric, long n, long k, long nshard, const float* all_D, const long long* all_I,
                                 long stride_D, long stride_I, const long long* translations, float* D, long long* I,
                                 cudaStream_t s);
// Index::train: k-means on the device (kmeans.cu), then set_centroids
void train_kmeans(IvfIndex& ix, long nx, const float* x_host, int niter, bool tune);

}  // namespace auncel
