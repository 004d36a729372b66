/*
 * auncel_b200 -- C ABI of the B200-native error-bounded IVF-Flat query path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Conventions
 * follow the reference's own C wrapper (/root/reference/Auncel/c_api): opaque handle, int
 * status (0 ok; -1 unknown, -2 invalid argument / failed FAISS_THROW_IF_NOT-style check,
 * -4 CUDA/runtime error: c_api/error_c.h:19-33), thread-local last error string
 * (c_api/error_impl.cpp:17-28).  idx_t is 64-bit (Index.h:67); matrices are row-major and
 * compact; the caller owns every buffer; missing results are label -1 with distance
 * FLT_MAX (L2) / -FLT_MAX (IP) (Heap.h:295-322).
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/Auncel).  "host" entry points take host pointers and do their own
 * host<->device copies; "_device" entry points take device pointers on the index's device
 * and enqueue on the index's own (non-blocking) stream, returning after the results are
 * complete.  STREAM CONTRACT: the index stream is not ordered against any other stream.  If
 * the inputs of a "_device" call are produced by work still pending on another stream (a
 * cudaMemcpyAsync, a kernel of the caller), call auncel_index_wait_stream(idx, that_stream)
 * first (or synchronise that stream); outputs are complete when the call returns.
 */
#ifndef AUNCEL_B200_H
#define AUNCEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AuncelIndex_H AuncelIndex;

#define AUNCEL_METRIC_INNER_PRODUCT 0 /* Index.h:49 */
#define AUNCEL_METRIC_L2 1            /* Index.h:50 */

/* c_api/error_c.h: faiss_get_last_error */
const char* auncel_get_last_error(void);

/* error bits accumulated by a tuned search where the reference would have thrown or read
 * out of bounds: 1 arcos domain (IVF_pro.cpp:180), 2 arcos x==1 (IVF_pro.cpp:182),
 * 4 cosine_theorem precondition (IVF_pro.cpp:42). */

/* IndexIVFFlat(quantizer, d, nlist, metric) with an IndexFlatL2/IP quantizer
 * (IndexIVFFlat.h:26-28, c_api/IndexIVFFlat_c.h faiss_IndexIVFFlat_new_with_metric). */
int auncel_index_new(AuncelIndex** out, int d, int64_t nlist, int metric, int device);
/* c_api/Index_c.h faiss_Index_free */
void auncel_index_free(AuncelIndex* idx);

int auncel_index_d(const AuncelIndex* idx);
int64_t auncel_index_nlist(const AuncelIndex* idx);
int64_t auncel_index_ntotal(const AuncelIndex* idx);   /* Index::ntotal */
int auncel_index_is_trained(const AuncelIndex* idx);   /* Index::is_trained */

/* Orders the index's stream behind everything enqueued so far on `cuda_stream` (a cudaStream_t;
 * NULL = the legacy default stream): event record + cudaStreamWaitEvent, no host wait. */
int auncel_index_wait_stream(AuncelIndex* idx, void* cuda_stream);

/* Import trained centroids (nlist x d, host) into the coarse quantizer -- the state
 * Level1Quantizer::train_q1 leaves behind (IndexIVF.cpp:71-137).  compute_interdis != 0 also
 * fills Auncel's pairwise centroid table interdis_cem (:97-117; what index->set_tune_mode()
 * before train() enables, eval/bound.cpp:261-263). */
int auncel_index_set_centroids(AuncelIndex* idx, const float* centroids, int compute_interdis);
int auncel_index_get_centroids(const AuncelIndex* idx, float* out);
/* packed strict upper triangle, nlist*(nlist-1)/2 floats (IVF_pro.cpp:21-39) */
int auncel_index_get_interdis(const AuncelIndex* idx, float* out);
int auncel_index_set_interdis(AuncelIndex* idx, const float* in);

/* Index::train (IndexIVF.cpp:995-1008): k-means on the device (Clustering.cpp:77-244
 * semantics: niter iterations, seed 1234, at most 256 points per centroid, spherical for IP)
 * followed by set_centroids.  tune != 0 <=> set_tune_mode() was called before train(). */
int auncel_index_train(AuncelIndex* idx, int64_t n, const float* x, int niter, int tune);

/* IndexIVFFlat::add_core (IndexIVFFlat.cpp:41-80).  ids == NULL: ntotal + i.  list_no ==
 * NULL: quantizer->assign (k = 1 coarse search); otherwise precomputed_idx (< 0 skips). */
int auncel_index_add(AuncelIndex* idx, int64_t n, const float* x, const int64_t* ids,
                     const int64_t* list_no);
int auncel_index_add_device(AuncelIndex* idx, int64_t n, const float* x_dev, const int64_t* ids,
                            const int64_t* list_no);
/* Index::assign (Index.cpp:42-47) */
int auncel_index_assign(AuncelIndex* idx, int64_t n, const float* x, int64_t* list_no);
/* Index::reset */
int auncel_index_reset(AuncelIndex* idx);
/* InvertedLists::list_size for every list (InvertedLists.h:43) */
int auncel_index_list_sizes(const AuncelIndex* idx, int64_t* out);

/* ArrayInvertedLists contents (InvertedLists.h:182-202) in list order: codes (sum of list sizes x d
 * floats) and ids; either pointer may be NULL.  What index_io's write_InvertedLists stores
 * (index_io.cpp:280-330). */
int auncel_index_get_lists(const AuncelIndex* idx, float* codes, int64_t* ids);
int auncel_index_get_params(const AuncelIndex* idx, float* multipler, float* std_m);
int auncel_index_has_interdis(const AuncelIndex* idx);

/* quantizer->search(n, x, nprobe, coarse_dis, idx) (IndexFlat.cpp:42-56): the nprobe best
 * centroids per query, best first, with the reference's exact per-pair arithmetic
 * (knn_L2sqr_sse / knn_inner_product_sse, utils.cpp:417-490). */
int auncel_index_coarse_search(AuncelIndex* idx, int64_t n, const float* x, int64_t nprobe,
                               float* coarse_dis, int64_t* keys);

/* IndexIVF::search (IndexIVF.cpp:335-353): fixed nprobe, optional max_codes (0 = off). */
int auncel_index_search(AuncelIndex* idx, int64_t n, const float* x, int64_t k, int64_t nprobe,
                        int64_t max_codes, float* distances, int64_t* labels);
int auncel_index_search_device(AuncelIndex* idx, int64_t n, const float* x_dev, int64_t k,
                               int64_t nprobe, int64_t max_codes, float* distances_dev,
                               int64_t* labels_dev);

/* error_pro state the online check needs (IVF_pro.h:82-114): n_traces ascending (phi, U,
 * sigma) tables (Trace, IVF_pro.h:44-62) concatenated with trace_off[n_traces+1], and the
 * two hyper-parameters of hyperparameter.txt / error_pro::setparam (IVF_pro.cpp:240-256). */
int auncel_index_set_error_model(AuncelIndex* idx, int n_traces, const int64_t* trace_off,
                                 const float* phi, const float* U, const float* sigma,
                                 float multipler, float std_m);
int auncel_index_set_params(AuncelIndex* idx, float multipler, float std_m);
int auncel_index_n_traces(const AuncelIndex* idx);
int64_t auncel_index_trace_size(const AuncelIndex* idx, int t);
int auncel_index_get_trace(const AuncelIndex* idx, int t, float* phi, float* U, float* sigma);

/* Error_sys::sys_train (profile.cpp:88-171) + error_pro::train (IVF_pro.cpp:186-194):
 * calibration search of n queries (training block, IndexIVF.cpp:640-673) against their exact
 * ground-truth distances gt_D (n x max_topk), then Trace::SB.  Installs the resulting error
 * model.  distances/labels (n x max_topk, may be NULL) receive the calibration search result. */
int auncel_index_calibrate(AuncelIndex* idx, int64_t n, const float* x, int64_t max_topk,
                           const float* gt_D, float* distances, int64_t* labels);

/* Error_sys::search (profile.cpp:211-227) -> IndexIVF::search(n, x, k, D, I, offset) with
 * tune = true, nprobe = nlist (IndexIVF.cpp:355-378, tune block :551-638).
 *   max_topk     width of the result heap (GT depth);  query_topk = error_pro::query_topk
 *   require_acc  n targets, 1 - error bound            (error_pro::require_acc[id])
 *   gt_kth       n ground-truth distances at rank query_topk-1, or NULL (only `profile`)
 *   my_nprobe    n, in/out: error_pro::my_nprobe[id] (0 = undecided on input)
 *   t_recalls    n, in/out, or NULL: error_pro::t_recalls[id]
 *   flags        bit 0 = error_pro::profile, bit 1 = error_pro::overhead_profile,
 *                bit 2 = error_pro::time_tune (the flag Error_sys::time_search leaves set, profile.cpp:242:
 *                require_acc is then ALSO read as a latency budget in ms, IndexIVF.cpp:545-549) */
int auncel_index_search_bounded(AuncelIndex* idx, int64_t n, const float* x, int64_t max_topk,
                                int64_t query_topk, const float* require_acc,
                                const float* gt_kth, uint64_t* my_nprobe, float* t_recalls,
                                int flags, float* distances, int64_t* labels);
int auncel_index_search_bounded_device(AuncelIndex* idx, int64_t n, const float* x_dev,
                                       int64_t max_topk, int64_t query_topk,
                                       const float* require_acc_dev, const float* gt_kth_dev,
                                       uint64_t* my_nprobe_dev, float* t_recalls_dev, int flags,
                                       float* distances_dev, int64_t* labels_dev);

/* Error_sys::time_search (profile.cpp:229-244): nprobe = nlist, no tune block; after every probed
 * list the reference breaks when  elapsed_ms >= 0.95 * budget_ms - elapsed_ms / lists_done
 * (IndexIVF.cpp:545-549, IndexIVF::time() = gettimeofday :329-333).  On the device the clock is a
 * MODEL, so the cut is deterministic: every probe iteration costs us_per_list microseconds plus
 * ns_per_code nanoseconds per code of the list; time() is evaluated on the resulting (tv_sec,
 * tv_usec) pair with the reference's double arithmetic.  budget_ms: n values (what the reference
 * keeps in require_acc[id] in this mode).  Default model: 2 us per list, 0 ns per code. */
int auncel_index_set_time_model(AuncelIndex* idx, int64_t us_per_list, int64_t ns_per_code);
int auncel_index_search_timed(AuncelIndex* idx, int64_t n, const float* x, int64_t k,
                              const float* budget_ms, float* distances, int64_t* labels);
int auncel_index_search_timed_device(AuncelIndex* idx, int64_t n, const float* x_dev, int64_t k,
                                     const float* budget_ms_dev, float* distances_dev,
                                     int64_t* labels_dev);

/* IndexIVF::range_search (IndexIVF.cpp:741-860, scan_codes_range IndexIVFFlat.cpp:139-155): all
 * vectors of the nprobe nearest lists with dis < radius (L2) / dis > radius (inner product).
 * Two calls, like RangeSearchResult (AuxIndexStructures.h:31-50): the first searches and fills
 * lims[n + 1] (result of query i = entries lims[i] .. lims[i+1]); the caller allocates lims[n]
 * entries (do_allocation) and the second call copies distances / labels, in the reference's scan
 * order (probe rank, then in-list order).  The results of a search are kept until the next one. */
int auncel_index_range_search(AuncelIndex* idx, int64_t n, const float* x, float radius,
                              int64_t nprobe, int64_t* lims);
int auncel_index_range_search_results(AuncelIndex* idx, float* distances, int64_t* labels);

/* IndexIVFStats (IndexIVF.h:361-374) of the last search on this index + engine counters.
 * out[0]=nq [1]=nlist visited [2]=ndis [3]=search ms (device) [4]=rounds [5]=scan tiles
 * [6]=(query,list) pairs scanned [7]=error bits [8]=scan-kernel ms (CUDA events around every
 * scan launch) [9]=kernel launches [10]=scan-kernel launches [11]=coarse ms [12]=rounds served by
 * the tensor-core filter [13]=its survivors [14]=rounds it handed back to the exact scan
 * [15]=tensor-core filter kernel ms [16]=distance evaluations those launches covered
 * [17]=scan-phase ms of exact-scan rounds [18]=distance evaluations of those rounds
 * [19]/[21]=vectors of the distinct lists touched per launch, summed (tensor-core / exact rounds)
 * [20]/[22]=vectors staged into shared memory (one list pass per query tile)
 * [23..25] option "tc_audit": slots that differ from the exact rescan / slots compared / candidates compared.
 * out: 32 doubles */
int auncel_index_get_stats(const AuncelIndex* idx, double* out32);

/* Per round of the last search (a round = one window of probe ranks for all still-active queries), 10 doubles each:
 * [0] first rank [1] window width [2] active queries [3] 1 = served by the tensor-core filter [4] distance
 * evaluations [5] vectors of the distinct lists touched (compulsory HBM traffic / 4d bytes) [6] vectors staged into
 * shared memory [7] scan-phase ms [8] tc_filter_kernel ms [9] reserved.  *n_rounds receives the number of rounds. */
int auncel_index_get_round_stats(const AuncelIndex* idx, int max_rounds, double* out, int* n_rounds);

/* engine switches (results never change): "tensor_core_filter" 0 off / 1 automatic / 2 whenever
 * every active query holds K results; "exact_ties" 0/1 replay of the reference's heap order for
 * equal centroid distances; "partial_rank" 0 / 1 / 2: rank only the best 1024 centroids of a query up front and
 * complete a row when a round reads past them -- never / for batches >= 2048 (default) / always; "tc_audit" 0/1 (tests) redo every tensor-core round with the exact scan and
 * compare the candidate slots; "tc_kernel" 0 or 1 the default filter kernel (queries resident in shared memory), 2 queries
 * resident in tensor memory (d <= 256), 3 CTA pairs (tcgen05 cta_group::2) -- measured alternatives, docs/TC_FILTER_VARIANTS.md;
 * "tc_stream_min" (default 96): at d > 256 the filter streams the query tile through its stage ring when the lists of a
 * round are probed by at least this many queries on average */
int auncel_index_set_option(AuncelIndex* idx, const char* name, int value);

/* scratch budget for per-round candidate pools, bytes (default 16 GiB) */
int auncel_index_set_pool_budget(AuncelIndex* idx, size_t bytes);

/* merge_tables (IndexShards.cpp:44-105): k-way merge of nshard sorted (n x k) result tables
 * (layout [shard][query][rank]); labels < 0 end a shard's row; translations may be NULL. */
int auncel_merge_tables(int metric, int64_t n, int64_t k, int64_t nshard, const float* all_distances,
                        const int64_t* all_labels, const int64_t* translations, float* distances,
                        int64_t* labels);
int auncel_merge_tables_device(int device, int metric, int64_t n, int64_t k, int64_t nshard,
                               const float* all_distances_dev, const int64_t* all_labels_dev,
                               const int64_t* translations_dev, float* distances_dev,
                               int64_t* labels_dev, void* cuda_stream);

/* ---- IndexShards over GPUs, one process per GPU (IndexShards.cpp:261-311 + merge_tables :44-105) ----
 * Every rank holds one shard: an AuncelIndex with the SHARED centroids and its slice of every inverted
 * list (copy_subset_to below, or add_with_ids of the rank's vectors).  A search runs the local
 * fixed-nprobe search of all n queries, ONE ncclAllGather of the packed (n x k distances | n x k labels)
 * table on the index stream, and merge_tables behind it on the same stream; every rank returns the
 * merged (n x k) tables, equal to the unsharded index (tests/test_merge.cpp:94-152 invariant).
 * Setup: rank 0 calls auncel_nccl_unique_id and hands the 128 bytes to all ranks by any means (MPI,
 * torch.distributed, a file); every rank then calls auncel_shard_group_new (collective: ncclCommInitRank).
 * NCCL is loaded at run time (the libnccl.so.2 already in the process, else the system's; override with
 * AUNCEL_NCCL_LIB); world == 1 needs neither an id nor NCCL. */
typedef struct AuncelShardGroup_H AuncelShardGroup;
int auncel_nccl_unique_id(void* out128);
int auncel_shard_group_new(AuncelShardGroup** out, AuncelIndex* local_shard, int rank, int world,
                           const void* nccl_id128);
void auncel_shard_group_free(AuncelShardGroup* g);
int auncel_shard_group_search(AuncelShardGroup* g, int64_t n, const float* x, int64_t k, int64_t nprobe,
                              int64_t max_codes, float* distances, int64_t* labels);
int auncel_shard_group_search_device(AuncelShardGroup* g, int64_t n, const float* x_dev, int64_t k,
                                     int64_t nprobe, int64_t max_codes, float* distances_dev,
                                     int64_t* labels_dev);
/* last call: out[0] local search ms, [1] all-gather ms, [2] merge ms (CUDA events on the index stream),
 * [3] bytes received by the all-gather, [4] world, [5] rank, [6] NCCL version code */
int auncel_shard_group_get_stats(const AuncelShardGroup* g, double* out8);

/* Error-bounded search over the shards with SINGLE-INDEX semantics.
 * The reference's own sharded mode (IndexShards.cpp:261-311) lets every shard stop on its own
 * partial top-k.  With set_bounded(g, 1) the local shard handle's auncel_index_search_bounded*,
 * auncel_index_calibrate and Error_sys-style calls become COLLECTIVE (all ranks, same queries and
 * arguments, same pool budget) and reproduce IndexIVF.cpp:515-660 as run on ONE index holding all
 * vectors: every rank scans its part of each probed list, the ranks exchange the per-(query, stage)
 * candidates of a round (one ncclAllGather of a compact list), and every rank replays the
 * stage-order merge + phi-U test on the union.  distances / my_nprobe / t_recalls are those of the
 * single index bit for bit; labels differ only inside groups of equal distances.  max_codes and
 * time_search are refused in this mode (they cut on list sizes, which are local).
 * set_bounded(g, 0) restores per-shard behaviour.  get_exchange_stats (last search):
 * out[0] exchanges (= rounds), [1] candidates this rank sent, [2] candidates all ranks sent,
 * [3] bytes this rank received. */
int auncel_shard_group_set_bounded(AuncelShardGroup* g, int on);
int auncel_shard_group_get_exchange_stats(const AuncelShardGroup* g, double* out4);

/* IndexIVF::copy_subset_to (IndexIVF.cpp:1055-1118): append to `other` (same centroids) the
 * entries selected by subset_type 1 (id % a1 == a2) or 2 (proportional in-list slice a1..a2
 * of ntotal).  This is how an index is split across GPUs (gpu/GpuAutoTune.cpp:201-220). */
int auncel_index_copy_subset_to(const AuncelIndex* idx, AuncelIndex* other, int subset_type,
                                int64_t a1, int64_t a2);

/* Host-only helper behind the exact coarse tie order (csrc/coarse.cu heap_order_kernel): the
 * structure of the reference's size-k result heap while it fills (knn_L2sqr_sse / heap_pop,
 * utils.cpp:417-490, Heap.h:88-117).  entry_out[j], j < k = node where insertion j's sift-down meets
 * its first data-dependent comparison; entry_out[k] = J, the number of leading insertions whose
 * sift-downs run in subtrees disjoint from the root-to-slot-k path (done level-parallel on the GPU).
 * entry_out must hold k + 1 ints.  Exposed so the CPU test-suite can check the table against a literal
 * replay. */
int auncel_heap_entry_table(int64_t k, int32_t* entry_out);

#ifdef __cplusplus
}
#endif
#endif
