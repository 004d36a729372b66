// IndexShards over GPUs, one process per GPU: the multi-GPU form of the path north_star names.
//
// Reference semantics: IndexShards::search (/root/reference/Auncel/IndexShards.cpp:261-311) -- every
// shard answers all queries, then merge_tables (:44-105) -- over sub-indexes that share the coarse
// quantizer and split every inverted list (copy_subset_to, IndexIVF.cpp:1055-1118; the way
// gpu/GpuAutoTune.cpp:201-220 distributes an index).  Here a shard is the AuncelIndex of this
// process; the exchange step is ONE ncclAllGather of a packed (distances | labels) table enqueued on
// the index stream right behind the local search, and merge_tables_kernel right behind that -- no
// host round trip between the three.  NCCL is resolved at run time (dlopen of the libnccl.so.2
// already mapped into the process, e.g. torch's, else the system one): the library itself links
// nothing but the CUDA runtime, and single-GPU users never load NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "../../include/auncel_b200.h"
#include "engine.h"

namespace auncel {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("AUNCEL_NCCL_LIB");
        void* h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.handle = h;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
    });
    AUNCEL_CHECK(api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce,
                 "NCCL (libnccl.so.2) could not be loaded; set AUNCEL_NCCL_LIB");
    return api;
}

#define NCCL_CHECK(x)                                                                                              \
    do {                                                                                                           \
        ncclResult_t r_ = (x);                                                                                     \
        if (r_ != ncclSuccess)                                                                                     \
            AUNCEL_THROW(-4, std::string("NCCL error: ") + (nccl().GetErrorString ? nccl().GetErrorString(r_) : "?") + \
                                 " in " #x);                                                                       \
    } while (0)

}  // namespace

struct ShardGroup {
    IvfIndex* ix = nullptr;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    DevBuf<unsigned char> pack_local, pack_all;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    // last call
    double local_ms = 0, allgather_ms = 0, merge_ms = 0;
    size_t allgather_bytes = 0;  // bytes this rank receives per call (world x packed table)
    ShardExchange xchg;          // error-bounded rounds with single-index semantics (shard_rounds.cu)

    // Route the local index's error-bounded and calibration searches through the round-wise candidate
    // exchange: auncel_index_search_bounded* / auncel_index_calibrate on the local handle then are
    // COLLECTIVE calls (every rank, same queries, same arguments) and return the single-index answer.
    void set_bounded(bool on) {
        if (!on || world == 1) {
            if (ix->shard_x == &xchg) ix->shard_x = nullptr;
            return;
        }
        xchg.rank = rank;
        xchg.world = world;
        ncclComm_t c = comm;
        xchg.all_gather = [c](const void* send, void* recv, size_t bytes, cudaStream_t s) {
            NCCL_CHECK(nccl().AllGather(send, recv, bytes, ncclInt8, c, s));
        };
        xchg.all_reduce_max = [c](long long* buf, size_t count, cudaStream_t s) {
            NCCL_CHECK(nccl().AllReduce(buf, buf, count, ncclInt64, ncclMax, c, s));
        };
        ix->shard_x = &xchg;
    }

    ~ShardGroup() {
        if (ix && ix->shard_x == &xchg) ix->shard_x = nullptr;
        if (comm) nccl().CommDestroy(comm);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (e2) cudaEventDestroy(e2);
    }

    // IndexShards::search: local search -> all-gather of the packed tables -> merge_tables
    void search(long n, const float* x_dev, int k, int nprobe, long max_codes, float* D_dev, long long* I_dev) {
        CUDA_CHECK(cudaSetDevice(ix->device));
        cudaStream_t s = ix->stream;
        local_ms = allgather_ms = merge_ms = 0;
        allgather_bytes = 0;
        if (n == 0) return;
        const size_t dbytes = ((size_t)n * k * sizeof(float) + 7) / 8 * 8;  // labels stay 8-byte aligned
        const size_t B = dbytes + (size_t)n * k * sizeof(long long);
        unsigned char* mine = pack_local.ensure(B);
        unsigned char* all = pack_all.ensure(B * world);
        QueryBatch qb;
        qb.n = n;
        qb.x = x_dev;
        qb.k = k;
        qb.nprobe = (int)std::min<long>(nprobe, ix->nlist);
        qb.max_codes = max_codes;
        qb.mode = 0;
        qb.D = reinterpret_cast<float*>(mine);
        qb.I = reinterpret_cast<long long*>(mine + dbytes);
        ix->search(qb);
        local_ms = ix->stats.search_ms;
        CUDA_CHECK(cudaEventRecord(e0, s));
        if (world > 1) {
            NCCL_CHECK(nccl().AllGather(mine, all, B, ncclInt8, comm, s));
        } else {
            CUDA_CHECK(cudaMemcpyAsync(all, mine, B, cudaMemcpyDeviceToDevice, s));
        }
        CUDA_CHECK(cudaEventRecord(e1, s));
        launch_merge_tables_strided(ix->metric, n, k, world, reinterpret_cast<const float*>(all),
                                    reinterpret_cast<const long long*>(all + dbytes), (long)(B / sizeof(float)),
                                    (long)(B / sizeof(long long)), nullptr, D_dev, I_dev, s);
        CUDA_CHECK(cudaEventRecord(e2, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        float a = 0.f, m = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&a, e0, e1));
        CUDA_CHECK(cudaEventElapsedTime(&m, e1, e2));
        allgather_ms = a;
        merge_ms = m;
        allgather_bytes = B * world;
    }
};

}  // namespace auncel

using namespace auncel;

struct AuncelShardGroup_H {
    ShardGroup g;
    DevBuf<float> x, D;
    DevBuf<long long> I;
};

static thread_local std::string g_shard_err;
extern "C" const char* auncel_get_last_error(void);
void auncel_set_last_error(const std::string& m);  // c_api.cu

#define SH_TRY try {
#define SH_CATCH                                 \
    }                                            \
    catch (const auncel::Error& e) {             \
        auncel_set_last_error(e.what());         \
        return e.code;                           \
    }                                            \
    catch (const std::exception& e) {            \
        auncel_set_last_error(e.what());         \
        return -4;                               \
    }                                            \
    return 0;

IvfIndex* auncel_index_engine(AuncelIndex* idx);  // c_api.cu

extern "C" {

int auncel_nccl_unique_id(void* out128) {
    SH_TRY
    AUNCEL_CHECK(out128 != nullptr, "null output");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_CHECK(nccl().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    SH_CATCH
}

int auncel_shard_group_new(AuncelShardGroup** out, AuncelIndex* local, int rank, int world, const void* nccl_id128) {
    SH_TRY
    AUNCEL_CHECK(out && local, "null argument");
    AUNCEL_CHECK(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank / world size (at most 64 shards)");
    AuncelShardGroup_H* h = new AuncelShardGroup_H;
    h->g.ix = auncel_index_engine(local);
    h->g.rank = rank;
    h->g.world = world;
    CUDA_CHECK(cudaSetDevice(h->g.ix->device));
    CUDA_CHECK(cudaEventCreate(&h->g.e0));
    CUDA_CHECK(cudaEventCreate(&h->g.e1));
    CUDA_CHECK(cudaEventCreate(&h->g.e2));
    if (world > 1) {
        AUNCEL_CHECK(nccl_id128 != nullptr, "a world of more than one shard needs the NCCL unique id of rank 0");
        ncclUniqueId id;
        memcpy(&id, nccl_id128, sizeof(id));
        ncclResult_t r = nccl().CommInitRank(&h->g.comm, world, id, rank);
        if (r != ncclSuccess) {
            delete h;
            AUNCEL_THROW(-4, std::string("ncclCommInitRank failed: ") + (nccl().GetErrorString ? nccl().GetErrorString(r) : "?"));
        }
    }
    *out = h;
    SH_CATCH
}

void auncel_shard_group_free(AuncelShardGroup* g) { delete g; }

int auncel_shard_group_search_device(AuncelShardGroup* g, int64_t n, const float* x_dev, int64_t k, int64_t nprobe,
                                     int64_t max_codes, float* distances_dev, int64_t* labels_dev) {
    SH_TRY
    AUNCEL_CHECK(k >= 1 && k <= MAX_K, "k must be in [1, 128]");
    g->g.search((long)n, x_dev, (int)k, (int)nprobe, (long)max_codes, distances_dev, (long long*)labels_dev);
    SH_CATCH
}

int auncel_shard_group_search(AuncelShardGroup* g, int64_t n, const float* x, int64_t k, int64_t nprobe,
                              int64_t max_codes, float* distances, int64_t* labels) {
    SH_TRY
    AUNCEL_CHECK(k >= 1 && k <= MAX_K, "k must be in [1, 128]");
    IvfIndex& ix = *g->g.ix;
    CUDA_CHECK(cudaSetDevice(ix.device));
    if (n == 0) return 0;
    g->x.ensure((size_t)n * ix.d);
    g->D.ensure((size_t)n * k);
    g->I.ensure((size_t)n * k);
    CUDA_CHECK(cudaMemcpyAsync(g->x.p, x, (size_t)n * ix.d * sizeof(float), cudaMemcpyHostToDevice, ix.stream));
    g->g.search((long)n, g->x.p, (int)k, (int)nprobe, (long)max_codes, g->D.p, g->I.p);
    CUDA_CHECK(cudaMemcpyAsync(distances, g->D.p, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToHost, ix.stream));
    CUDA_CHECK(cudaMemcpyAsync(labels, g->I.p, (size_t)n * k * sizeof(long long), cudaMemcpyDeviceToHost, ix.stream));
    CUDA_CHECK(cudaStreamSynchronize(ix.stream));
    SH_CATCH
}

int auncel_shard_group_set_bounded(AuncelShardGroup* g, int on) {
    SH_TRY
    AUNCEL_CHECK(g != nullptr, "null group");
    g->g.set_bounded(on != 0);
    SH_CATCH
}

int auncel_shard_group_get_exchange_stats(const AuncelShardGroup* g, double* out4) {
    out4[0] = (double)g->g.xchg.exchanges;
    out4[1] = (double)g->g.xchg.entries_sent;
    out4[2] = (double)g->g.xchg.entries_recv;
    out4[3] = (double)g->g.xchg.bytes_recv;
    return 0;
}

int auncel_shard_group_get_stats(const AuncelShardGroup* g, double* out8) {
    out8[0] = g->g.local_ms;
    out8[1] = g->g.allgather_ms;
    out8[2] = g->g.merge_ms;
    out8[3] = (double)g->g.allgather_bytes;
    out8[4] = (double)g->g.world;
    out8[5] = (double)g->g.rank;
    int v = 0;
    if (g->g.world > 1 && nccl().GetVersion) nccl().GetVersion(&v);
    out8[6] = (double)v;
    out8[7] = 0;
    return 0;
}

}  // extern "C"
