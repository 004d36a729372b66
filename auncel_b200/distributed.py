"""Multi-GPU wrappers: one process per GPU, torch.distributed for the plumbing.

* ReplicaGroup -- IndexReplicas semantics (/root/reference/Auncel/IndexReplicas.cpp:79-118):
  every rank holds the whole index, the query batch is cut into contiguous chunks of
  ceil(n / world), results land in disjoint row ranges.  No data-path collective.
* ShardGroup -- IndexShards semantics (IndexShards.cpp:261-311) with a shared quantizer and
  every inverted list split over the ranks (copy_subset_to type 1/2, IndexIVF.cpp:1055-1118,
  the way gpu/GpuAutoTune.cpp:201-220 clones an index over GPUs): every rank searches all
  queries in its shard, one all_gather of the (n x k) distance / label tables over
  NCCL/NVLink, then merge_tables (IndexShards.cpp:44-105) on the device.

The index object only needs `search_device(x_t, k, D_t, I_t)` / `search(x, k)`, so the CPU
tests drive these classes with a recording mock (tests/test_threaded_index.cpp style) over gloo.
"""
import numpy as np


def replica_slice(n, world, rank):
    """IndexReplicas.cpp:95-112: queriesPerIndex = ceil(n / count); base = i * queriesPerIndex."""
    per = (n + world - 1) // world
    base = rank * per
    return base, max(0, min(per, n - base))


def shard_mask(ids, world, rank, subset_type=1, ntotal=None):
    """Which vectors rank `rank` owns. type 1: id % world == rank (IndexIVF.cpp:1086-1095)."""
    ids = np.asarray(ids)
    if subset_type == 1:
        return ids % world == rank
    raise ValueError("subset types other than 1 are list-local; use IndexIVFFlat.copy_subset_to")


class ReplicaGroup:
    def __init__(self, index, group=None):
        import torch.distributed as dist
        self.index, self.group, self.dist = index, group, dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def search(self, x, k):
        """x: the full (n, d) host batch, identical on every rank. Returns this rank's rows
        (base, D, I); nothing is exchanged."""
        base, m = replica_slice(len(x), self.world, self.rank)
        if m == 0:
            return base, np.zeros((0, k), np.float32), np.zeros((0, k), np.int64)
        D, I = self.index.search(x[base:base + m], k)
        return base, D, I

    def search_gathered(self, x, k):
        """Same, then all ranks assemble the full (n, k) tables (what IndexReplicas::search
        leaves in the caller's buffers)."""
        import torch
        n = len(x)
        base, D, I = self.search(x, k)
        per = (n + self.world - 1) // self.world
        if self.world == 1:
            return D, I
        dev = torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend(self.group) == "nccl" \
            else torch.device("cpu")
        Dp = torch.full((per, k), np.nan, dtype=torch.float32, device=dev)
        Ip = torch.full((per, k), -1, dtype=torch.int64, device=dev)
        Dp[:len(D)] = torch.from_numpy(D).to(dev)
        Ip[:len(I)] = torch.from_numpy(I).to(dev)
        Dall = [torch.empty_like(Dp) for _ in range(self.world)]
        Iall = [torch.empty_like(Ip) for _ in range(self.world)]
        self.dist.all_gather(Dall, Dp, group=self.group)
        self.dist.all_gather(Iall, Ip, group=self.group)
        return torch.cat(Dall)[:n].cpu().numpy(), torch.cat(Iall)[:n].cpu().numpy()


class BoundedShardGroup:
    """Error-bounded search over database shards with the reference's distributed semantics
    (Auncel/dist/worker.cpp:153-231,243-267, dist/reduce.cpp:98-119): every worker is an
    independent Auncel index over its slice -- its own calibration against its own ground
    truth, its own termination -- and the per-worker top-k tables are merged at the end.  One
    exchange step: an all_gather of (n x k) distances / labels, then merge_tables."""

    def __init__(self, error_sys, metric, group=None, merge_fn=None):
        self.es, self.shards = error_sys, ShardGroup(_EsAdapter(error_sys), metric, group, merge_fn=merge_fn)

    def search(self, start, n):
        import torch
        D, I = self.es.search(start, n)
        dev = torch.device("cuda", self.es.index.device) if self.shards.dist.is_initialized() and \
            self.shards.dist.get_backend(self.shards.group) == "nccl" else torch.device("cpu")
        x_t = torch.zeros(n, 1, device=dev)
        self.shards.index.tables = (torch.from_numpy(D).to(dev), torch.from_numpy(I).to(dev))
        Dm, Im = self.shards.search_device(x_t, D.shape[1])
        return Dm.cpu().numpy(), Im.cpu().numpy()


class _EsAdapter:
    """lets ShardGroup gather tables that Error_sys.search already produced"""

    def __init__(self, es):
        self.es, self.tables = es, None

    def search_device(self, x_t, k, D_t, I_t):
        D_t.copy_(self.tables[0])
        I_t.copy_(self.tables[1])


class NcclShardGroup:
    """IndexShards over GPUs inside the library (csrc/shards.cu): local search, ONE ncclAllGather of the
    packed (distances | labels) table on the index stream, merge_tables behind it.  torch.distributed
    only carries the 128-byte NCCL unique id from rank 0 to the other ranks at construction."""

    def __init__(self, index, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from ._lib import lib
        from .index import _ck
        self.index = index
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                buf = (C.c_uint8 * 128)()
                _ck(lib().auncel_nccl_unique_id(buf))
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            dev = torch.device("cuda", index.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            uid = uid.to(dev)
            dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uid = uid.cpu()
        raw = (C.c_uint8 * 128)(*uid.tolist())
        h = C.c_void_p()
        _ck(lib().auncel_shard_group_new(C.byref(h), index.h, self.rank, self.world, raw))
        self.h = h

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            from ._lib import lib
            lib().auncel_shard_group_free(h)

    def search_device(self, x_t, k, D_t=None, I_t=None):
        """x_t: (n, d) CUDA tensor, identical on every rank.  Every rank gets the merged (n, k) tables."""
        import torch

        from ._lib import lib
        from .index import _ck
        n = x_t.shape[0]
        if D_t is None:
            D_t = torch.empty(n, k, device=x_t.device, dtype=torch.float32)
            I_t = torch.empty(n, k, device=x_t.device, dtype=torch.int64)
        self.index._after(x_t)
        _ck(lib().auncel_shard_group_search_device(self.h, n, x_t.data_ptr(), k, self.index.nprobe,
                                                   self.index.max_codes, D_t.data_ptr(), I_t.data_ptr()))
        return D_t, I_t

    def search(self, x, k):
        import numpy as np

        from ._lib import _f, _l, lib
        from .index import _ck, _f32, _p
        x = _f32(x)
        D = np.empty((len(x), k), np.float32)
        I = np.empty((len(x), k), np.int64)
        _ck(lib().auncel_shard_group_search(self.h, len(x), _p(x, _f), k, self.index.nprobe, self.index.max_codes,
                                            _p(D, _f), _p(I, _l)))
        return D, I

    def set_bounded(self, on=True):
        """Error-bounded search with SINGLE-INDEX semantics over the shards (csrc/shard_rounds.cu): while on,
        Error_sys.search / sys_train / IndexIVFFlat.search_bounded on the local shard index are collective
        calls -- every rank passes the same queries -- and answer as one index holding all vectors would
        (IndexIVF.cpp:515-660): the ranks exchange each round's candidates before the stage replay."""
        from ._lib import lib
        from .index import _ck
        _ck(lib().auncel_shard_group_set_bounded(self.h, 1 if on else 0))

    def exchange_stats(self):
        import ctypes as C

        from ._lib import lib
        out = (C.c_double * 4)()
        lib().auncel_shard_group_get_exchange_stats(self.h, out)
        return dict(zip(["exchanges", "entries_sent", "entries_all", "bytes_received"], [float(v) for v in out]))

    def stats(self):
        import ctypes as C

        from ._lib import lib
        out = (C.c_double * 8)()
        lib().auncel_shard_group_get_stats(self.h, out)
        return dict(zip(["local_ms", "allgather_ms", "merge_ms", "allgather_bytes", "world", "rank", "nccl_version"],
                        [float(v) for v in out]))


class ShardGroup:
    """The same semantics with torch.distributed doing the gather (any backend): what the CPU tests drive
    with a recording mock over gloo, and what BoundedShardGroup builds on."""

    def __init__(self, index, metric, group=None, translations=None, merge_fn=None, device=None):
        import torch.distributed as dist
        self.index, self.metric, self.group, self.dist = index, metric, group, dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.translations = translations  # successive_ids shifts, IndexShards.cpp:289-297
        self.merge_fn = merge_fn
        self.device = device

    def _merge_device(self, allD, allI, n, k):
        import torch

        from ._lib import lib
        from .index import _ck
        D = torch.empty(n, k, device=allD.device, dtype=torch.float32)
        I = torch.empty(n, k, device=allD.device, dtype=torch.int64)
        tr = None
        if self.translations is not None:
            tr = torch.as_tensor(self.translations, dtype=torch.int64, device=allD.device)
        stream = torch.cuda.current_stream(allD.device).cuda_stream
        _ck(lib().auncel_merge_tables_device(allD.device.index, self.metric, n, k, self.world, allD.data_ptr(),
                                             allI.data_ptr(), None if tr is None else tr.data_ptr(),
                                             D.data_ptr(), I.data_ptr(), stream))
        return D, I

    def search_device(self, x_t, k):
        """x_t: (n, d) tensor, identical on every rank (CUDA for NCCL, CPU for gloo tests).
        Every rank returns the merged (n, k) tables."""
        import torch
        n = x_t.shape[0]
        D = torch.empty(n, k, device=x_t.device, dtype=torch.float32)
        I = torch.empty(n, k, device=x_t.device, dtype=torch.int64)
        self.index.search_device(x_t, k, D, I)
        if x_t.is_cuda:
            torch.cuda.current_stream(x_t.device).synchronize()  # the library ran on its own stream
        if self.world == 1:
            allD, allI = D[None], I[None]
        else:
            allD = torch.empty(self.world, n, k, device=x_t.device, dtype=torch.float32)
            allI = torch.empty(self.world, n, k, device=x_t.device, dtype=torch.int64)
            self.dist.all_gather_into_tensor(allD.view(self.world * n, k), D, group=self.group)
            self.dist.all_gather_into_tensor(allI.view(self.world * n, k), I, group=self.group)
        if self.merge_fn is not None:
            return self.merge_fn(self.metric, allD, allI, self.translations)
        return self._merge_device(allD, allI, n, k)
