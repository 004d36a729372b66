"""CPU, world_size 2 over gloo: the N>1 host logic (query slicing of replicas, shard
all-gather + merge plumbing) with a recording mock index -- the way the reference tests its
multi-device wrappers without devices (tests/test_threaded_index.cpp:21-59,150-253)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from auncel_b200 import distributed as AD
from oracle import oracle as O


class MockIndex:
    """records calls; returns a deterministic sorted table that depends on (rank, query)."""

    def __init__(self, rank, k_valid=None):
        self.rank, self.calls, self.k_valid = rank, [], k_valid

    def _table(self, x, k):
        n = len(x)
        base = np.asarray(x)[:, 0].astype(np.float64)
        D = np.sort((base[:, None] * 0.01 + np.arange(k)[None, :] * (1.0 + 0.37 * self.rank) + 0.1 * self.rank)
                    .astype(np.float32), axis=1)
        I = (np.arange(n)[:, None] * 1000 + np.arange(k)[None, :] * 10 + self.rank).astype(np.int64)
        if self.k_valid is not None:
            I[:, self.k_valid:] = -1
        return D, I

    def search(self, x, k):
        self.calls.append(("search", len(x), k, float(np.asarray(x)[0, 0]) if len(x) else None))
        return self._table(x, k)

    def search_device(self, x_t, k, D_t, I_t):
        self.calls.append(("search_device", x_t.shape[0], k))
        D, I = self._table(x_t.numpy(), k)
        D_t.copy_(torch.from_numpy(D))
        I_t.copy_(torch.from_numpy(I))


def _oracle_merge(metric, allD, allI, tr):
    D, I = O.merge_tables(metric, allD.numpy(), allI.numpy(), tr)
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, k = 11, 4, 5
        x = np.arange(n * d, dtype=np.float32).reshape(n, d)
        # replicas: contiguous ceil(n/world) slices, disjoint rows, no collective needed
        rg = AD.ReplicaGroup(MockIndex(rank))
        base, D, I = rg.search(x, k)
        assert (base, len(D)) == AD.replica_slice(n, world, rank)
        assert rg.index.calls == [("search", len(D), k, float(x[base, 0]))]
        Dg, Ig = rg.search_gathered(x, k)
        exp = np.concatenate([MockIndex(r)._table(x[slice(*(lambda b, m: (b, b + m))(*AD.replica_slice(n, world, r)))], k)[0]
                              for r in range(world)])
        assert np.array_equal(Dg, exp)
        # shards: all queries on every rank, all_gather, merge_tables
        for metric in (O.L2,):
            sg = AD.ShardGroup(MockIndex(rank, k_valid=3 if rank == 1 else None), metric, merge_fn=_oracle_merge,
                               translations=[0, 500])
            Dm, Im = sg.search_device(torch.from_numpy(x), k)
            allD = np.stack([MockIndex(r, 3 if r == 1 else None)._table(x, k)[0] for r in range(world)])
            allI = np.stack([MockIndex(r, 3 if r == 1 else None)._table(x, k)[1] for r in range(world)])
            De, Ie = O.merge_tables(metric, allD, allI, [0, 500])
            assert np.array_equal(Dm.numpy(), De) and np.array_equal(Im.numpy(), Ie)
            assert sg.index.calls == [("search_device", n, k)]
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_replicas_and_shards_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}


def test_replica_slice_matches_reference_rule():
    # IndexReplicas.cpp:95-112
    for n in (0, 1, 7, 8, 9, 100):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                b, m = AD.replica_slice(n, world, r)
                rows += list(range(b, b + m))
            assert rows == list(range(n))


def test_shard_mask_type1():
    ids = np.arange(20)
    got = [AD.shard_mask(ids, 3, r).nonzero()[0].tolist() for r in range(3)]
    assert sorted(sum(got, [])) == list(range(20)) and got[1] == [1, 4, 7, 10, 13, 16, 19]
