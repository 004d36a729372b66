"""Host-side mirror of the reference's Index API for this path, above the C ABI.

Names, argument meaning and error behaviour follow the reference
(/root/reference/Auncel): IndexIVFFlat (IndexIVFFlat.h, IndexIVF.h:97-308) with `nprobe`,
`max_codes`, train/add/add_core/search/reset, and Error_sys (profile.h:29-91) with
set_gt / sys_train / set_topk / set_queries / search.  Errors surface as FaissException
like the reference's FAISS_THROW_* (FaissAssert.h:56-93).  numpy arrays = host buffers;
`*_device` variants take torch CUDA tensors (their storage is only addressed, never
copied) -- torch is plumbing for device memory here, nothing else.
"""
import ctypes as C

import numpy as np

from ._lib import _f, _l, _u, lib

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1


class FaissException(RuntimeError):
    """faiss::FaissException (FaissException.h)"""


def _ck(rc):
    if rc != 0:
        raise FaissException(lib().auncel_get_last_error().decode())


def _p(a, t):
    return C.cast(None, t) if a is None else a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class IndexIVFFlat:
    def __init__(self, d, nlist, metric=METRIC_L2, device=0):
        self.d, self.nlist, self.metric_type, self.device = int(d), int(nlist), int(metric), int(device)
        self.nprobe = 1          # IndexIVF.h: nprobe
        self.max_codes = 0       # IndexIVF.h: max_codes
        self.tune = False        # Index.h:42-77 set_tune_mode/off
        h = C.c_void_p()
        _ck(lib().auncel_index_new(C.byref(h), self.d, self.nlist, self.metric_type, self.device))
        self.h = h

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            lib().auncel_index_free(h)

    # ---- Index.h fields
    @property
    def ntotal(self):
        return lib().auncel_index_ntotal(self.h)

    @property
    def is_trained(self):
        return bool(lib().auncel_index_is_trained(self.h))

    def set_tune_mode(self):
        self.tune = True

    def set_tune_off(self):
        self.tune = False

    # ---- train / centroids
    def train(self, x, niter=25):
        x = _f32(x)
        _ck(lib().auncel_index_train(self.h, len(x), _p(x, _f), niter, int(self.tune)))

    def set_centroids(self, c, compute_interdis=True):
        c = _f32(c)
        if c.shape != (self.nlist, self.d):
            raise FaissException("centroids must be nlist x d")
        _ck(lib().auncel_index_set_centroids(self.h, _p(c, _f), int(compute_interdis)))

    def centroids(self):
        out = np.empty((self.nlist, self.d), np.float32)
        _ck(lib().auncel_index_get_centroids(self.h, _p(out, _f)))
        return out

    def interdis_cem(self):
        out = np.empty(self.nlist * (self.nlist - 1) // 2, np.float32)
        _ck(lib().auncel_index_get_interdis(self.h, _p(out, _f)))
        return out

    # ---- add
    def add(self, x):
        self.add_core(x, None, None)

    def add_with_ids(self, x, xids):
        self.add_core(x, xids, None)

    def add_core(self, x, xids=None, precomputed_idx=None):
        x = _f32(x)
        xids = None if xids is None else _i64(xids)
        pre = None if precomputed_idx is None else _i64(precomputed_idx)
        _ck(lib().auncel_index_add(self.h, len(x), _p(x, _f), _p(xids, _l), _p(pre, _l)))

    def _after(self, t):
        """Stream contract of the `_device` entry points (include/auncel_b200.h): the index runs on
        its own stream, so it is ordered behind whatever torch still has pending on the current
        stream of the tensor's device (the copy or kernel that produces the inputs)."""
        import torch
        _ck(lib().auncel_index_wait_stream(self.h, torch.cuda.current_stream(t.device).cuda_stream))

    def add_device(self, x_t, xids=None, precomputed_idx=None):
        assert x_t.is_cuda and x_t.is_contiguous() and x_t.dtype.is_floating_point and x_t.element_size() == 4
        xids = None if xids is None else _i64(xids)
        pre = None if precomputed_idx is None else _i64(precomputed_idx)
        self._after(x_t)
        _ck(lib().auncel_index_add_device(self.h, x_t.shape[0], x_t.data_ptr(), _p(xids, _l), _p(pre, _l)))

    def assign(self, x):
        x = _f32(x)
        out = np.empty(len(x), np.int64)
        _ck(lib().auncel_index_assign(self.h, len(x), _p(x, _f), _p(out, _l)))
        return out

    def reset(self):
        _ck(lib().auncel_index_reset(self.h))

    def list_sizes(self):
        out = np.empty(self.nlist, np.int64)
        lib().auncel_index_list_sizes(self.h, _p(out, _l))
        return out

    # ---- search
    def coarse_search(self, x, nprobe):
        x = _f32(x)
        dis = np.empty((len(x), nprobe), np.float32)
        keys = np.empty((len(x), nprobe), np.int64)
        _ck(lib().auncel_index_coarse_search(self.h, len(x), _p(x, _f), nprobe, _p(dis, _f), _p(keys, _l)))
        return dis, keys

    def search(self, x, k):
        x = _f32(x)
        D = np.empty((len(x), k), np.float32)
        I = np.empty((len(x), k), np.int64)
        _ck(lib().auncel_index_search(self.h, len(x), _p(x, _f), k, self.nprobe, self.max_codes, _p(D, _f),
                                      _p(I, _l)))
        return D, I

    def search_device(self, x_t, k, D_t, I_t):
        self._after(x_t)
        _ck(lib().auncel_index_search_device(self.h, x_t.shape[0], x_t.data_ptr(), k, self.nprobe, self.max_codes,
                                             D_t.data_ptr(), I_t.data_ptr()))

    def range_search(self, x, radius):
        """IndexIVF::range_search (IndexIVF.cpp:741-860) -> (lims, D, I): result of query i =
        entries lims[i]:lims[i+1], unsorted, in scan order (RangeSearchResult)."""
        x = _f32(x)
        lims = np.empty(len(x) + 1, np.int64)
        _ck(lib().auncel_index_range_search(self.h, len(x), _p(x, _f), radius, self.nprobe, _p(lims, _l)))
        D = np.empty(int(lims[-1]), np.float32)
        I = np.empty(int(lims[-1]), np.int64)
        _ck(lib().auncel_index_range_search_results(self.h, _p(D, _f), _p(I, _l)))
        return lims, D, I

    def set_time_model(self, us_per_list, ns_per_code):
        _ck(lib().auncel_index_set_time_model(self.h, int(us_per_list), int(ns_per_code)))

    def search_timed(self, x, k, budget_ms):
        """Error_sys::time_search's search (profile.cpp:229-244): latency budget per query, ms."""
        x, budget_ms = _f32(x), _f32(budget_ms)
        assert len(budget_ms) == len(x)
        D = np.empty((len(x), k), np.float32)
        I = np.empty((len(x), k), np.int64)
        _ck(lib().auncel_index_search_timed(self.h, len(x), _p(x, _f), k, _p(budget_ms, _f), _p(D, _f), _p(I, _l)))
        return D, I

    # ---- error model
    def set_error_model(self, traces, multipler=1.0, std_m=1.0):
        off = np.zeros(len(traces) + 1, np.int64)
        np.cumsum([len(t[0]) for t in traces], out=off[1:])
        phi = _f32(np.concatenate([t[0] for t in traces]))
        U = _f32(np.concatenate([t[1] for t in traces]))
        sg = _f32(np.concatenate([t[2] for t in traces]))
        _ck(lib().auncel_index_set_error_model(self.h, len(traces), _p(off, _l), _p(phi, _f), _p(U, _f), _p(sg, _f),
                                               multipler, std_m))

    def set_params(self, multipler, std_m):
        lib().auncel_index_set_params(self.h, multipler, std_m)

    def traces(self):
        out = []
        for t in range(lib().auncel_index_n_traces(self.h)):
            n = lib().auncel_index_trace_size(self.h, t)
            phi, U, sg = (np.empty(n, np.float32) for _ in range(3))
            _ck(lib().auncel_index_get_trace(self.h, t, _p(phi, _f), _p(U, _f), _p(sg, _f)))
            out.append((phi, U, sg))
        return out

    def calibrate(self, xq, max_topk, gt_D):
        xq, gt_D = _f32(xq), _f32(gt_D)
        D = np.empty((len(xq), max_topk), np.float32)
        I = np.empty((len(xq), max_topk), np.int64)
        _ck(lib().auncel_index_calibrate(self.h, len(xq), _p(xq, _f), max_topk, _p(gt_D, _f), _p(D, _f), _p(I, _l)))
        return D, I

    def search_bounded(self, x, max_topk, query_topk, require_acc, my_nprobe=None, gt_kth=None, t_recalls=None,
                       profile=False, overhead_profile=False):
        x, require_acc = _f32(x), _f32(require_acc)
        n = len(x)
        if my_nprobe is None:
            my_nprobe = np.zeros(n, np.uint64)
        assert my_nprobe.dtype == np.uint64 and len(my_nprobe) == n and len(require_acc) == n
        gt_kth = None if gt_kth is None else _f32(gt_kth)
        D = np.empty((n, max_topk), np.float32)
        I = np.empty((n, max_topk), np.int64)
        flags = int(profile) | (int(overhead_profile) << 1) | (int(getattr(self, "time_tune", False)) << 2)
        _ck(lib().auncel_index_search_bounded(self.h, n, _p(x, _f), max_topk, query_topk, _p(require_acc, _f),
                                              _p(gt_kth, _f), _p(my_nprobe, _u), _p(t_recalls, _f), flags,
                                              _p(D, _f), _p(I, _l)))
        return D, I, my_nprobe

    def search_bounded_device(self, x_t, max_topk, query_topk, acc_t, np_t, D_t, I_t, gt_t=None, trec_t=None,
                              flags=0):
        self._after(x_t)
        _ck(lib().auncel_index_search_bounded_device(
            self.h, x_t.shape[0], x_t.data_ptr(), max_topk, query_topk, acc_t.data_ptr(),
            None if gt_t is None else gt_t.data_ptr(), np_t.data_ptr(),
            None if trec_t is None else trec_t.data_ptr(), flags, D_t.data_ptr(), I_t.data_ptr()))

    def stats(self):
        out = (C.c_double * 32)()
        lib().auncel_index_get_stats(self.h, out)
        return dict(zip(["nq", "nlist", "ndis", "search_ms", "rounds", "scan_tiles", "scan_pairs", "err_bits",
                         "scan_ms", "launches", "scan_launches", "coarse_ms", "tc_rounds", "tc_candidates",
                         "tc_fallbacks", "tc_ms", "tc_ndis", "simt_ms", "simt_ndis", "tc_uniq", "tc_staged",
                         "simt_uniq", "simt_staged", "tc_audit_bad", "tc_audit_slots", "tc_audit_cands"],
                        [float(v) for v in out]))

    def round_stats(self):
        """per round of the last search: dicts with r0, w, active, tc, ndis, uniq, staged, scan_ms, tc_ms"""
        out = (C.c_double * (64 * 10))()
        n = C.c_int(0)
        lib().auncel_index_get_round_stats(self.h, 64, out, C.byref(n))
        keys = ["r0", "w", "active", "tc", "ndis", "uniq", "staged", "scan_ms", "tc_ms"]
        return [dict(zip(keys, [float(out[r * 10 + j]) for j in range(9)])) for r in range(min(n.value, 64))]

    def set_option(self, name, value):
        _ck(lib().auncel_index_set_option(self.h, name.encode(), int(value)))

    def set_pool_budget(self, nbytes):
        lib().auncel_index_set_pool_budget(self.h, nbytes)

    def copy_subset_to(self, other, subset_type, a1, a2):
        _ck(lib().auncel_index_copy_subset_to(self.h, other.h, subset_type, a1, a2))


class Error_sys:
    """Error_sys (profile.h:29-91, profile.cpp)."""

    def __init__(self, index, nq, topk):
        if nq % 10 != 0:  # profile.cpp:31-32
            raise FaissException("Error: 'nq%10 == 0' failed: Train num must be evenly divided by ten")
        self.index, self.train_num, self.max_topk = index, nq, topk
        self.is_trained = False
        self.train_D = self.train_I = None
        self.query_topk = None
        self.profile = False
        self.overhead_profile = False

    def set_gt(self, gt_D, gt_I):
        if gt_D is None or gt_I is None:  # profile.cpp:46-50
            raise FaissException("the ground truth must not be null ptr when setting up")
        self.train_D = _f32(gt_D).reshape(self.train_num, self.max_topk).copy()
        self.train_I = _i64(gt_I).reshape(self.train_num, self.max_topk).copy()

    def sys_train(self, nq, xq):
        if nq > self.train_num:  # profile.cpp:89-90
            raise FaissException("Error sys training does not have the same nb of queries compared with creation")
        if self.train_I is None:
            raise FaissException("ground truth not initialized")
        self.index.calibrate(_f32(xq)[:nq], self.max_topk, self.train_D[:nq])
        self.is_trained = True

    def set_topk(self, new_topk):
        self.query_topk = new_topk

    def set_queries(self, n, q, acc, allo_size):
        self.num, self.queries, self.require_acc = n, _f32(q), _f32(acc)
        self.my_nprobe = np.zeros(allo_size, np.uint64)   # profile.cpp:178-181
        self.t_recalls = np.zeros(allo_size, np.float32)

    def setparam(self, multipler, std_m):
        """error_pro::setparam (IVF_pro.cpp:240-256) with the two values given directly."""
        self.index.set_params(multipler, std_m)
        self.profile = False

    def time_search(self, start, search_size=-1):
        """Error_sys::time_search (profile.cpp:229-244): require_acc[id] is read as a latency budget
        in ms; the tune block stays off.  Like the reference (:242) this leaves error_pro::time_tune
        SET, so a later search() applies the latency cut as well until `index.time_tune = False`."""
        if not self.is_trained:
            raise FaissException("Error sys must be trained before searching")
        if self.num > self.train_num:
            raise FaissException("Error sys search num must be lower than all qeuries num")
        n = self.num if search_size == -1 else search_size
        sl = slice(start, start + n)
        self.index.time_tune = True
        return self.index.search_timed(self.queries[sl], self.max_topk, self.require_acc[sl])

    def search(self, start, search_size=-1):
        if not self.is_trained:  # profile.cpp:212-213
            raise FaissException("Error sys must be trained before searching")
        if self.num > self.train_num:
            raise FaissException("Error sys search num must be lower than all qeuries num")
        n = self.num if search_size == -1 else search_size
        sl = slice(start, start + n)
        np_ = self.my_nprobe[sl].copy()
        tr = self.t_recalls[sl].copy()
        gt_kth = np.ascontiguousarray(self.train_D[sl, self.query_topk - 1]) if self.train_D is not None else None
        D, I, np_ = self.index.search_bounded(self.queries[sl], self.max_topk, self.query_topk,
                                              self.require_acc[sl], np_, gt_kth, tr, self.profile,
                                              self.overhead_profile)
        self.my_nprobe[sl] = np_
        self.t_recalls[sl] = tr
        return D, I


def merge_tables(metric, all_D, all_I, translations=None):
    """merge_tables (IndexShards.cpp:44-105); all_D/all_I: (nshard, n, k) host arrays."""
    all_D, all_I = _f32(all_D), _i64(all_I)
    nshard, n, k = all_D.shape
    tr = None if translations is None else _i64(translations)
    D = np.empty((n, k), np.float32)
    I = np.empty((n, k), np.int64)
    _ck(lib().auncel_merge_tables(metric, n, k, nshard, _p(all_D, _f), _p(all_I, _l), _p(tr, _l), _p(D, _f),
                                  _p(I, _l)))
    return D, I
