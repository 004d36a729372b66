"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

* ``liboracle.so``            the plain-C restatement (oracle/auncel_oracle.c)
* ``_ref/libauncel_ref.so``   the unmodified reference compiled from /root/reference
                              (oracle/ref_driver.cpp); optional on a box without it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (auncel_b200/) never does.
"""
import ctypes as C
import glob
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IP, L2 = 0, 1

_f = C.POINTER(C.c_float)
_l = C.POINTER(C.c_long)
_ul = C.POINTER(C.c_ulong)
_i = C.POINTER(C.c_int)


def _p(a, t):
    if a is None:
        return C.cast(None, t)
    return a.ctypes.data_as(t)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def build(force=False):
    """make -C oracle (restatement always; _ref only when /root/reference exists)."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so"))
    ref_src = os.path.exists("/root/reference/Auncel/IndexIVF.cpp")
    if ref_src and not os.path.exists(os.path.join(HERE, "_ref", "libauncel_ref.so")):
        need = True
    if need:
        subprocess.check_call(["make", "-C", HERE, "-j8"], stdout=subprocess.DEVNULL)


_orc = None
_ref = None


def orc():
    global _orc
    if _orc is None:
        build()
        lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
        lib.orc_fvec_L2sqr.restype = C.c_float
        lib.orc_fvec_L2sqr.argtypes = [_f, _f, C.c_long]
        lib.orc_fvec_inner_product.restype = C.c_float
        lib.orc_fvec_inner_product.argtypes = [_f, _f, C.c_long]
        lib.orc_coarse.argtypes = [C.c_int, C.c_long, _f, C.c_long, _f, C.c_int, C.c_long, _f, _l]
        lib.orc_interdis.argtypes = [C.c_int, C.c_long, C.c_int, _f, _f]
        lib.orc_construct_arcos.argtypes = [C.c_int, _f]
        lib.orc_arcos.restype = C.c_float
        lib.orc_arcos.argtypes = [_f, C.c_int, C.c_float, _i]
        lib.orc_cosine_theorem.restype = C.c_float
        lib.orc_cosine_theorem.argtypes = [C.c_float, C.c_float, C.c_float, _i]
        lib.orc_set_online.argtypes = [C.c_int, C.c_long, _f, _l, _f, _f, C.c_int, _f, _f, _i]
        lib.orc_sum_angle.restype = C.c_float
        lib.orc_sum_angle.argtypes = [C.c_float, _f, C.c_long, C.c_long, _f, C.c_int, _i]
        lib.orc_trace_search.restype = C.c_float
        lib.orc_trace_search.argtypes = [_f, _f, _f, C.c_long, C.c_float, C.c_float]
        lib.orc_trace_SB.restype = C.c_long
        lib.orc_trace_SB.argtypes = [_f, C.c_long, C.c_long, _f, _f, _f]
        lib.orc_kscaling.restype = C.c_float
        lib.orc_kscaling.argtypes = [C.c_float, C.c_long, _f, C.c_long]
        lib.orc_cur_num.restype = C.c_size_t
        lib.orc_cur_num.argtypes = [_f, C.c_int, C.c_int, _l, _f, _f, _f, C.c_float, _f, _f,
                                    C.c_long, C.c_long, _i]
        lib.orc_search_preassigned.restype = C.c_int
        lib.orc_search_preassigned.argtypes = [
            C.c_int, C.c_int, C.c_long, _f, _l, _l, C.c_long, _f, C.c_long, C.c_long, C.c_long,
            _l, _f, C.c_int, C.c_long, _f, _f, C.c_int, C.c_int, _l, _f, _f, _f, C.c_float,
            C.c_float, C.c_long, _f, _f, C.c_int, C.c_int, _ul, _f, _f, C.c_long, _f, _l, _l,
            C.c_long, _f]
        lib.orc_merge_tables.argtypes = [C.c_int, C.c_long, C.c_long, C.c_long, _f, _l, _f, _l, _l]
        lib.orc_set_time_tune.argtypes = [C.c_int, C.c_long, C.c_long]
        lib.orc_range_search.argtypes = [C.c_int, C.c_int, _f, _l, _l, C.c_long, _f, C.c_float, C.c_long, _l, _l,
                                         _f, _l, _l]
        _orc = lib
    return _orc


def find_openblas():
    pats = [os.path.join(p, "opencv_python_headless.libs", "libopenblas*.so*")
            for p in __import__("site").getsitepackages()]
    for pat in pats:
        for f in glob.glob(pat):
            return f
    return None


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libauncel_ref.so"))


def ref():
    global _ref
    if _ref is None:
        build()
        lib = C.CDLL(os.path.join(HERE, "_ref", "libauncel_ref.so"))
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_create.restype = C.c_void_p
        lib.ref_create.argtypes = [C.c_int, C.c_long, C.c_int]
        lib.ref_free.argtypes = [C.c_void_p]
        lib.ref_train.argtypes = [C.c_void_p, C.c_long, _f, C.c_int, C.c_int]
        lib.ref_set_centroids.argtypes = [C.c_void_p, _f]
        lib.ref_get_centroids.argtypes = [C.c_void_p, _f]
        lib.ref_interdis_size.restype = C.c_long
        lib.ref_interdis_size.argtypes = [C.c_void_p]
        lib.ref_get_interdis.argtypes = [C.c_void_p, _f]
        lib.ref_add.argtypes = [C.c_void_p, C.c_long, _f, _l, _l]
        lib.ref_ntotal.restype = C.c_long
        lib.ref_ntotal.argtypes = [C.c_void_p]
        lib.ref_list_sizes.argtypes = [C.c_void_p, _l]
        lib.ref_list_ids.argtypes = [C.c_void_p, C.c_long, _l]
        lib.ref_assign.argtypes = [C.c_void_p, C.c_long, _f, _l]
        lib.ref_coarse.argtypes = [C.c_void_p, C.c_long, _f, C.c_long, _f, _l]
        lib.ref_search_fixed.argtypes = [C.c_void_p, C.c_long, _f, C.c_long, C.c_long, C.c_long, _f, _l]
        lib.ref_es_create.argtypes = [C.c_void_p, C.c_long, C.c_long, _f, _l]
        lib.ref_es_sys_train.argtypes = [C.c_void_p, C.c_long, _f, C.c_char_p]
        lib.ref_n_traces.restype = C.c_long
        lib.ref_n_traces.argtypes = [C.c_void_p]
        lib.ref_trace_size.restype = C.c_long
        lib.ref_trace_size.argtypes = [C.c_void_p, C.c_long]
        lib.ref_trace_nstd.restype = C.c_long
        lib.ref_trace_nstd.argtypes = [C.c_void_p, C.c_long]
        lib.ref_get_trace.argtypes = [C.c_void_p, C.c_long, _f, _f, _f]
        lib.ref_set_trace.argtypes = [C.c_void_p, C.c_long, C.c_long, _f, _f, _f]
        lib.ref_get_arcos.argtypes = [C.c_void_p, _f]
        lib.ref_es_set_queries.argtypes = [C.c_void_p, C.c_long, C.c_long, _f, C.c_long, _f,
                                           C.c_long, C.c_float, C.c_float, C.c_int, C.c_int]
        lib.ref_es_search.argtypes = [C.c_void_p, _f, _l, C.c_long, C.c_long]
        lib.ref_es_search_threads.argtypes = [C.c_void_p, _f, _l, C.c_long, C.c_long, C.c_int]
        lib.ref_es_search_threads_chunk.argtypes = [C.c_void_p, _f, _l, C.c_long, C.c_long, C.c_int, C.c_long]
        lib.ref_search_fixed_threads.argtypes = [C.c_void_p, C.c_long, _f, C.c_long, C.c_long, _f, _l, C.c_int]
        lib.ref_get_my_nprobe.argtypes = [C.c_void_p, C.c_long, C.c_long, _ul]
        lib.ref_clear_my_nprobe.argtypes = [C.c_void_p]
        lib.ref_get_t_recalls.argtypes = [C.c_void_p, C.c_long, C.c_long, _f]
        lib.ref_stats_get.argtypes = [C.POINTER(C.c_double)]
        lib.ref_shards_search.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_long, _f, C.c_long,
                                          C.c_long, _f, _l, C.c_int]
        lib.ref_replicas_search.argtypes = lib.ref_shards_search.argtypes
        lib.ref_copy_subset_to.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_long]
        lib.ref_fvec_L2sqr.restype = C.c_float
        lib.ref_fvec_L2sqr.argtypes = [_f, _f, C.c_long]
        lib.ref_fvec_inner_product.restype = C.c_float
        lib.ref_fvec_inner_product.argtypes = [_f, _f, C.c_long]
        lib.ref_cosine_theorem.restype = C.c_float
        lib.ref_cosine_theorem.argtypes = [C.c_float, C.c_float, C.c_float]
        lib.ref_kscaling.restype = C.c_float
        lib.ref_kscaling.argtypes = [C.c_float, C.c_long, _f, C.c_long]
        lib.ref_set_virtual_clock.argtypes = [C.c_int]
        lib.ref_es_time_search.argtypes = [C.c_void_p, _f, _l, C.c_long, C.c_long, C.c_int]
        lib.ref_range_search.argtypes = [C.c_void_p, C.c_long, _f, C.c_float, C.c_long, _l]
        lib.ref_range_results.argtypes = [C.c_long, _f, _l]
        lib.ref_shim_set_blas.argtypes = [C.c_char_p]
        blas = find_openblas()
        if blas:
            # the bundled OpenBLAS needs its sibling libquadmath/libgfortran preloaded
            for pat in ("libquadmath*", "libgfortran*"):
                for dep in glob.glob(os.path.join(os.path.dirname(blas), pat)):
                    try:
                        C.CDLL(dep, mode=C.RTLD_GLOBAL)
                    except OSError:
                        pass
            lib.ref_shim_set_blas(blas.encode())
        _ref = lib
    return _ref


class RefError(RuntimeError):
    pass


def _ck(rc):
    if rc != 0:
        raise RefError(ref().ref_last_error().decode())


class RefIndex:
    """The real reference IndexIVFFlat (+ Error_sys) behind oracle/ref_driver.cpp."""

    def __init__(self, d, nlist, metric=L2):
        self.lib = ref()
        self.d, self.nlist, self.metric = d, nlist, metric
        self.h = self.lib.ref_create(d, nlist, metric)
        self.max_topk = None

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    @staticmethod
    def set_blas_threshold(v):
        ref().ref_set_blas_threshold(int(v))

    @staticmethod
    def has_blas():
        return bool(ref().ref_shim_has_blas())

    def train(self, x, niter=0, verbose=False):
        x = f32(x)
        _ck(self.lib.ref_train(self.h, len(x), _p(x, _f), niter, int(verbose)))

    def set_centroids(self, c):
        c = f32(c)
        assert c.shape == (self.nlist, self.d)
        _ck(self.lib.ref_set_centroids(self.h, _p(c, _f)))

    def centroids(self):
        out = np.empty((self.nlist, self.d), np.float32)
        self.lib.ref_get_centroids(self.h, _p(out, _f))
        return out

    def interdis(self):
        out = np.empty(self.lib.ref_interdis_size(self.h), np.float32)
        self.lib.ref_get_interdis(self.h, _p(out, _f))
        return out

    def add(self, x, ids=None, list_no=None):
        x = f32(x)
        ids = None if ids is None else i64(ids)
        list_no = None if list_no is None else i64(list_no)
        _ck(self.lib.ref_add(self.h, len(x), _p(x, _f), _p(ids, _l), _p(list_no, _l)))

    def list_sizes(self):
        out = np.empty(self.nlist, np.int64)
        self.lib.ref_list_sizes(self.h, _p(out, _l))
        return out

    def list_ids(self, l):
        n = int(self.list_sizes()[l])
        out = np.empty(n, np.int64)
        self.lib.ref_list_ids(self.h, l, _p(out, _l))
        return out

    def assign(self, x):
        x = f32(x)
        out = np.empty(len(x), np.int64)
        _ck(self.lib.ref_assign(self.h, len(x), _p(x, _f), _p(out, _l)))
        return out

    def coarse(self, x, nprobe):
        x = f32(x)
        dis = np.empty((len(x), nprobe), np.float32)
        keys = np.empty((len(x), nprobe), np.int64)
        _ck(self.lib.ref_coarse(self.h, len(x), _p(x, _f), nprobe, _p(dis, _f), _p(keys, _l)))
        return dis, keys

    def search_fixed(self, x, k, nprobe, max_codes=0, threads=0):
        x = f32(x)
        D = np.empty((len(x), k), np.float32)
        I = np.empty((len(x), k), np.int64)
        if threads > 1:
            _ck(self.lib.ref_search_fixed_threads(self.h, len(x), _p(x, _f), k, nprobe, _p(D, _f),
                                                  _p(I, _l), threads))
        else:
            _ck(self.lib.ref_search_fixed(self.h, len(x), _p(x, _f), k, nprobe, max_codes,
                                          _p(D, _f), _p(I, _l)))
        return D, I

    # ---- Error_sys ----
    def es_create(self, gtD, gtI):
        gtD, gtI = f32(gtD), i64(gtI)
        self.nq_total, self.max_topk = gtD.shape
        _ck(self.lib.ref_es_create(self.h, self.nq_total, self.max_topk, _p(gtD, _f), _p(gtI, _l)))

    def sys_train(self, ts, xq):
        xq = f32(xq)
        with tempfile.TemporaryDirectory() as td:
            _ck(self.lib.ref_es_sys_train(self.h, ts, _p(xq, _f), td.encode()))

    def traces(self):
        out = []
        for t in range(self.lib.ref_n_traces(self.h)):
            n = self.lib.ref_trace_size(self.h, t)
            phi, U, sg = (np.empty(n, np.float32) for _ in range(3))
            assert self.lib.ref_trace_nstd(self.h, t) == n
            self.lib.ref_get_trace(self.h, t, _p(phi, _f), _p(U, _f), _p(sg, _f))
            out.append((phi, U, sg))
        return out

    def set_traces(self, traces):
        for t, (phi, U, sg) in enumerate(traces):
            phi, U, sg = f32(phi), f32(U), f32(sg)
            self.lib.ref_set_trace(self.h, t, len(phi), _p(phi, _f), _p(U, _f), _p(sg, _f))

    def arcos(self):
        out = np.empty(500, np.float32)
        self.lib.ref_get_arcos(self.h, _p(out, _f))
        return out

    def set_queries(self, query_topk, num, xq, acc, multipler=1.0, std_m=1.0, profile=False,
                    overhead_profile=False):
        xq, acc = f32(xq), f32(acc)
        _ck(self.lib.ref_es_set_queries(self.h, query_topk, num, _p(xq, _f), len(xq), _p(acc, _f),
                                        len(acc), multipler, std_m, int(profile),
                                        int(overhead_profile)))

    def es_search(self, start, num, search_size=-1, threads=0, virtual_clock=False, chunk=0):
        """search_size=-1: one batched call over `num` queries; 1: the eval/bound.cpp loop."""
        if virtual_clock:  # error_pro::time_tune left set by time_search: the cut needs the test clock
            self.lib.ref_set_virtual_clock(1)
            try:
                return self.es_search(start, num, search_size, threads)
            finally:
                self.lib.ref_set_virtual_clock(0)
        k = self.max_topk
        D = np.empty((num, k), np.float32)
        I = np.empty((num, k), np.int64)
        if threads > 1:
            _ck(self.lib.ref_es_search_threads_chunk(self.h, _p(D, _f), _p(I, _l), start, num, threads, chunk))
        elif search_size == -1:
            _ck(self.lib.ref_es_search(self.h, _p(D, _f), _p(I, _l), start, -1))
        else:
            for q in range(0, num, search_size):
                m = min(search_size, num - q)
                _ck(self.lib.ref_es_search(self.h, _p(D[q:], _f), _p(I[q:], _l), start + q, m))
        return D, I

    def es_time_search(self, start, num, virtual_clock=True, keep_flag=False):
        """Error_sys::time_search (profile.cpp:229-244); require_acc = budget in ms.  With the
        virtual clock every IndexIVF::time() call advances by exactly 1 s (ref_driver.cpp)."""
        k = self.max_topk
        D = np.empty((num, k), np.float32)
        I = np.empty((num, k), np.int64)
        self.lib.ref_set_virtual_clock(int(virtual_clock))
        try:
            _ck(self.lib.ref_es_time_search(self.h, _p(D, _f), _p(I, _l), start, num, 0 if keep_flag else 1))
        finally:
            self.lib.ref_set_virtual_clock(0)
        return D, I

    def range_search(self, x, radius, nprobe):
        """IndexIVF::range_search (IndexIVF.cpp:741-860) -> (lims, D, I)."""
        x = f32(x)
        lims = np.empty(len(x) + 1, np.int64)
        _ck(self.lib.ref_range_search(self.h, len(x), _p(x, _f), radius, nprobe, _p(lims, _l)))
        D = np.empty(int(lims[-1]), np.float32)
        I = np.empty(int(lims[-1]), np.int64)
        self.lib.ref_range_results(int(lims[-1]), _p(D, _f), _p(I, _l))
        return lims, D, I

    def my_nprobe(self, start, n):
        out = np.empty(n, np.uint64)
        self.lib.ref_get_my_nprobe(self.h, start, n, _p(out, _ul))
        return out

    def clear_my_nprobe(self):
        self.lib.ref_clear_my_nprobe(self.h)

    def t_recalls(self, start, n):
        out = np.empty(n, np.float32)
        self.lib.ref_get_t_recalls(self.h, start, n, _p(out, _f))
        return out

    @staticmethod
    def stats():
        out = (C.c_double * 6)()
        ref().ref_stats_get(out)
        return dict(zip(["nq", "nlist", "ndis", "nheap_updates", "quantization_ms", "search_ms"], out))

    @staticmethod
    def stats_reset():
        ref().ref_stats_reset()


def ref_shards_search(subs, x, k, nprobe, threaded=False, replicas=False):
    x = f32(x)
    arr = (C.c_void_p * len(subs))(*[s.h for s in subs])
    D = np.empty((len(x), k), np.float32)
    I = np.empty((len(x), k), np.int64)
    fn = ref().ref_replicas_search if replicas else ref().ref_shards_search
    _ck(fn(arr, len(subs), len(x), _p(x, _f), k, nprobe, _p(D, _f), _p(I, _l), int(threaded)))
    return D, I


# ----------------------------------------------------------------------------------
# The restatement, assembled into the same object model as the reference.
# ----------------------------------------------------------------------------------

def fvec_L2sqr(x, y):
    x, y = f32(x), f32(y)
    return orc().orc_fvec_L2sqr(_p(x, _f), _p(y, _f), len(x))


def fvec_inner_product(x, y):
    x, y = f32(x), f32(y)
    return orc().orc_fvec_inner_product(_p(x, _f), _p(y, _f), len(x))


def construct_arcos(size=500):
    out = np.empty(size, np.float32)
    orc().orc_construct_arcos(size, _p(out, _f))
    return out


def merge_tables(metric, all_D, all_I, translations=None):
    """all_D/all_I: (nshard, n, k)."""
    all_D, all_I = f32(all_D), i64(all_I)
    nshard, n, k = all_D.shape
    tr = i64(np.zeros(nshard) if translations is None else translations)
    D = np.empty((n, k), np.float32)
    I = np.empty((n, k), np.int64)
    orc().orc_merge_tables(metric, n, k, nshard, _p(D, _f), _p(I, _l), _p(all_D, _f), _p(all_I, _l),
                           _p(tr, _l))
    return D, I


class OracleIndex:
    """IndexIVFFlat + Error_sys restated (oracle/auncel_oracle.c)."""

    def __init__(self, d, nlist, metric=L2):
        self.lib = orc()
        self.d, self.nlist, self.metric = d, nlist, metric
        self.centroids = None
        self.interdis = None
        self.arcos = construct_arcos(500)
        self.traces = None
        self._x = []
        self._ids = []
        self._list_no = []
        self.ntotal = 0
        self._csr = None
        self.multipler, self.std_m = 1.0, 1.0

    def set_centroids(self, c):
        self.centroids = f32(c)
        n = self.nlist
        self.interdis = np.empty(n * (n - 1) // 2, np.float32)
        self.lib.orc_interdis(self.metric, n, self.d, _p(self.centroids, _f), _p(self.interdis, _f))

    def coarse(self, x, nprobe):
        x = f32(x)
        dis = np.empty((len(x), nprobe), np.float32)
        keys = np.empty((len(x), nprobe), np.int64)
        self.lib.orc_coarse(self.metric, len(x), _p(x, _f), self.nlist, _p(self.centroids, _f),
                            self.d, nprobe, _p(dis, _f), _p(keys, _l))
        return dis, keys

    def assign(self, x):
        return self.coarse(x, 1)[1][:, 0].copy()

    def add(self, x, ids=None, list_no=None):
        x = f32(x)
        if list_no is None:
            list_no = self.assign(x)
        if ids is None:
            ids = np.arange(self.ntotal, self.ntotal + len(x), dtype=np.int64)
        self._x.append(x)
        self._ids.append(i64(ids))
        self._list_no.append(i64(list_no))
        self.ntotal += len(x)
        self._csr = None

    def csr(self):
        if self._csr is None:
            x = np.concatenate(self._x) if self._x else np.zeros((0, self.d), np.float32)
            ids = np.concatenate(self._ids) if self._ids else np.zeros(0, np.int64)
            ln = np.concatenate(self._list_no) if self._list_no else np.zeros(0, np.int64)
            keep = ln >= 0
            x, ids, ln = x[keep], ids[keep], ln[keep]
            order = np.argsort(ln, kind="stable")
            counts = np.bincount(ln, minlength=self.nlist)
            off = np.zeros(self.nlist + 1, np.int64)
            np.cumsum(counts, out=off[1:])
            self._csr = (f32(x[order]), off, i64(ids[order]))
        return self._csr

    def _model_arrays(self):
        if self.traces is None:
            off = np.zeros(1, np.int64)
            z = np.zeros(1, np.float32)
            return 0, off, z, z, z
        off = np.zeros(len(self.traces) + 1, np.int64)
        np.cumsum([len(t[0]) for t in self.traces], out=off[1:])
        phi = f32(np.concatenate([t[0] for t in self.traces]))
        U = f32(np.concatenate([t[1] for t in self.traces]))
        sg = f32(np.concatenate([t[2] for t in self.traces]))
        return len(self.traces), off, phi, U, sg

    def search_preassigned(self, x, k, keys, coarse_dis, mode=0, max_codes=0, offset=0,
                           query_topk=0, require_acc=None, gt_D=None, profile=False,
                           overhead_profile=False, my_nprobe=None, t_recalls=None,
                           train_pairs=None, train_num=0, dump_q=-1):
        x, keys, coarse_dis = f32(x), i64(keys), f32(coarse_dis)
        n, nprobe = keys.shape
        codes, off, ids = self.csr()
        ntr, toff, phi, U, sg = self._model_arrays()
        D = np.empty((n, k), np.float32)
        I = np.empty((n, k), np.int64)
        stats = np.zeros(3, np.int64)
        dump = np.full((nprobe, 4), -2, np.float32) if dump_q >= 0 else None
        require_acc = None if require_acc is None else f32(require_acc)
        gt_D = None if gt_D is None else f32(gt_D)
        err = self.lib.orc_search_preassigned(
            self.metric, self.d, self.nlist, _p(codes, _f), _p(off, _l), _p(ids, _l), n, _p(x, _f),
            k, nprobe, max_codes, _p(keys, _l), _p(coarse_dis, _f), mode, offset,
            _p(self.interdis, _f), _p(self.arcos, _f), len(self.arcos), ntr, _p(toff, _l),
            _p(phi, _f), _p(U, _f), _p(sg, _f), self.multipler, self.std_m, query_topk,
            _p(require_acc, _f), _p(gt_D, _f), int(profile), int(overhead_profile),
            _p(my_nprobe, _ul), _p(t_recalls, _f), _p(train_pairs, _f), train_num, _p(D, _f),
            _p(I, _l), _p(stats, _l), dump_q, _p(dump, _f))
        self.last_err = err
        self.last_stats = dict(nlist=int(stats[0]), ndis=int(stats[1]), nheap_updates=int(stats[2]))
        self.last_dump = dump
        return D, I

    def search_fixed(self, x, k, nprobe, max_codes=0):
        dis, keys = self.coarse(x, nprobe)
        return self.search_preassigned(x, k, keys, dis, 0, max_codes)

    def search_timed(self, x, k, budget_ms, us_per_list, ns_per_code, offset=0):
        """Error_sys::time_search (profile.cpp:229-244): nprobe = nlist, no tune block, the
        latency cut of IndexIVF.cpp:545-549 on the modelled clock (orc_set_time_tune)."""
        dis, keys = self.coarse(x, self.nlist)
        self.lib.orc_set_time_tune(1, int(us_per_list), int(ns_per_code))
        try:
            return self.search_preassigned(x, k, keys, dis, 0, 0, offset=offset, require_acc=budget_ms)
        finally:
            self.lib.orc_set_time_tune(0, 0, 0)

    def range_search(self, x, radius, nprobe):
        """IndexIVF::range_search (IndexIVF.cpp:741-860) -> (lims, D, I), hits in scan order."""
        x = f32(x)
        dis, keys = self.coarse(x, nprobe)
        keys = i64(keys)
        codes, off, ids = self.csr()
        lims = np.empty(len(x) + 1, np.int64)
        stats = np.zeros(2, np.int64)
        args = (self.metric, self.d, _p(codes, _f), _p(off, _l), _p(ids, _l), len(x), _p(x, _f), radius, nprobe,
                _p(keys, _l), _p(lims, _l))
        self.lib.orc_range_search(*args, None, None, _p(stats, _l))
        D = np.empty(int(lims[-1]), np.float32)
        I = np.empty(int(lims[-1]), np.int64)
        self.lib.orc_range_search(*args, _p(D, _f), _p(I, _l), None)
        self.last_stats = dict(nlist=int(stats[0]), ndis=int(stats[1]))
        return lims, D, I

    def n_traces(self):
        n, np_ = 0, 1
        while np_ <= self.nlist // 8:
            n += 1
            np_ <<= 1
        return n

    def calibrate(self, xq, gt_D, bs=250):
        """Error_sys::sys_train (profile.cpp:88-171) + error_pro::train (IVF_pro.cpp:186-194).
        Returns (D, I) of the calibration search (top-k after nlist/8 lists)."""
        xq, gt_D = f32(xq), f32(gt_D)
        ts, k = gt_D.shape
        ntr = self.n_traces()
        per = (k // 4) * ts
        pairs = np.full((ntr, per, 2), -1, np.float32)  # IndexIVF.cpp:213-217
        dis, keys = self.coarse(xq, self.nlist)
        D, I = self.search_preassigned(xq, k, keys, dis, mode=2, gt_D=gt_D, train_pairs=pairs,
                                       train_num=ts)
        self.raw_pairs = pairs.copy()
        traces = []
        for t in range(ntr):
            cap = (per + bs - 1) // bs + 1
            phi, U, sg = (np.empty(cap, np.float32) for _ in range(3))
            p = f32(pairs[t].reshape(-1))
            sz = self.lib.orc_trace_SB(_p(p, _f), per, bs, _p(phi, _f), _p(U, _f), _p(sg, _f))
            traces.append((phi[:sz].copy(), U[:sz].copy(), sg[:sz].copy()))
        self.traces = traces
        return D, I

    def search_bounded(self, x, max_topk, query_topk, require_acc, gt_D=None, offset=0,
                       my_nprobe=None, profile=False, overhead_profile=False, dump_q=-1, time_model=None):
        """Error_sys::search (profile.cpp:211-227): nprobe = nlist, tune block on.
        require_acc / gt_D / my_nprobe are indexed by GLOBAL id (offset + i)."""
        x = f32(x)
        n = len(x)
        dis, keys = self.coarse(x, self.nlist)
        if my_nprobe is None:
            my_nprobe = np.zeros(offset + n, np.uint64)
        t_rec = np.zeros(offset + n, np.float32)
        if time_model is not None:  # error_pro::time_tune still set (profile.cpp:242)
            self.lib.orc_set_time_tune(1, int(time_model[0]), int(time_model[1]))
        try:
            D, I = self.search_preassigned(x, max_topk, keys, dis, mode=1, offset=offset,
                                           query_topk=query_topk, require_acc=require_acc, gt_D=gt_D,
                                           profile=profile, overhead_profile=overhead_profile,
                                           my_nprobe=my_nprobe, t_recalls=t_rec, dump_q=dump_q)
        finally:
            self.lib.orc_set_time_tune(0, 0, 0)
        return D, I, my_nprobe, t_rec
