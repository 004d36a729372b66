"""ctypes loader for the C ABI (include/auncel_b200.h).  No fallback: if the CUDA library
was not built, importing the product fails loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AUNCEL_LIB", os.path.join(HERE, "libauncel_b200.so"))  # override: kernel experiments

_f = C.POINTER(C.c_float)
_l = C.POINTER(C.c_int64)
_u = C.POINTER(C.c_uint64)
_d = C.POINTER(C.c_double)
_h = C.c_void_p

# every symbol declared in include/auncel_b200.h: name -> (restype, argtypes)
SYMBOLS = {
    "auncel_get_last_error": (C.c_char_p, []),
    "auncel_index_new": (C.c_int, [C.POINTER(_h), C.c_int, C.c_int64, C.c_int, C.c_int]),
    "auncel_index_free": (None, [_h]),
    "auncel_index_d": (C.c_int, [_h]),
    "auncel_index_nlist": (C.c_int64, [_h]),
    "auncel_index_ntotal": (C.c_int64, [_h]),
    "auncel_index_is_trained": (C.c_int, [_h]),
    "auncel_index_wait_stream": (C.c_int, [_h, C.c_void_p]),
    "auncel_index_set_centroids": (C.c_int, [_h, _f, C.c_int]),
    "auncel_index_get_centroids": (C.c_int, [_h, _f]),
    "auncel_index_get_interdis": (C.c_int, [_h, _f]),
    "auncel_index_set_interdis": (C.c_int, [_h, _f]),
    "auncel_index_train": (C.c_int, [_h, C.c_int64, _f, C.c_int, C.c_int]),
    "auncel_index_add": (C.c_int, [_h, C.c_int64, _f, _l, _l]),
    "auncel_index_add_device": (C.c_int, [_h, C.c_int64, C.c_void_p, _l, _l]),
    "auncel_index_assign": (C.c_int, [_h, C.c_int64, _f, _l]),
    "auncel_index_reset": (C.c_int, [_h]),
    "auncel_index_list_sizes": (C.c_int, [_h, _l]),
    "auncel_index_get_lists": (C.c_int, [_h, _f, _l]),
    "auncel_index_get_params": (C.c_int, [_h, _f, _f]),
    "auncel_index_has_interdis": (C.c_int, [_h]),
    "auncel_index_coarse_search": (C.c_int, [_h, C.c_int64, _f, C.c_int64, _f, _l]),
    "auncel_index_search": (C.c_int, [_h, C.c_int64, _f, C.c_int64, C.c_int64, C.c_int64, _f, _l]),
    "auncel_index_search_device": (C.c_int, [_h, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                             C.c_void_p, C.c_void_p]),
    "auncel_index_set_error_model": (C.c_int, [_h, C.c_int, _l, _f, _f, _f, C.c_float, C.c_float]),
    "auncel_index_set_params": (C.c_int, [_h, C.c_float, C.c_float]),
    "auncel_index_n_traces": (C.c_int, [_h]),
    "auncel_index_trace_size": (C.c_int64, [_h, C.c_int]),
    "auncel_index_get_trace": (C.c_int, [_h, C.c_int, _f, _f, _f]),
    "auncel_index_calibrate": (C.c_int, [_h, C.c_int64, _f, C.c_int64, _f, _f, _l]),
    "auncel_index_search_bounded": (C.c_int, [_h, C.c_int64, _f, C.c_int64, C.c_int64, _f, _f, _u, _f,
                                              C.c_int, _f, _l]),
    "auncel_index_search_bounded_device": (C.c_int, [_h, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_int, C.c_void_p, C.c_void_p]),
    "auncel_index_set_time_model": (C.c_int, [_h, C.c_int64, C.c_int64]),
    "auncel_index_search_timed": (C.c_int, [_h, C.c_int64, _f, C.c_int64, _f, _f, _l]),
    "auncel_index_search_timed_device": (C.c_int, [_h, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                   C.c_void_p]),
    "auncel_index_range_search": (C.c_int, [_h, C.c_int64, _f, C.c_float, C.c_int64, _l]),
    "auncel_index_range_search_results": (C.c_int, [_h, _f, _l]),
    "auncel_index_get_stats": (C.c_int, [_h, _d]),
    "auncel_index_get_round_stats": (C.c_int, [_h, C.c_int, _d, C.POINTER(C.c_int)]),
    "auncel_index_set_pool_budget": (C.c_int, [_h, C.c_size_t]),
    "auncel_index_set_option": (C.c_int, [_h, C.c_char_p, C.c_int]),
    "auncel_merge_tables": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, _f, _l, _l, _f, _l]),
    "auncel_heap_entry_table": (C.c_int, [C.c_int64, C.POINTER(C.c_int32)]),
    "auncel_merge_tables_device": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "auncel_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "auncel_shard_group_new": (C.c_int, [C.POINTER(_h), _h, C.c_int, C.c_int, C.c_void_p]),
    "auncel_shard_group_free": (None, [_h]),
    "auncel_shard_group_search": (C.c_int, [_h, C.c_int64, _f, C.c_int64, C.c_int64, C.c_int64, _f, _l]),
    "auncel_shard_group_search_device": (C.c_int, [_h, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                                   C.c_void_p, C.c_void_p]),
    "auncel_shard_group_get_stats": (C.c_int, [_h, _d]),
    "auncel_shard_group_set_bounded": (C.c_int, [_h, C.c_int]),
    "auncel_shard_group_get_exchange_stats": (C.c_int, [_h, _d]),
    "auncel_index_copy_subset_to": (C.c_int, [_h, _h, C.c_int, C.c_int64, C.c_int64]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
