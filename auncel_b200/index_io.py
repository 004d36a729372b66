"""write_index / read_index for IndexIVFFlat in the reference's on-disk format
(/root/reference/Auncel/index_io.cpp): fourcc "IwFl" (:448-453) = index header (:196-203) + nlist,
nprobe + the quantizer as "IxF2"/"IxFI" (:384-390) + direct map + ArrayInvertedLists "ilar"
(:280-330, "full" or "sprs" size table, then per non-empty list codes and ids).

The reference does not persist Auncel's state (interdis_cem, traces, hyper-parameters):
a reloaded index cannot run bounded search without retraining.  Here that state is appended
after the standard payload in a block tagged "AuNc"; the reference's reader stops before it, so
files stay loadable by the reference, and the reference's own files (e.g. the shipped
eval/trained_index/*_IVF1024,Flat_trained.index) load here.
"""
import struct

import numpy as np


def _fourcc(s):
    return struct.unpack("<I", s.encode())[0]


class _R:
    def __init__(self, b):
        self.b, self.o = b, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def arr(self, dtype, n):
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o).copy()
        self.o += a.nbytes
        return a

    def vec(self, dtype):
        return self.arr(dtype, self.take("q"))


def _read_header(r):
    d, ntotal, _, _, trained, metric = r.take("i"), r.take("q"), r.take("q"), r.take("q"), r.take("?"), r.take("i")
    return d, ntotal, trained, metric


def parse_ivfflat(data):
    """bytes of an "IwFl" file -> dict; pure numpy (no GPU needed)."""
    r = _R(data)
    h = r.take("I")
    if h != _fourcc("IwFl"):
        raise ValueError("not an IndexIVFFlat file (fourcc %r)" % struct.pack("<I", h))
    d, ntotal, trained, metric = _read_header(r)
    nlist, nprobe = r.take("Q"), r.take("Q")
    hq = r.take("I")
    if hq not in (_fourcc("IxF2"), _fourcc("IxFI")):
        raise ValueError("quantizer is not an IndexFlat")
    dq, nq, _, mq = _read_header(r)
    xb = r.vec(np.float32).reshape(nq, dq)
    r.take("?")            # maintain_direct_map
    r.vec(np.int64)        # direct_map
    hl = r.take("I")
    sizes = np.zeros(nlist, np.int64)
    codes = np.zeros((0, d), np.float32)
    ids = np.zeros(0, np.int64)
    if hl == _fourcc("ilar"):
        nl, code_size = r.take("Q"), r.take("Q")
        assert nl == nlist and code_size == 4 * d
        lt = r.take("I")
        tab = r.vec(np.uint64).astype(np.int64)
        if lt == _fourcc("full"):
            sizes = tab
        elif lt == _fourcc("sprs"):
            sizes[tab[0::2]] = tab[1::2]
        else:
            raise ValueError("unknown list type")
        cl, il = [], []
        for n in sizes:
            if n > 0:
                cl.append(r.arr(np.float32, int(n) * d).reshape(int(n), d))
                il.append(r.arr(np.int64, int(n)))
        if cl:
            codes, ids = np.concatenate(cl), np.concatenate(il)
    elif hl != _fourcc("il00"):
        raise ValueError("unsupported inverted lists")
    out = dict(d=d, ntotal=ntotal, is_trained=trained, metric=metric, nlist=nlist, nprobe=nprobe, centroids=xb,
               list_sizes=sizes, codes=codes, ids=ids, auncel=None)
    if r.o + 4 <= len(data) and r.take("I") == _fourcc("AuNc"):
        has_inter, mult, stdm, ntr = r.take("i"), r.take("f"), r.take("f"), r.take("i")
        interdis = r.vec(np.float32) if has_inter else None
        traces = [(r.vec(np.float32), r.vec(np.float32), r.vec(np.float32)) for _ in range(ntr)]
        out["auncel"] = dict(interdis=interdis, multipler=mult, std_m=stdm, traces=traces)
    return out


def serialize_ivfflat(d, metric, nlist, nprobe, ntotal, centroids, list_sizes, codes, ids, auncel=None):
    def header(dd, nt, trained, m):
        return struct.pack("<iqqq?i", dd, nt, 1 << 20, 1 << 20, trained, m)

    def vec(a):
        a = np.ascontiguousarray(a)
        return struct.pack("<Q", a.size) + a.tobytes()

    out = [struct.pack("<I", _fourcc("IwFl")), header(d, ntotal, True, metric), struct.pack("<QQ", nlist, nprobe),
           struct.pack("<I", _fourcc("IxF2" if metric == 1 else "IxFI")), header(d, nlist, True, metric),
           vec(np.asarray(centroids, np.float32)), struct.pack("<?", False), vec(np.zeros(0, np.int64))]
    sizes = np.asarray(list_sizes, np.int64)
    out.append(struct.pack("<IQQ", _fourcc("ilar"), nlist, 4 * d))
    if (sizes > 0).sum() > nlist // 2:  # index_io.cpp:293-313
        out += [struct.pack("<I", _fourcc("full")), vec(sizes.astype(np.uint64))]
    else:
        nz = np.flatnonzero(sizes)
        tab = np.empty(2 * len(nz), np.uint64)
        tab[0::2], tab[1::2] = nz, sizes[nz]
        out += [struct.pack("<I", _fourcc("sprs")), vec(tab)]
    off = 0
    for n in sizes:
        if n > 0:
            out += [np.ascontiguousarray(codes[off:off + n], np.float32).tobytes(),
                    np.ascontiguousarray(ids[off:off + n], np.int64).tobytes()]
            off += int(n)
    if auncel is not None:
        inter = auncel.get("interdis")
        out.append(struct.pack("<Iiffi", _fourcc("AuNc"), int(inter is not None), auncel["multipler"], auncel["std_m"],
                               len(auncel["traces"])))
        if inter is not None:
            out.append(vec(np.asarray(inter, np.float32)))
        for phi, U, sg in auncel["traces"]:
            out += [vec(np.asarray(phi, np.float32)), vec(np.asarray(U, np.float32)), vec(np.asarray(sg, np.float32))]
    return b"".join(out)


def write_index(ix, fname, auncel_state=True):
    """faiss::write_index (index_io.cpp:383+) for an auncel_b200.IndexIVFFlat."""
    import ctypes as C

    from ._lib import _f, _l, lib
    from .index import _ck
    nt = int(ix.list_sizes().sum())
    codes = np.empty((nt, ix.d), np.float32)
    ids = np.empty(nt, np.int64)
    _ck(lib().auncel_index_get_lists(ix.h, codes.ctypes.data_as(_f), ids.ctypes.data_as(_l)))
    st = None
    if auncel_state:
        m, s = C.c_float(), C.c_float()
        lib().auncel_index_get_params(ix.h, C.byref(m), C.byref(s))
        st = dict(interdis=ix.interdis_cem() if lib().auncel_index_has_interdis(ix.h) else None,
                  multipler=m.value, std_m=s.value, traces=ix.traces())
    with open(fname, "wb") as f:
        f.write(serialize_ivfflat(ix.d, ix.metric_type, ix.nlist, ix.nprobe, ix.ntotal, ix.centroids(),
                                  ix.list_sizes(), codes, ids, st))


def read_index(fname, device=0):
    """faiss::read_index (index_io.cpp:890+) -> auncel_b200.IndexIVFFlat on `device`."""
    from .index import IndexIVFFlat
    p = parse_ivfflat(open(fname, "rb").read())
    ix = IndexIVFFlat(p["d"], p["nlist"], p["metric"], device)
    ix.nprobe = p["nprobe"]
    a = p["auncel"]
    ix.set_centroids(p["centroids"], compute_interdis=False)
    if a is not None and a["interdis"] is not None:
        from ._lib import _f, lib
        from .index import _ck
        inter = np.ascontiguousarray(a["interdis"], np.float32)
        _ck(lib().auncel_index_set_interdis(ix.h, inter.ctypes.data_as(_f)))
    if len(p["ids"]):
        ix.add_core(p["codes"], p["ids"], np.repeat(np.arange(p["nlist"]), p["list_sizes"]))
    if a is not None and a["traces"]:
        ix.set_error_model(a["traces"], a["multipler"], a["std_m"])
    return ix
