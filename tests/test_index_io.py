"""index_io format: CPU parsing of the reference's shipped trained indexes (only where the
reference tree is mounted) and byte-level round trip; GPU write/read round trip incl. the
Auncel state extension."""
import os

import numpy as np
import pytest

from auncel_b200 import index_io as IO
from auncel_b200 import synth

REF_DIR = "/root/reference/Auncel/eval/trained_index"


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference tree not mounted")
@pytest.mark.parametrize("name,d,metric", [("sift10M", 128, 1), ("deep10M", 96, 1), ("gist", 960, 1), ("text", 200, 0)])
def test_parse_reference_trained_index(name, d, metric):
    data = open(os.path.join(REF_DIR, f"{name}_IVF1024,Flat_trained.index"), "rb").read()
    p = IO.parse_ivfflat(data)
    assert (p["d"], p["nlist"], p["metric"], p["ntotal"]) == (d, 1024, metric, 0)
    assert p["centroids"].shape == (1024, d) and p["list_sizes"].sum() == 0 and p["auncel"] is None
    # re-serialising gives the same bytes: the writer matches index_io.cpp field for field
    again = IO.serialize_ivfflat(d, metric, 1024, p["nprobe"], 0, p["centroids"], p["list_sizes"], p["codes"], p["ids"])
    assert again == data
    if name == "text":
        assert np.allclose((p["centroids"] ** 2).sum(1), 1.0, atol=1e-4)


def test_serialize_parse_roundtrip_cpu():
    d, nlist = 8, 16
    cent = synth.clustered(1, nlist, d)
    sizes = np.array([3, 0, 2, 0] * 4)
    codes = synth.clustered(2, int(sizes.sum()), d)
    ids = np.arange(sizes.sum(), dtype=np.int64) * 7
    tr = [(np.array([0.1, 0.5], np.float32), np.array([1.0, 2.0], np.float32), np.array([0.0, 0.1], np.float32))]
    for aun in (None, dict(interdis=np.arange(nlist * (nlist - 1) // 2, dtype=np.float32), multipler=2.5, std_m=1.5, traces=tr)):
        b = IO.serialize_ivfflat(d, 1, nlist, 4, 20, cent, sizes, codes, ids, aun)
        p = IO.parse_ivfflat(b)
        assert np.array_equal(p["centroids"], cent) and np.array_equal(p["codes"], codes) and np.array_equal(p["ids"], ids)
        assert np.array_equal(p["list_sizes"], sizes) and p["nprobe"] == 4
        assert (p["auncel"] is None) == (aun is None)
        if aun:
            assert p["auncel"]["multipler"] == 2.5 and np.array_equal(p["auncel"]["traces"][0][1], tr[0][1])


@pytest.mark.gpu
def test_write_read_roundtrip_gpu(tmp_path):
    import auncel_b200 as ab
    from oracle import oracle as O
    d, nlist, nb, k, qk = 16, 64, 8000, 16, 4
    xb = synth.clustered(3, nb, d, 40)
    xq = synth.clustered(21, 100, d, 40)
    ix = ab.IndexIVFFlat(d, nlist)
    ix.set_centroids(synth.clustered(53, nlist, d, 40))
    ix.add(xb)
    ix.nprobe = nlist
    gD, gI = ix.search(xq, k)
    es = ab.Error_sys(ix, 100, k)
    es.set_gt(gD, gI)
    es.sys_train(50, xq)
    ix.set_params(2.0, 1.0)
    acc = np.full(50, 0.9, np.float32)
    D1, I1, np1 = ix.search_bounded(xq[50:], k, qk, acc)
    f = str(tmp_path / "ivf.index")
    IO.write_index(ix, f)
    ix2 = IO.read_index(f)
    assert ix2.ntotal == nb and np.array_equal(ix2.list_sizes(), ix.list_sizes())
    D2, I2, np2 = ix2.search_bounded(xq[50:], k, qk, acc)  # bounded search works right after loading
    assert np.array_equal(D1, D2) and np.array_equal(I1, I2) and np.array_equal(np1, np2)
    p = IO.parse_ivfflat(open(f, "rb").read())
    assert p["auncel"] is not None and len(p["auncel"]["traces"]) == len(ix.traces())


def test_vecs_roundtrip(tmp_path):
    from auncel_b200 import vecs_io
    x = synth.clustered(9, 37, 12)
    f = str(tmp_path / "x.fvecs")
    vecs_io.fvecs_write(f, x)
    assert np.array_equal(vecs_io.fvecs_read(f), x)
    ids = (np.arange(37 * 5).reshape(37, 5) * 3).astype(np.int32)
    g = str(tmp_path / "i.ivecs")
    vecs_io.fvecs_write(g, ids)
    assert np.array_equal(vecs_io.ivecs_read(g), ids)
