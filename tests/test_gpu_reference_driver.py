"""GPU: the reference's own evaluation driver, Auncel/eval/bound.cpp (:216-426), compiled
UNMODIFIED against include/auncel/faiss_api.h (oracle/Makefile, target _ref/bound_b200) and run on
synthetic fvecs/ivecs files.  It trains "IVF1024,Flat" through index_factory, writes the trained
index (write_index), adds the base, calibrates (Error_sys::sys_train), searches in latency mode
(one query per call) and prints its own verdict: "Error bound is guaranteed"."""
import os
import subprocess

import numpy as np
import pytest

import auncel_b200 as ab
from auncel_b200 import index_io, vecs_io
from tests.util import mixture

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOUND = os.path.join(ROOT, "oracle", "_ref", "bound_b200")
REMAP = os.path.join(ROOT, "oracle", "_ref", "libpathremap.so")


def test_reference_bound_driver_unmodified(tmp_path):
    if not (os.path.exists(BOUND) and os.path.exists(REMAP)):
        pytest.skip("oracle/_ref/bound_b200 was not built (needs /root/reference at build time)")
    d, nb, ts, ses, k = 32, 120_000, 200, 100, 100
    xb = mixture(31, nb, d)
    xq = mixture(32, ts + ses, d)
    # ground truth by exhaustive search with the exact kernel (== brute force with fvec_L2sqr)
    ix = ab.IndexIVFFlat(d, 64)
    ix.train(xb[::40], niter=2)
    ix.add(xb)
    ix.nprobe = 64
    gD, gI = ix.search(xq, k)
    del ix
    data = tmp_path / "data" / "sift"
    run = tmp_path / "run" / "eval"
    os.makedirs(data)
    os.makedirs(run / "trained_index")
    vecs_io.fvecs_write(str(data / "sift1M.fvecs"), xb)
    vecs_io.fvecs_write(str(data / "1M_query.fvecs"), xq)
    vecs_io.ivecs_write(str(data / "idx_1M.ivecs"), gI.astype(np.int32))
    vecs_io.fvecs_write(str(data / "dis_1M.fvecs"), gD)
    # ../hyperparameter.txt relative to the driver's cwd (IVF_pro.cpp:240-256): line 6 = SIFT10M, k=10, eb=0.1
    with open(tmp_path / "run" / "hyperparameter.txt", "w") as f:
        f.write("\n".join(["9.3 1.0", "6.9 1.0", "2.7 12.0", "11.0 8.0", "6.7 1.0", "7.9 6.0", "10.2 6.0", "26.5 1.0",
                           "10.0 0.2", "4.2 1.0", "4.5 1.0", "15.0 1.0"]) + "\n")
    env = dict(os.environ, LD_PRELOAD=REMAP, AUNCEL_DATA_ROOT=str(tmp_path / "data"))
    # ./bound <dataset> <train size> <query size> <topk> <error bound> <figure id>   (eval/run.sh)
    p = subprocess.run([BOUND, "sift1M", str(ts), str(ses), "10", "0.1", "6"], cwd=run, env=env, capture_output=True,
                       text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert "Error bound is guaranteed" in out, out[-3000:]
    assert "Output index type: 0" in out  # IndexType IVF
    # the latency log the driver writes (eval/bound.cpp:417-426) and the index it saved with write_index
    lat = np.loadtxt(run / "Auncel_Latency_sift1M_10_10.log")
    assert lat.shape == (ses,) and (lat > 0).all()
    parsed = index_io.parse_ivfflat(open(run / "trained_index" / "sift1M_IVF1024,Flat_trained.index", "rb").read())
    assert parsed["nlist"] == 1024 and parsed["d"] == d and parsed["centroids"].shape == (1024, d)
    assert parsed["ntotal"] == 0  # saved right after training, before add (eval/bound.cpp:261-268)
