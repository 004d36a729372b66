"""GPU: Index::train (k-means on the device, csrc/kmeans.cu) against the UNMODIFIED reference's
Clustering::train (Auncel/Clustering.cpp:77-244, km_update_centroids utils.cpp:1078-1161):
same seed -> bit-identical centroids -- sub-sampling (rand_perm over mt19937), void-cluster
splits, spherical k-means for inner product, dimensions that are not a multiple of 4 or 32."""
import numpy as np
import pytest

import auncel_b200 as ab
from auncel_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _train_both(metric, d, nlist, xb, niter):
    if not O.have_ref():
        pytest.skip("oracle/_ref/libauncel_ref.so not present")
    O.RefIndex.set_blas_threshold(1 << 30)  # index.search(nx, x, 1) on the exact path, like the device
    R = O.RefIndex(d, nlist, metric)
    R.train(xb, niter=niter)
    ix = ab.IndexIVFFlat(d, nlist, metric)
    ix.set_tune_mode()
    ix.train(xb, niter=niter)
    ix.set_tune_off()
    cr, cg = R.centroids(), ix.centroids()
    inter_r, inter_g = R.interdis(), ix.interdis_cem()
    R.close()
    return cr, cg, inter_r, inter_g


@pytest.mark.parametrize("metric,d,nlist,nb,niter", [
    (O.L2, 16, 64, 8000, 6),        # plain
    (O.L2, 37, 48, 5000, 4),        # d not a multiple of 4
    (O.L2, 300, 32, 3000, 3),       # more than one 256-dimension pass of the accumulation
    (O.IP, 24, 64, 9000, 5),        # spherical (IndexIVF.cpp:160-162)
    (O.L2, 8, 16, 16 * 256 + 777, 5),  # more than 256 points per centroid: seeded sub-sampling
])
def test_kmeans_centroids_bit_equal(metric, d, nlist, nb, niter):
    xb = synth.clustered(41, nb, d, 30, 0.4, normalize=metric == O.IP)
    cr, cg, ir, ig = _train_both(metric, d, nlist, xb, niter)
    assert np.array_equal(cr, cg)
    assert np.array_equal(ir, ig)  # train_q1's interdis_cem on the trained centroids (IndexIVF.cpp:97-117)


def test_kmeans_void_clusters_split_like_reference():
    """Fewer distinct points than centroids: void clusters every iteration, split by the seeded draw
    of utils.cpp:1121-1157."""
    d, nlist = 12, 64
    proto = synth.clustered(43, 40, d, 40, 0.5)
    xb = np.repeat(proto, 60, axis=0)[np.random.default_rng(3).permutation(2400)]
    cr, cg, _, _ = _train_both(O.L2, d, nlist, np.ascontiguousarray(xb), 5)
    assert np.array_equal(cr, cg)


def test_train_rejects_bad_input():
    ix = ab.IndexIVFFlat(8, 16)
    with pytest.raises(ab.FaissException):
        ix.train(synth.clustered(1, 10, 8))  # fewer points than centroids (Clustering.cpp:78-80)
    x = synth.clustered(1, 100, 8)
    x[17, 3] = np.inf
    with pytest.raises(ab.FaissException):
        ix.train(x)  # Clustering.cpp:86-89
