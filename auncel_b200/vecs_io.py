"""fvecs / ivecs / fbin readers and writers of the reference's drivers
(/root/reference/Auncel/eval/bound.cpp:29-114, dist/gt.cpp): every vector is stored as an int32
dimension followed by d 4-byte components; .fbin = int32 n, int32 d, then n*d floats."""
import numpy as np


def fvecs_read(fname, dtype=np.float32):
    a = np.fromfile(fname, dtype=np.int32)
    if a.size == 0:
        return np.zeros((0, 0), dtype)
    d = int(a[0])
    a = a.reshape(-1, d + 1)
    if not np.all(a[:, 0] == d):
        raise ValueError("inconsistent vector dimensions in " + fname)
    return np.ascontiguousarray(a[:, 1:]).view(dtype)


def ivecs_read(fname):
    return fvecs_read(fname, np.int32)


def fvecs_write(fname, x):
    x = np.ascontiguousarray(x)
    assert x.dtype.itemsize == 4 and x.ndim == 2
    n, d = x.shape
    out = np.empty((n, d + 1), np.int32)
    out[:, 0] = d
    out[:, 1:] = x.view(np.int32)
    out.tofile(fname)


def fbin_read(fname, dtype=np.float32, max_n=None):
    with open(fname, "rb") as f:
        n, d = np.fromfile(f, dtype=np.int32, count=2)
        if max_n is not None:
            n = min(int(n), max_n)
        return np.fromfile(f, dtype=dtype, count=int(n) * int(d)).reshape(int(n), int(d))


def ivecs_write(fname, x):
    fvecs_write(fname, np.ascontiguousarray(x, dtype=np.int32))
