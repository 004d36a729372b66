"""Build libauncel_b200.so in-tree with nvcc for sm_100a (no torch dependency)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libauncel_b200.so")
SOURCES = ["coarse.cu", "scan.cu", "tcfilter.cu", "tcfilter2.cu", "tcfilter3.cu", "merge.cu", "merge_tables.cu", "kmeans.cu", "range.cu", "shards.cu", "shard_rounds.cu", "index.cu", "c_api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",  # the reference build has no FMA; every product/sum rounds separately
         "-Xcompiler", "-fPIC,-O2,-fno-fast-math", "-ccbin", "/usr/bin/g++"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "auncel_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "_obj", src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
        ok = ok and p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++"])
    return LIB


def build_examples():
    """C++ driver written against include/auncel/faiss_api.h (the reference-API mirror)."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "examples", "bound_demo")
    src = os.path.join(root, "examples", "bound_demo.cpp")
    hdrs = [os.path.join(root, "include", "auncel", "faiss_api.h"), os.path.join(root, "include", "auncel_b200.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(f) for f in [src, LIB] + hdrs):
        return out
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++14", "-I", os.path.join(root, "include"), src, "-o", out,
                           LIB, "-Wl,-rpath," + HERE, "-lpthread"])
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
