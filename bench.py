#!/usr/bin/env python
"""bench.py -- QPS of the error-bounded IVF-Flat query path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (configs[1] of BASELINE.json): IVF-Flat nlist=4096, 10M x 128 synthetic
SIFT-shaped base, 10k test queries (+5k calibration queries), result heap max_topk=100,
query_topk=10, error bound 0.1 (0.05 / 0.2 are timed once each under "error_bounds").
One step = one Error_sys::search over the whole query batch.
  value  : queries/s with the queries already resident in HBM (device C-ABI entry point)
  e2e    : queries/s through the host-pointer C-ABI call (pinned host buffers; H2D of the
           queries/targets and D2H of distances/labels/my_nprobe inside the timed region)
N > 1 (torchrun, one rank per GPU): replicas -- IndexReplicas semantics
(Auncel/IndexReplicas.cpp:79-118): every rank holds the index and serves its own query batch,
no data-path collective; value = all ranks' queries / max-over-ranks time ("weak").

Explanatory sections of the same JSON line (none of them changes `value`):
  roofline     dominant kernel (tc_filter_kernel): fraction of the ACTIVE bound -- compulsory (unique)
               HBM bytes per launch / launch time against the measured copy peak, or TF32 rate against
               the measured tensor peak, chosen by arithmetic intensity; step-level floor
  hbm_bound    the regime where "fraction of HBM peak" is defined: batch 1 and 64 on the
               headline index and on config 4's TEXT shape (10M x 200, normalised inner product)
  shards       config 3: 10M x 96 DEEP-shaped inner product, k = 100, fixed nprobe, the inverted lists
               split over the N ranks (id % N), ncclAllGather of the packed tables + merge inside the
               library (IndexShards.cpp:261-311); strong scaling; result compared with the unsharded index
  cpu_baseline the unmodified reference (oracle/_ref) on all host threads, bounded sample
--impl reference: that reference arm alone, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

# stdout carries exactly one JSON line.  The unmodified reference (oracle/_ref) prints progress with
# printf; file descriptor 1 is therefore pointed at stderr and the JSON goes out through a private copy.
_JSON_OUT = None


def emit(obj):
    global _JSON_OUT
    if _JSON_OUT is None:
        _JSON_OUT = sys.stdout
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def isolate_stdout():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


sys.path.insert(0, ROOT)

QUERY_TOPK = 10
MAX_TOPK = 100
# (multipler, std_m) per error bound: Auncel/hyperparameter.txt lines 6 / 7 are the authors'
# SIFT10M k=10 settings for eb=0.1 / 0.05 (eval/run.sh:13-15); eb=0.2 reuses line 6.
HYPER = {0.1: (7.9, 6.0), 0.05: (10.2, 6.0), 0.2: (7.9, 6.0)}
# dram__bytes_read.sum + dram__bytes_write.sum per tc_filter launch of the default step under ncu
# (profiles/r02_tc_traffic.txt, refreshed whenever the kernel changes); None = not measured for this build
TC_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_tc_traffic.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nb", type=int, default=10_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--ncal", type=int, default=5_000)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--shape", default="sift")
    ap.add_argument("--eb", type=float, default=0.1)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the hbm_bound / shards sections")
    ap.add_argument("--shards-nprobe", type=int, default=64)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        mp = json.load(open(p))
        return dict(hbm=mp["hbm_gbs"], tf32=mp.get("bf16_tflops", 0) / 2.0 or None,
                    tf32_sustained=mp.get("bf16_tflops_sustained", 0) / 2.0 or None,
                    source="measured (MEASURED_PEAKS.json: hbm_gbs; tf32 = bf16_tflops / 2)")
    return dict(hbm=6650.0, tf32=None, tf32_sustained=None, source="fallback (B200_PROFILING.md)")


def build_everything(a, dev_index, rank, shape=None, nlist=None, calibrate=True):
    """Synthetic base + queries on the device, index build, exact ground truth, calibration."""
    import torch

    import auncel_b200 as ab
    from auncel_b200 import workload as W
    shape = shape or a.shape
    nlist = nlist or a.nlist
    dev = torch.device(f"cuda:{dev_index}")
    t0 = time.time()
    base = W.make_vectors(shape, a.nb, 123, dev)
    # queries: calibration set shared by all ranks, test set per rank (replicas serve different queries)
    qcal = W.make_vectors(shape, a.ncal, 456, dev)
    qtest = W.make_vectors(shape, a.nq, 789 + rank, dev)
    ix = W.build_index(ab, shape, base, nlist, dev_index, niter=10)
    q_all = torch.cat([qcal, qtest])
    gD, gI = W.ground_truth(ix, q_all, MAX_TOPK)
    es = ab.Error_sys(ix, a.ncal + a.nq, MAX_TOPK)
    gD_h = gD.cpu().numpy()
    es.set_gt(gD_h, gI.cpu().numpy())
    if calibrate:
        es.sys_train(a.ncal, qcal.cpu().numpy())
    torch.cuda.synchronize()
    return dict(ab=ab, W=W, dev=dev, base=base, qcal=qcal, qtest=qtest, ix=ix, gD=gD_h, es=es, shape=shape,
                setup_s=time.time() - t0)


# ----------------------------------------------------------------------------- small-batch regime
def hbm_bound_section(a, S, peaks, batches=(1, 64), calls=48):
    """Batch 1 (the reference's latency mode, eval/bound.cpp:390-396) and batch 64: every query streams
    its own lists, so algorithmic bytes = ndis * 4d are (nearly) the unique bytes and the HBM roofline
    applies.  Two fractions: bytes / scan-kernel time, and bytes / whole call (what a user sees)."""
    import torch
    ix, W, dev = S["ix"], S["W"], S["dev"]
    d = W.SHAPES[S["shape"]]["d"]
    out = {}
    for b in batches:
        ncalls = max(4, min(calls, a.nq // b))
        acc = torch.full((b,), 1.0 - a.eb, device=dev)
        npb = torch.zeros(b, device=dev, dtype=torch.int64)
        D = torch.empty(b, MAX_TOPK, device=dev)
        I = torch.empty(b, MAX_TOPK, device=dev, dtype=torch.int64)
        ix.set_params(*HYPER[a.eb])
        ms, scan, ndis, wall = [], [], [], []
        for c in range(ncalls + 3):
            q = S["qtest"][(c * b) % (a.nq - b + 1):][:b]
            npb.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ix.search_bounded_device(q, MAX_TOPK, QUERY_TOPK, acc, npb, D, I)
            w = time.perf_counter() - t0
            if c >= 3:
                st = ix.stats()
                ms.append(st["search_ms"])
                scan.append(st["scan_ms"])
                ndis.append(st["ndis"])
                wall.append(w * 1e3)
        bytes_call = float(np.mean(ndis)) * 4 * d
        out[str(b)] = {
            "calls": ncalls, "ms_per_call_device": float(np.mean(ms)), "ms_per_call_wall": float(np.mean(wall)),
            "scan_ms": float(np.mean(scan)), "qps": b / (float(np.mean(wall)) / 1e3),
            "alg_bytes_per_call": bytes_call,
            "scan_gbs": bytes_call / (float(np.mean(scan)) / 1e3) / 1e9,
            "scan_frac_of_hbm_peak": bytes_call / (float(np.mean(scan)) / 1e3) / 1e9 / peaks["hbm"],
            "call_gbs": bytes_call / (float(np.mean(ms)) / 1e3) / 1e9,
            "call_frac_of_hbm_peak": bytes_call / (float(np.mean(ms)) / 1e3) / 1e9 / peaks["hbm"]}
    return out


# ----------------------------------------------------------------------------- shards (config 3)
def shards_section(a, world, rank, local, steps, warmup):
    """BASELINE config 3: 10M x 96 DEEP-shaped inner product, k = 100, fixed nprobe; the database is
    split over the ranks (id % world, copy_subset_to type 1), queries go to every rank, one
    ncclAllGather + merge inside the library.  Strong scaling (total work fixed)."""
    import torch
    import torch.distributed as dist

    import auncel_b200 as ab
    from auncel_b200 import distributed as AD
    from auncel_b200 import workload as W
    dev = torch.device(f"cuda:{local}")
    shape, K, nprobe = "deep", 100, a.shards_nprobe
    d, metric = W.SHAPES[shape]["d"], W.SHAPES[shape]["metric"]
    t0 = time.time()
    base = W.make_vectors(shape, a.nb, 321, dev)       # every rank draws the same base ...
    q = W.make_vectors(shape, a.nq, 654, dev)           # ... and the same queries
    # shared quantizer: trained identically on every rank (same data, same seed, deterministic k-means)
    sh = ab.IndexIVFFlat(d, a.nlist, metric, device=local)
    gp = torch.Generator(device=dev)
    gp.manual_seed(5)
    sel = torch.randperm(a.nb, generator=gp, device=dev)[:min(a.nb, 64 * a.nlist)]
    sh.set_tune_mode()  # (the centroid-distance table interdis_cem: the error-bounded part below needs it)
    sh.train(base[sel].cpu().numpy(), niter=6)
    sh.set_tune_off()
    cent = sh.centroids()
    ids = torch.arange(rank, a.nb, world, device=dev)
    sh.add_device(base[ids].contiguous(), ids.cpu().numpy())
    sh.nprobe = nprobe
    g = AD.NcclShardGroup(sh)
    D = torch.empty(a.nq, K, device=dev)
    I = torch.empty(a.nq, K, device=dev, dtype=torch.int64)
    setup_s = time.time() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        g.search_device(q, K, D, I)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lms, ams, mms = [], [], []
    t_wall = time.perf_counter()
    for _ in range(steps):
        g.search_device(q, K, D, I)
        st = g.stats()
        lms.append(st["local_ms"])
        ams.append(st["allgather_ms"])
        mms.append(st["merge_ms"])
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3 / steps
    step_ms = float(np.mean(lms)) + float(np.mean(ams)) + float(np.mean(mms))  # device time on the index stream
    t = torch.tensor([step_ms, wall_ms, float(np.mean(lms)), float(np.mean(ams)), float(np.mean(mms))], device=dev,
                     dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, wall_ms, l_ms, a_ms, m_ms = [float(v) for v in t]
    out = {"workload": f"IVF-Flat nlist={a.nlist}, {a.nb}x{d} deep-shaped inner product, {a.nq} queries, k={K}, "
                       f"nprobe={nprobe}, lists split id % {world}",
           "parallelism": f"shards x{world}: 1 ncclAllGather of the packed (D|I) tables + merge per step",
           "scaling": "strong", "value": a.nq / (step_ms / 1e3), "unit": "queries/s", "ms_per_step": step_ms,
           "ms_per_step_wall": wall_ms, "local_search_ms": l_ms, "allgather_ms": a_ms, "merge_ms": m_ms,
           "allgather_bytes_per_step_per_rank": st["allgather_bytes"], "nccl_version": st["nccl_version"],
           "collective_share_of_step": (a_ms + m_ms) / step_ms, "setup_s": setup_s}
    # ---- error-bounded search over the same shards, single-index semantics (csrc/shard_rounds.cu):
    # the ranks exchange every round's candidates, so my_nprobe / distances are those of ONE index
    bounded = None
    if world > 1:
        ncal = min(a.ncal, 1000)
        qcal = W.make_vectors(shape, ncal, 987, dev)
        sh.nprobe = a.nlist
        gD = torch.empty(ncal, K, device=dev)
        gI = torch.empty(ncal, K, device=dev, dtype=torch.int64)
        g.search_device(qcal, K, gD, gI)                      # exhaustive over all shards: the ground truth
        sh.nprobe = nprobe
        g.set_bounded(True)
        qcal_h, gD_h = qcal.cpu().numpy(), gD.cpu().numpy()
        sh.calibrate(qcal_h, K, gD_h)                         # collective: the traces of the whole database
        sh.set_params(*HYPER[a.eb])
        acc_t = torch.full((a.nq,), 1.0 - a.eb, device=dev)
        np_t = torch.zeros(a.nq, device=dev, dtype=torch.int64)
        Db = torch.empty(a.nq, K, device=dev)
        Ib = torch.empty(a.nq, K, device=dev, dtype=torch.int64)
        bms = []
        for it in range(2 + steps):
            np_t.zero_()
            sh.search_bounded_device(q, K, QUERY_TOPK, acc_t, np_t, Db, Ib)
            if it >= 2:
                bms.append(sh.stats()["search_ms"])
        barrier()
        tb = torch.tensor([float(np.mean(bms))], device=dev, dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        xs, stb = g.exchange_stats(), sh.stats()
        bounded = {"workload": f"error bound {a.eb}, top-{QUERY_TOPK} of {K}, same shards and queries; single-index semantics",
                   "parallelism": f"shards x{world}: per round 1 ncclAllGather of the candidates each rank's scan produced, "
                                  "stage replay on every rank; 1 ncclAllReduce(max) of the labels",
                   "value": a.nq / (float(tb[0]) / 1e3), "unit": "queries/s", "ms_per_step": float(tb[0]),
                   "rounds": stb["rounds"], "tc_rounds": stb["tc_rounds"], "mean_my_nprobe": float(np_t.float().mean()),
                   "candidates_sent_per_step_this_rank": xs["entries_sent"], "candidates_all_ranks": xs["entries_all"],
                   "exchange_bytes_received_per_step_per_rank": xs["bytes_received"]}
        g.set_bounded(False)
    # parity: the merged result equals the unsharded index (tests/test_merge.cpp:94-152 invariant)
    if world > 1:
        nchk = 512
        if rank == 0:
            one = ab.IndexIVFFlat(d, a.nlist, metric, device=local)
            one.set_centroids(cent)
            one.add_device(base)
            one.nprobe = nprobe
            D1 = torch.empty(nchk, K, device=dev)
            I1 = torch.empty(nchk, K, device=dev, dtype=torch.int64)
            one.search_device(q[:nchk], K, D1, I1)
            out["parity_vs_single_index"] = {
                "queries": nchk, "distances_bit_equal": bool(torch.equal(D1, D[:nchk])),
                "labels_equal_frac": float((I1 == I[:nchk]).float().mean())}
            one.calibrate(qcal_h, K, gD_h)
            one.set_params(*HYPER[a.eb])
            np1 = torch.zeros(nchk, device=dev, dtype=torch.int64)
            one.search_bounded_device(q[:nchk], K, QUERY_TOPK, acc_t[:nchk].contiguous(), np1, D1, I1)
            bounded["parity_vs_single_index"] = {
                "queries": nchk, "my_nprobe_equal": bool(torch.equal(np1, np_t[:nchk])),
                "distances_bit_equal": bool(torch.equal(D1, Db[:nchk])),
                "labels_equal_frac": float((I1 == Ib[:nchk]).float().mean())}
            # the same bounded batch on ONE GPU holding the whole database (what the shards are compared with)
            one_ms = []
            for it in range(2 + steps):
                np_t.zero_()
                one.search_bounded_device(q, K, QUERY_TOPK, acc_t, np_t, Db, Ib)
                if it >= 2:
                    one_ms.append(one.stats()["search_ms"])
            bounded["single_gpu_ms_per_step"] = float(np.mean(one_ms))
            del one
        dist.barrier()
        out["error_bounded"] = bounded
    del g, sh, base
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- main arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    S = build_everything(a, local, rank)
    ix, W, dev = S["ix"], S["W"], S["dev"]
    n, d = a.nq, W.SHAPES[a.shape]["d"]
    metric = W.SHAPES[a.shape]["metric"]
    gt_test = S["gD"][a.ncal:]

    # device-resident buffers (value) and pinned host buffers (e2e)
    acc_t = torch.empty(n, device=dev, dtype=torch.float32)
    np_t = torch.zeros(n, device=dev, dtype=torch.int64)
    D_t = torch.empty(n, MAX_TOPK, device=dev, dtype=torch.float32)
    I_t = torch.empty(n, MAX_TOPK, device=dev, dtype=torch.int64)
    hx = S["qtest"].cpu().pin_memory()
    hacc = torch.empty(n, dtype=torch.float32).pin_memory()
    hnp = torch.zeros(n, dtype=torch.int64).pin_memory()
    hD = torch.empty(n, MAX_TOPK, dtype=torch.float32).pin_memory()
    hI = torch.empty(n, MAX_TOPK, dtype=torch.int64).pin_memory()
    from auncel_b200._lib import lib
    from auncel_b200.index import _ck
    L = lib()

    def step_device(eb):
        acc_t.fill_(1.0 - eb)
        np_t.zero_()
        ix.set_params(*HYPER[eb])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ix.search_bounded_device(S["qtest"], MAX_TOPK, QUERY_TOPK, acc_t, np_t, D_t, I_t)
        wall = time.perf_counter() - t0
        return ix.stats(), wall

    import ctypes as C
    L.auncel_index_search_bounded.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]

    def step_host(eb):
        hacc.fill_(1.0 - eb)
        hnp.zero_()
        ix.set_params(*HYPER[eb])
        t0 = time.perf_counter()
        _ck(L.auncel_index_search_bounded(ix.h, n, hx.data_ptr(), MAX_TOPK, QUERY_TOPK, hacc.data_ptr(), None,
                                          hnp.data_ptr(), None, 0, hD.data_ptr(), hI.data_ptr()))
        return time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step_device(a.eb)
        step_host(a.eb)

    # ---- timed: device-resident (value).  The library times each call with CUDA events on its
    # own stream (stats.search_ms); the K-step region is bracketed by barrier + synchronize.
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    dev_ms, scan_ms, ndis, launches, scan_launches = [], [], [], 0, 0
    keys = ("tc_ms", "tc_ndis", "simt_ms", "simt_ndis", "tc_rounds", "tc_uniq", "tc_staged", "simt_uniq", "simt_staged",
            "rounds", "coarse_ms")
    tcs = {k: [] for k in keys}
    t_region = time.perf_counter()
    per_round = []
    for _ in range(a.steps):
        st, _ = step_device(a.eb)
        per_round.append(ix.round_stats())
        dev_ms.append(st["search_ms"])
        scan_ms.append(st["scan_ms"])
        ndis.append(st["ndis"])
        launches += int(st["launches"])
        scan_launches += int(st["scan_launches"])
        for k in tcs:
            tcs[k].append(st[k])
    barrier()
    region_s = time.perf_counter() - t_region
    # ---- timed: host buffers (e2e)
    barrier()
    e2e_s = [step_host(a.eb) for _ in range(a.steps)]
    barrier()
    clocks = sampler.stop()
    D_eb = hD.numpy().copy()
    np_eb = hnp.numpy().copy()

    ms_step = float(np.mean(dev_ms))
    e2e_step = float(np.mean(e2e_s))
    per_rank_ms = [ms_step]
    # ---- quality on this rank's queries
    rec = W.recall_at(gt_test, D_eb, QUERY_TOPK, metric)
    quality = {"mean_recall@10": float(rec.mean()), "min_recall@10": float(rec.min()),
               "satisfied_frac": float((rec >= 1.0 - a.eb - 1e-6).mean()),
               "mean_my_nprobe": float(np_eb.mean()), "max_my_nprobe": int(np_eb.max())}
    if world > 1:
        mine = torch.tensor([ms_step, quality["satisfied_frac"], quality["mean_recall@10"]], device=dev, dtype=torch.float64)
        allr = torch.empty(world, 3, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine[None])
        per_rank_ms = [float(v) for v in allr[:, 0].cpu()]
        quality["satisfied_frac_min_over_ranks"] = float(allr[:, 1].min())
        quality["mean_recall@10_min_over_ranks"] = float(allr[:, 2].min())
        t = torch.tensor([ms_step, e2e_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step = float(t[0]), float(t[1])
    total_q = n * world

    other, curve = {}, {}
    if rank == 0:
        for eb in (0.05, 0.2):
            if abs(eb - a.eb) < 1e-9:
                continue
            step_device(eb)            # (first call at another bound grows the scratch pools: untimed)
            st, _ = step_device(eb)
            r = W.recall_at(gt_test, D_t.cpu().numpy(), QUERY_TOPK, metric)
            other[str(eb)] = {"qps": n / (st["search_ms"] / 1e3), "mean_recall@10": float(r.mean()),
                              "satisfied_frac": float((r >= 1.0 - eb - 1e-6).mean()),
                              "mean_my_nprobe": float(np_t.cpu().numpy().mean())}
        # how hard the synthetic workload is: fixed-nprobe recall@10 (cf. benchs/README.md:229-241 of the reference)
        m = min(n, 2000)
        for npb in (1, 4, 16, 64, 256):
            ix.nprobe = npb
            ix.search_device(S["qtest"][:m], MAX_TOPK, D_t[:m], I_t[:m])
            curve[str(npb)] = float(W.recall_at(gt_test[:m], D_t[:m].cpu().numpy(), QUERY_TOPK, metric).mean())

    # ---- roofline of the dominant kernel, per launch.  Unit of work = one (query, vector) distance
    # evaluation: 4d algorithmic bytes, 2d TF32 flops in tc_filter_kernel (survivors recomputed exactly).
    # A list-batched kernel stages a list tile once for up to 256 queries, so the HBM-level figure is the
    # UNIQUE bytes of the lists a launch touches (counted by the plan kernel), SURVEY 8d "B_uniq".
    peaks = load_peaks()
    tc_ms, tc_ndis = float(np.mean(tcs["tc_ms"])), float(np.mean(tcs["tc_ndis"]))
    simt_ms, simt_ndis = float(np.mean(tcs["simt_ms"])), float(np.mean(tcs["simt_ndis"]))
    tc_launches = max(float(np.mean(tcs["tc_rounds"])), 1.0)
    uniq_bytes = float(np.mean(tcs["tc_uniq"])) * 4 * d
    staged_bytes = float(np.mean(tcs["tc_staged"])) * 4 * d
    flops = tc_ndis * 2 * d
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6 / 1e12  # FP32 lane-ops/s at the measured clock (no FMA on the exact path)
    flop_per_dis = 3 * d if metric == 1 else 2 * d
    hbm_gbs = uniq_bytes / (tc_ms / 1e3) / 1e9 if tc_ms > 0 else None
    tf32_tflops = flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else None
    hbm_frac = hbm_gbs / peaks["hbm"] if hbm_gbs else None
    tensor_frac = tf32_tflops / peaks["tf32"] if (tf32_tflops and peaks["tf32"]) else None
    intensity = flops / uniq_bytes if uniq_bytes else None
    ridge = peaks["tf32"] * 1e12 / (peaks["hbm"] * 1e9) if peaks["tf32"] else None
    bound = "hbm" if (intensity is None or ridge is None or intensity < ridge) else "tensor"
    # per launch: every tensor-core round is one tc_filter_kernel launch; its active bound is the resource it
    # drives closest to the peak (HBM on unique bytes, or the tensor pipe on TF32 flops)
    launches_rf = []
    nr = min(len(r) for r in per_round) if per_round else 0
    for ri in range(nr):
        rows = [r[ri] for r in per_round]
        if not all(x["tc"] for x in rows):
            continue
        ms_i = float(np.mean([x["tc_ms"] for x in rows]))
        ub = float(np.mean([x["uniq"] for x in rows])) * 4 * d
        fl = float(np.mean([x["ndis"] for x in rows])) * 2 * d
        hf = ub / (ms_i / 1e3) / 1e9 / peaks["hbm"]
        tf = fl / (ms_i / 1e3) / 1e12 / peaks["tf32"] if peaks["tf32"] else 0.0
        launches_rf.append({"ranks": [int(rows[0]["r0"]), int(rows[0]["r0"] + rows[0]["w"])], "active_queries": int(rows[0]["active"]),
                            "ms": ms_i, "unique_bytes": ub, "tf32_flops": fl, "hbm_gbs": ub / (ms_i / 1e3) / 1e9,
                            "hbm_frac": hf, "tf32_tflops": fl / (ms_i / 1e3) / 1e12, "tensor_frac": tf,
                            "bound": "hbm" if hf >= tf else "tensor", "frac": max(hf, tf)})
    tw = sum(x["ms"] for x in launches_rf)
    frac_active = sum(x["ms"] * x["frac"] for x in launches_rf) / tw if tw > 0 else None
    longest = max(launches_rf, key=lambda x: x["ms"]) if launches_rf else None
    traffic = None
    if os.path.exists(TC_TRAFFIC_FILE) and a.nb == 10_000_000 and d == 128 and a.nlist == 4096:
        try:
            traffic = json.load(open(TC_TRAFFIC_FILE)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    arena_bytes = a.nb * d * 4
    all_flops = float(np.mean(ndis)) * 2 * d
    floor_ms = max(arena_bytes / (peaks["hbm"] * 1e9), all_flops / (peaks["tf32"] * 1e12) if peaks["tf32"] else 0.0) * 1e3
    if longest is not None:
        bound = longest["bound"]
    roofline = {
        "kernel": "tc_filter_kernel", "bound": bound, "peak_source": peaks["source"],
        "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
        "achieved": (longest["hbm_gbs"] if bound == "hbm" else longest["tf32_tflops"]) if longest else None,
        "peak": peaks["hbm"] if bound == "hbm" else peaks["tf32"],
        "frac": frac_active,
        "frac_definition": "time-weighted over the tc_filter launches of a step: each launch's fraction of ITS active "
                           "bound (unique list bytes / launch time vs the measured HBM copy peak, or TF32 flops / launch "
                           "time vs the measured tensor peak, whichever is higher); bound/achieved/peak describe the "
                           "longest launch; `launches` lists all of them; `hbm`/`tensor` are the aggregates over all launches",
        "launches": launches_rf,
        "traffic": traffic,
        "per_launch": {"unique_bytes": uniq_bytes / tc_launches, "staged_bytes": staged_bytes / tc_launches,
                       "tf32_flops": flops / tc_launches, "ms": tc_ms / tc_launches, "launches_per_step": tc_launches},
        "arithmetic_intensity_flop_per_byte": intensity, "ridge_flop_per_byte": ridge,
        "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks["hbm"], "frac": hbm_frac,
                "staged_gbs": staged_bytes / (tc_ms / 1e3) / 1e9 if tc_ms > 0 else None},
        "tensor": {"achieved_tflops_tf32": tf32_tflops, "peak_tflops_tf32": peaks["tf32"], "frac": tensor_frac},
        "reuse": {"note": "algorithmic bytes (ndis*4d, one list pass per probing query, as the reference streams "
                          "them) over kernel time; NOT a bandwidth: one staged tile serves up to 256 queries",
                  "algorithmic_gbs": tc_ndis * 4 * d / (tc_ms / 1e3) / 1e9 if tc_ms > 0 else None,
                  "queries_per_staged_byte": tc_ndis * 4 * d / staged_bytes if staged_bytes else None},
        "exact_scan": {"kernel": "scan_kernel", "ms_per_step": simt_ms, "ndis": simt_ndis,
                       "unique_gbs": float(np.mean(tcs["simt_uniq"])) * 4 * d / (simt_ms / 1e3) / 1e9 if simt_ms > 0 else None,
                       "fp32_pipe_frac": simt_ndis * flop_per_dis / (simt_ms / 1e3) / 1e12 / fp32_peak if simt_ms > 0 else None},
        "step": {"ms_per_step": ms_step, "floor_ms": floor_ms, "frac_of_floor": floor_ms / ms_step,
                 "note": "floor = max(one pass over the list arena at the HBM peak, all distance evaluations of the "
                         "step as TF32 flops at the tensor peak)",
                 "tc_filter_share": tc_ms / ms_step, "exact_scan_share": simt_ms / ms_step,
                 "coarse_share": float(np.mean(tcs["coarse_ms"])) / ms_step,
                 "rounds": float(np.mean(tcs["rounds"]))},
    }

    line = {
        "metric": "QPS at fixed error bound & recall@10, 10M x 128 SIFT-shape", "value": total_q / (ms_step / 1e3),
        "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"IVF-Flat nlist={a.nlist}, {a.nb}x{d} {a.shape}-shaped, {n} queries/GPU, "
                               f"max_topk={MAX_TOPK}, query_topk={QUERY_TOPK}, error bound {a.eb}, "
                               f"(multipler,std_m)={HYPER[a.eb]}, calibration {a.ncal} queries",
                   "parallelism": "replicas x%d (query split, no collective)" % world,
                   "l2": "inputs larger than L2: every step streams the %.1f GB list arena" % (a.nb * d * 4 / 1e9)},
        "e2e": {"value": total_q / e2e_step, "unit": "queries/s", "h2d_bytes_per_step": n * d * 4 + n * 4 + n * 8,
                "d2h_bytes_per_step": n * MAX_TOPK * 12 + n * 8, "ms_per_step": 1e3 * e2e_step},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "quality": quality,
        "error_bounds": other, "fixed_nprobe_recall@10": curve, "setup_s": S["setup_s"], "timed_region_s": region_s,
        "per_rank_ms_per_step": per_rank_ms,
    }
    if rank == 0 and world == 1 and not a.no_cpu:
        line["cpu_baseline"] = cpu_baseline(a, S, np_eb, D_eb)
    if not a.no_extras:
        if rank == 0:
            hb = {"sift": hbm_bound_section(a, S, peaks)}
        # free the headline index before the other shapes are built
        del ix
        S.clear()
        torch.cuda.empty_cache()
        if rank == 0 and world == 1 and a.shape == "sift":
            St = build_everything(a, local, rank, shape="text")
            hb["text"] = hbm_bound_section(a, St, peaks)
            hb["text"]["setup_s"] = St["setup_s"]
            St.clear()
            torch.cuda.empty_cache()
        if rank == 0:
            line["hbm_bound"] = hb
        sh = shards_section(a, world, rank, local, max(3, a.steps), max(3, a.warmup))
        if rank == 0:
            line["shards"] = sh
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- reference arm
def build_reference(a, S):
    """The unmodified reference with the same index content: same centroids, the same list
    assignment (precomputed_idx, IndexIVFFlat.cpp:41-59), same ground truth, same traces."""
    from oracle import oracle as O
    ix, W = S["ix"], S["W"]
    d, metric = W.SHAPES[a.shape]["d"], W.SHAPES[a.shape]["metric"]
    O.RefIndex.set_blas_threshold(1 << 30)  # exact-difference coarse path, like the GPU side
    R = O.RefIndex(d, a.nlist, metric)
    R.set_centroids(ix.centroids())
    base = S["base"]
    bs = 1 << 20
    for i0 in range(0, a.nb, bs):
        xb = base[i0:i0 + bs].cpu().numpy()
        R.add(xb, ids=np.arange(i0, i0 + len(xb), dtype=np.int64), list_no=ix.assign(xb))
    return R, O


SETUP_NOTE = ("setup (synthetic data, k-means, list assignment, ground truth, calibration traces) ran through "
              "auncel_b200 on the GPU and was handed to the reference (centroids, precomputed list numbers, traces); "
              "the timed region is the unmodified reference alone")


def cpu_sample_search(a, S, R, O, nsample, threads, chunk=4):
    """Error_sys::search of the reference over `nsample` test queries on `threads` host threads."""
    ix = S["ix"]
    ncal = 10  # Error_sys needs a trained error_pro; its traces are then replaced by the full
    # calibration's (bit-identical to the reference's own sys_train, tests/test_gpu_golden.py)
    q = np.concatenate([S["qcal"][:ncal].cpu().numpy(), S["qtest"][:nsample].cpu().numpy()])
    gD = np.concatenate([S["gD"][:ncal], S["gD"][a.ncal:a.ncal + nsample]])
    gI = np.zeros_like(gD, dtype=np.int64)
    R.es_create(gD, gI)
    R.sys_train(ncal, q)
    R.set_traces(ix.traces())
    acc = np.full(ncal + nsample, 1.0 - a.eb, np.float32)
    R.set_queries(QUERY_TOPK, nsample, q, acc, *HYPER[a.eb])
    t0 = time.perf_counter()
    D, I = R.es_search(ncal, nsample, threads=threads, chunk=chunk)  # dynamic hand-out, `chunk` queries at a time
    dt = time.perf_counter() - t0
    return dt, D, R.my_nprobe(ncal, nsample)


def cpu_sample_size(a, cores):
    nround = max(10, (a.nq // 10) * 10)
    return a.cpu_sample or min(nround, max(1000, (32 * cores) // 10 * 10))


def cpu_baseline(a, S, np_gpu, D_gpu):
    from oracle import oracle as O
    if not O.have_ref():
        return {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/libauncel_ref.so missing"}
    cores = os.cpu_count() or 1
    nsample = cpu_sample_size(a, cores)
    R, O = build_reference(a, S)
    dt, D, mynp = cpu_sample_search(a, S, R, O, nsample, cores)
    # the reference's latency mode: one query per call, one thread (eval/bound.cpp:390-396)
    nlat = min(32, nsample)
    R.clear_my_nprobe()
    lat = []
    for i in range(nlat):
        t0 = time.perf_counter()
        R.es_search(10 + i, 1, search_size=1, threads=1)
        lat.append((time.perf_counter() - t0) * 1e3)
    R.close()
    return {"value": nsample / dt, "unit": "queries/s", "cores": cores, "kind": "reference",
            "latency_mode": {"queries": nlat, "ms_per_query_mean": float(np.mean(lat)), "ms_per_query_p50": float(np.median(lat)),
                             "note": "one Error_sys::search call per query on one thread (eval/bound.cpp:390-396); compare "
                                     "hbm_bound.sift['1'].ms_per_call_wall"},
            "sample": f"first {nsample} test queries, one batched Error_sys::search on {cores} host threads drawing 4 "
                      f"queries at a time (unmodified reference, exact-difference coarse path, OpenBLAS unused), {dt:.2f} s",
            "setup": SETUP_NOTE,
            "parity_on_sample": {"my_nprobe_equal": bool(np.array_equal(mynp.astype(np.int64), np_gpu[:nsample])),
                                 "distances_bit_equal": bool(np.array_equal(D, D_gpu[:nsample])),
                                 "my_nprobe_mismatches": [[int(i), int(mynp[i]), int(np_gpu[i])] for i in
                                                          np.flatnonzero(mynp.astype(np.int64) != np_gpu[:nsample])[:8]]}}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    if not O.have_ref():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libauncel_ref.so not present"})
        return
    import torch
    torch.cuda.set_device(0)
    S = build_everything(a, 0, 0)  # data generation + ground truth + calibration tables only
    W = S["W"]
    d = W.SHAPES[a.shape]["d"]
    cores = os.cpu_count() or 1
    nsample = cpu_sample_size(a, cores)
    R, O = build_reference(a, S)
    times = []
    for i in range(a.warmup + a.steps):
        R.clear_my_nprobe() if i else None
        dt, D, mynp = cpu_sample_search(a, S, R, O, nsample, cores)
        if i >= a.warmup:
            times.append(dt)
    R.close()
    t = float(np.mean(times))
    v = nsample / t
    emit({
        "impl": "reference", "metric": "QPS at fixed error bound & recall@10, 10M x 128 SIFT-shape", "value": v,
        "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"IVF-Flat nlist={a.nlist}, {a.nb}x{d} {a.shape}-shaped, sample of {nsample} queries, "
                               f"max_topk={MAX_TOPK}, query_topk={QUERY_TOPK}, error bound {a.eb}, "
                               f"(multipler,std_m)={HYPER[a.eb]}"},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "reference",
                         "sample": f"{nsample} test queries per step, batched Error_sys::search on {cores} threads "
                                   f"drawing 4 queries at a time", "setup": SETUP_NOTE},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


if __name__ == "__main__":
    args = parse()
    isolate_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
