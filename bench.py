#!/usr/bin/env python
"""bench.py -- QPS of the error-bounded IVF-Flat query path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (configs[1] of BASELINE.json): IVF-Flat nlist=4096, 10M x 128 synthetic SIFT-shaped
base, 10k test queries (+5k calibration queries), result heap max_topk=100, query_topk=10,
error bound 0.1 (targets 0.05 / 0.2 are timed once each and reported under "error_bounds").
One step = one Error_sys::search over the whole query batch.
  value  : queries/s with the queries already resident in HBM (device C-ABI entry point)
  e2e    : queries/s through the host-pointer C-ABI call (pinned host buffers; H2D of the
           queries/targets and D2H of distances/labels/my_nprobe inside the timed region)
N > 1 (torchrun, one rank per GPU): replicas -- IndexReplicas semantics
(Auncel/IndexReplicas.cpp:79-118): every rank holds the index and serves its own query batch,
no data-path collective; value = all ranks' queries / max-over-ranks time ("weak").
--impl reference: the unmodified reference (oracle/_ref, all host threads) on a bounded
sample of the same workload, rank 0 only.
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
# dram__bytes_read.sum + dram__bytes_write.sum per tc_filter launch, mean of the four launches of one
# default step under ncu (profiles/r01_tc_traffic.txt: 5.48 + 7.26 + 5.73 + 3.16 GB in 4.42 ms)
TC_TRAFFIC_PER_LAUNCH = 5.40e9
MAX_TOPK = 100
# (multipler, std_m) per error bound: Auncel/hyperparameter.txt lines 6 / 7 are the authors'
# SIFT10M k=10 settings for eb=0.1 / 0.05 (eval/run.sh:13-15); eb=0.2 reuses line 6.
HYPER = {0.1: (7.9, 6.0), 0.05: (10.2, 6.0), 0.2: (7.9, 6.0)}


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
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_everything(a, dev_index, rank):
    """Synthetic base + queries on the device, index build, exact ground truth, calibration."""
    import torch

    import auncel_b200 as ab
    from auncel_b200 import workload as W
    dev = torch.device(f"cuda:{dev_index}")
    t0 = time.time()
    base = W.make_vectors(a.shape, a.nb, 123, dev)
    # queries: calibration set shared by all ranks, test set per rank (replicas serve different queries)
    qcal = W.make_vectors(a.shape, a.ncal, 456, dev)
    qtest = W.make_vectors(a.shape, a.nq, 789 + rank, dev)
    ix = W.build_index(ab, a.shape, base, a.nlist, dev_index, niter=10)
    q_all = torch.cat([qcal, qtest])
    gD, gI = W.ground_truth(ix, q_all, MAX_TOPK)
    es = ab.Error_sys(ix, a.ncal + a.nq, MAX_TOPK)
    gD_h = gD.cpu().numpy()
    es.set_gt(gD_h, gI.cpu().numpy())
    es.sys_train(a.ncal, qcal.cpu().numpy())
    torch.cuda.synchronize()
    return dict(ab=ab, W=W, dev=dev, base=base, qcal=qcal, qtest=qtest, ix=ix, gD=gD_h, es=es,
                setup_s=time.time() - t0)


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

    def step_host(eb):
        hacc.fill_(1.0 - eb)
        hnp.zero_()
        ix.set_params(*HYPER[eb])
        t0 = time.perf_counter()
        _ck(L.auncel_index_search_bounded(ix.h, n, hx.data_ptr(), MAX_TOPK, QUERY_TOPK, hacc.data_ptr(), None,
                                          hnp.data_ptr(), None, 0, hD.data_ptr(), hI.data_ptr()))
        return time.perf_counter() - t0

    import ctypes as C
    L.auncel_index_search_bounded.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]

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
    tcs = {k: [] for k in ("tc_ms", "tc_ndis", "simt_ms", "simt_ndis", "tc_rounds", "tc_uniq", "tc_staged", "simt_uniq",
                           "simt_staged")}
    t_region = time.perf_counter()
    for _ in range(a.steps):
        st, _ = step_device(a.eb)
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
    if world > 1:
        mine = torch.tensor([ms_step], device=dev, dtype=torch.float64)
        allr = torch.empty(world, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        per_rank_ms = [float(v) for v in allr.cpu()]
        t = torch.tensor([ms_step, e2e_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step = float(t[0]), float(t[1])
    total_q = n * world

    # ---- quality on this rank's queries
    rec = W.recall_at(gt_test, D_eb, QUERY_TOPK, metric)
    quality = {"mean_recall@10": float(rec.mean()), "min_recall@10": float(rec.min()),
               "satisfied_frac": float((rec >= 1.0 - a.eb - 1e-6).mean()),
               "mean_my_nprobe": float(np_eb.mean()), "max_my_nprobe": int(np_eb.max())}
    other = {}
    if rank == 0:
        for eb in (0.05, 0.2):
            if abs(eb - a.eb) < 1e-9:
                continue
            st, _ = step_device(eb)
            r = W.recall_at(gt_test, D_t.cpu().numpy(), QUERY_TOPK, metric)
            other[str(eb)] = {"qps": n / (st["search_ms"] / 1e3), "mean_recall@10": float(r.mean()),
                              "satisfied_frac": float((r >= 1.0 - eb - 1e-6).mean()),
                              "mean_my_nprobe": float(np_t.cpu().numpy().mean())}

    # ---- roofline of the dominant kernel.  Unit of work = one (query, vector) distance evaluation,
    # 4d algorithmic bytes / 2d (IP) or 3d (L2) flops each.  The bulk of the work runs
    # in tc_filter_kernel (TF32 tcgen05 filter; survivors recomputed exactly), the rest in the
    # exact FP32 scan_kernel.  Both are timed inside the library with CUDA events on its stream.
    peak, peak_src = load_peaks()
    tc_ms, tc_ndis = float(np.mean(tcs["tc_ms"])), float(np.mean(tcs["tc_ndis"]))
    simt_ms, simt_ndis = float(np.mean(tcs["simt_ms"])), float(np.mean(tcs["simt_ndis"]))
    tc_launches = float(np.mean(tcs["tc_rounds"]))
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6 / 1e12  # FP32 lane-ops/s at the measured clock (no FMA on the exact path)
    flop_per_dis = 3 * d if metric == 1 else 2 * d
    # DRAM bytes per tensor-core launch: measured with ncu on this very command (dram__bytes_read+write of
    # the four tc_filter launches of a step, profiles/r01_scan_tc_ncu_v3.txt / r01_tc_traffic.txt)
    tf32_peak = None
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        tf32_peak = mp.get("bf16_tflops", 0) / 2.0 or None  # tf32 runs at half the measured dense bf16 rate
    except Exception:
        pass
    tc_tflops = tc_ndis * 2 * d / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else None
    roofline = {
        "kernel": "tc_filter_kernel", "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
        "achieved": tc_ndis * 4 * d / (tc_ms / 1e3) / 1e9 if tc_ms > 0 else None,
        "frac": tc_ndis * 4 * d / (tc_ms / 1e3) / 1e9 / peak if tc_ms > 0 else None,
        "traffic": TC_TRAFFIC_PER_LAUNCH if (a.nb == 10_000_000 and d == 128 and a.nlist == 4096) else None,
        "per_launch": {"alg_bytes": tc_ndis * 4 * d / max(tc_launches, 1), "ms": tc_ms / max(tc_launches, 1),
                       "launches_per_step": tc_launches},
        "dram": {"note": "compulsory traffic = bytes of the distinct lists each launch touches (counted by the plan "
                         "kernel); staged = bytes TMA moved into shared memory (one list pass per 256-query tile). "
                         "ncu per launch (profiles/r01_scan_tc_ncu_v3.txt): launches with ~1 query tile per list run "
                         "at 6.76 TB/s DRAM (103 % of the measured copy peak); launches with 3-4 query tiles per "
                         "list are bound by the tf32 MMA rate (tensor pipe 60 % active, 3.7 TB/s DRAM, 1.9x the "
                         "arena because L2 keeps only part of a list between its query tiles)",
                 "compulsory_bytes_per_step": float(np.mean(tcs["tc_uniq"])) * 4 * d,
                 "staged_bytes_per_step": float(np.mean(tcs["tc_staged"])) * 4 * d,
                 "compulsory_gbs": float(np.mean(tcs["tc_uniq"])) * 4 * d / (tc_ms / 1e3) / 1e9 if tc_ms > 0 else None,
                 "staged_gbs": float(np.mean(tcs["tc_staged"])) * 4 * d / (tc_ms / 1e3) / 1e9 if tc_ms > 0 else None,
                 "compulsory_frac_of_peak": float(np.mean(tcs["tc_uniq"])) * 4 * d / (tc_ms / 1e3) / 1e9 / peak if tc_ms > 0 else None},
        "tensor": {"achieved_tflops_tf32": tc_tflops, "peak_tflops_tf32": tf32_peak,
                   "frac": tc_tflops / tf32_peak if (tc_tflops and tf32_peak) else None,
                   "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2"},
        "note": "algorithmic bytes = ndis*4d as the reference streams them (one list pass per probing query); one "
                "staged tile serves up to 256 queries, so achieved/peak > 1 is reuse -- the HBM-level figure is "
                "`dram` (each launch reads the arena about once)",
        "exact_scan": {"kernel": "scan_kernel", "ms_per_step": simt_ms, "ndis": simt_ndis,
                       "achieved_gbs": simt_ndis * 4 * d / (simt_ms / 1e3) / 1e9 if simt_ms > 0 else None,
                       "compulsory_gbs": float(np.mean(tcs["simt_uniq"])) * 4 * d / (simt_ms / 1e3) / 1e9 if simt_ms > 0 else None,
                       "fp32_pipe_frac": simt_ndis * flop_per_dis / (simt_ms / 1e3) / 1e12 / fp32_peak if simt_ms > 0 else None},
        "scan_share_of_step": float(np.mean(scan_ms)) / ms_step,
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
        "error_bounds": other, "setup_s": S["setup_s"], "timed_region_s": region_s,
        "per_rank_ms_per_step": per_rank_ms,
    }
    if rank == 0 and world == 1 and not a.no_cpu:
        line["cpu_baseline"] = cpu_baseline(a, S, np_eb, D_eb)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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


def cpu_sample_search(a, S, R, O, nsample, threads):
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
    D, I = R.es_search(ncal, nsample, threads=threads)
    dt = time.perf_counter() - t0
    return dt, D, R.my_nprobe(ncal, nsample)


def cpu_baseline(a, S, np_gpu, D_gpu):
    from oracle import oracle as O
    if not O.have_ref():
        return {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/libauncel_ref.so missing"}
    cores = os.cpu_count() or 1
    nround = max(10, (a.nq // 10) * 10)
    nsample = a.cpu_sample or min(nround, max(200, (8 * cores) // 10 * 10))
    R, O = build_reference(a, S)
    dt, D, mynp = cpu_sample_search(a, S, R, O, nsample, cores)
    R.close()
    return {"value": nsample / dt, "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": f"first {nsample} test queries, one batched Error_sys::search on {cores} host threads "
                      f"(unmodified reference, exact-difference coarse path, OpenBLAS unused), {dt:.2f} s",
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
    nsample = a.cpu_sample or min(max(10, a.nq // 10 * 10), max(200, (8 * cores) // 10 * 10))
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
                         "sample": f"{nsample} test queries per step, batched Error_sys::search on {cores} threads"},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


if __name__ == "__main__":
    args = parse()
    isolate_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
