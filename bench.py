#!/usr/bin/env python
"""bench.py -- QPS of the IVF-PQ search hot path on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA kernels through the C-ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)
    torchrun ... bench.py --gpus N ...                       # N ranks, index sharded by vector

A "step" is one pass of the hot path (coarse -> LUT -> ADC scan -> top-k [-> all-gather -> merge]) over one
10 000-query batch.  Default workload at N = 1 is BASELINE.json configs[1]: 100M x 128, IVF8192,PQ16x8,
nprobe = 32, k = 10 (fits one B200: 1.6 GB codes + 0.8 GB ids).  With N > 1 the same 100M database is sharded
by vector over the N GPUs (strong scaling: fixed database and batch, value = queries / time).

Everything printed besides the JSON line goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "chameleon-rag-acceleration_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIGS = {
    # name: (nb, d, nlist, M, nprobe, k, nq)   -- BASELINE.json configs
    "c1": (1_000_000, 128, 1024, 16, 16, 10, 10_000),
    "c2": (100_000_000, 128, 8192, 16, 32, 10, 10_000),
    "c3": (1_000_000_000, 96, 65536, 16, 64, 100, 10_000),
    "c4": (100_000_000, 768, 16384, 64, 32, 10, 128),     # RALM: batch-1 latency is reported beside the 128-query step
    "c5": (100_000_000, 128, 8192, 32, 32, 10, 10_000),
}
METRIC = "QPS at recall@10 parity (10k-query batch) + p50 batch-1 latency, 1/2/4/8 B200"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, for one) write to fd 1, so the real
# stdout is kept aside for the JSON line and fd 1 is pointed at stderr for everything else.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def workload_name(cfg, args):
    nb, d, nlist, M, nprobe, k, nq = cfg
    return (f"{args.config}: {nb // 1_000_000}M x {d}, IVF{nlist},PQ{M}x8, nprobe={nprobe}, k={k}, "
            f"{nq}-query batch, synthetic clustered data")


# ---------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every few ms from a thread (the timed region
    of the default run is ~150 ms, too short for nvidia-smi's own loop); nvidia-smi -lms as the fallback."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll_nvml(self):
        nv, h = self.nvml
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, bit in bits.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                break
            self._stop.wait(0.004)

    def start(self):
        try:
            self.nvml = self._nvml_handle()
            nv, h = self.nvml
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self._stop.set()
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": self.mx[0] if self.mx else None, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------------
REF_SHRINK = {"c1": 1, "c2": 8, "c3": 64, "c4": 32, "c5": 8}   # the CPU arm's index is nb / shrink vectors in nlist / shrink lists


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads(oracle):
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: the CPU arm must not inherit that."""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    oracle.C.set_num_threads(n)
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    return oracle.C.num_threads()


def build_index_cpu(cfg, shrink, args):
    """The GPU arm's data (same ClusteredGenerator, same seeds, same 12-dimensional blobs), built with host code only
    (torch CPU GEMMs + numpy; none of this repo's kernels): nb / shrink vectors in nlist / shrink lists, so that a list is
    as long as in the full-size index and a query scans the same number of codes and builds the same number of look-up
    tables with the SAME nprobe -- the per-query work of the path -- while the build stays within a minute of host time.
    Only the coarse stage is smaller (nlist / shrink centroids)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
    from b200ivfpq.datasets import SEED_BASE, SEED_QUERY, SEED_TRAIN, ClusteredGenerator
    from b200ivfpq.kmeans import _assign, kmeans, kmeans_subspaces

    nb, d, nlist, M, nprobe, k, nq = cfg
    nb_r, nlist_r = max(nb // shrink, 1), max(nlist // shrink, nprobe)
    cpu = torch.device("cpu")
    gen = ClusteredGenerator(d, ncentres=max(16, nlist_r // 2), sigma=args.sigma, device=cpu, seed=7,
                             latent_dim=args.latent_dim, sigma_iso=args.sigma_iso)
    ntrain = min(nb_r, max(64 * nlist_r, 65536))
    xt = torch.cat([gen.chunk(SEED_TRAIN, i, min(1 << 18, ntrain - (i << 18))) for i in range((ntrain + (1 << 18) - 1) >> 18)])
    coarse = kmeans(xt, nlist_r, niter=8, seed=1234)
    lab, _ = _assign(xt[:65536], coarse)
    pq = kmeans_subspaces(xt[:65536] - coarse[lab], M, 256, niter=8, seed=4321)
    del xt
    dsub = d // M
    list_no = np.empty(nb_r, np.int32)
    codes = np.empty((nb_r, M), np.uint8)
    chunk = 1 << 18
    pn = (pq * pq).sum(2)                                              # (M, 256)
    for i0 in range(0, nb_r, chunk):
        x = gen.chunk(SEED_BASE, i0 // chunk, min(chunk, nb_r - i0))
        l, _ = _assign(x, coarse)
        r = (x - coarse[l]).reshape(-1, M, dsub).permute(1, 0, 2)       # (M, n, dsub)
        dist = torch.baddbmm(pn.unsqueeze(1), r, pq.transpose(1, 2), alpha=-2.0)
        codes[i0:i0 + x.shape[0]] = dist.argmin(2).t().to(torch.uint8).numpy()
        list_no[i0:i0 + x.shape[0]] = l.numpy()
    order = np.argsort(list_no, kind="stable")
    offsets = np.zeros(nlist_r + 1, np.int64)
    offsets[1:] = np.cumsum(np.bincount(list_no, minlength=nlist_r))
    xq = gen.chunk(SEED_QUERY, 0, nq).numpy()
    arrays = (np.ascontiguousarray(coarse.numpy()), np.ascontiguousarray(pq.numpy()), offsets,
              np.ascontiguousarray(codes[order]), order.astype(np.int64))
    return arrays, np.ascontiguousarray(xq), nb_r, nlist_r


def time_oracle(oracle, xq, arrays, nprobe, k, budget_s=15.0, max_q=None):
    """Bounded CPU sample: pilot on 64 queries, then as many queries as fit ~budget_s.  Returns (qps, nq_used, s)."""
    coarse, pq, offsets, codes, ids = arrays
    npilot = min(64, xq.shape[0])
    t0 = time.perf_counter()
    oracle.C.search(xq[:npilot], coarse, pq, offsets, codes, ids, nprobe, k)
    per_q = (time.perf_counter() - t0) / npilot
    n = int(max(npilot, min(xq.shape[0] if max_q is None else max_q, budget_s / max(per_q, 1e-9))))
    n = min(n, xq.shape[0])
    t0 = time.perf_counter()
    D, I = oracle.C.search(xq[:n], coarse, pq, offsets, codes, ids, nprobe, k)
    dt = time.perf_counter() - t0
    return n / dt, n, dt, D, I


def bench_config(cfg, args):
    """The `config` object of the JSON line: the workload and nothing run-dependent, identical in both arms."""
    nb, d, nlist, M, nprobe, k, nq = cfg
    return {"workload": workload_name(cfg, args), "nb": nb, "d": d, "index": f"IVF{nlist},PQ{M}x8", "nprobe": nprobe,
            "k": k, "batch": nq}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ivfpq_oracle as oracle
    oracle.build()
    cores = use_all_host_threads(oracle)
    cfg = apply_overrides(CONFIGS[args.config], args)
    nb, d, nlist, M, nprobe, k, nq = cfg
    shrink = REF_SHRINK[args.config] if not args.nb else max(1, nb // 12_500_000)
    t0 = time.perf_counter()
    log(f"[reference] {cores} host threads; building the clustered index on the CPU ({nb // shrink} vectors, "
        f"IVF{max(nlist // shrink, nprobe)}) ...")
    arrays, xq, nb_r, nlist_r = build_index_cpu(cfg, shrink, args)
    log(f"[reference] built in {time.perf_counter() - t0:.1f} s")
    # a step = the 10k-query batch, or the part of it that keeps the whole run within a few minutes
    npilot = min(64, nq)
    t0 = time.perf_counter()
    D0, I0 = oracle.C.search(xq[:npilot], *arrays, nprobe, k)
    per_q = (time.perf_counter() - t0) / npilot
    offsets = arrays[2]
    total = args.steps + args.warmup
    nq_step = int(max(1, min(nq, (150.0 / max(total, 1)) / max(per_q, 1e-9))))
    # scan bytes per query of this index (what the CPU arm streams per query), for comparison with the GPU arm's
    pl = oracle.C.search(xq[:npilot], *arrays, nprobe, k, return_probes=True)[3]
    sizes = np.diff(offsets)
    scan_mb = float(sizes[np.maximum(pl, 0)].sum() * M / npilot / 1e6)
    log(f"[reference] {per_q * 1e3:.2f} ms/query on {cores} threads -> {nq_step} queries per step; "
        f"{scan_mb:.2f} MB of codes scanned per query")
    for _ in range(args.warmup):
        oracle.C.search(xq[:nq_step], *arrays, nprobe, k)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle.C.search(xq[:nq_step], *arrays, nprobe, k)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    qps = nq_step / (ms / 1e3)
    sample = (f"{nq_step} of {nq} queries per step, {cores} OpenMP threads, same generator / seeds / nprobe / k as the GPU "
              f"arm on a 1/{shrink}-size index ({nb_r} vectors, IVF{nlist_r}: lists as long as the full index's, "
              f"{scan_mb:.2f} MB of codes and {nprobe} look-up tables per query; only the coarse stage is 1/{shrink}); "
              f"restated Faiss-CPU algorithm (oracle/ivfpq_oracle.c, residual look-up tables built per probe, no "
              f"precomputed-table shortcut), not the Faiss binary")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32/u8", "data": "synthetic",
        "config": bench_config(cfg, args), "queries_per_step": nq_step, "scan_mbytes_per_query": scan_mb,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def apply_overrides(cfg, args):
    if args.nb:
        cfg = (args.nb,) + cfg[1:]
    if args.nq:
        cfg = cfg[:6] + (args.nq,)
    if args.nprobe:
        cfg = cfg[:4] + (args.nprobe,) + cfg[5:]
    return cfg


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def build_index(cfg, rank, world, device, dist, args, layout=None):
    """Generate -> train -> assign/encode -> discard, chunk by chunk on the GPU.  Rank r
    keeps the 2M-vector chunks c with c % world == r.  Exact ground truth for a query sample is computed alongside."""
    import torch

    import b200ivfpq as faiss
    from b200ivfpq.datasets import SEED_BASE, SEED_QUERY, SEED_TRAIN, ClusteredGenerator

    nb, d, nlist, M, nprobe, k, nq = cfg
    gen = ClusteredGenerator(d, ncentres=args.ncentres or max(16, nlist // 2), sigma=args.sigma,
                             device=device, seed=7, latent_dim=args.latent_dim, sigma_iso=args.sigma_iso)
    index = faiss.index_factory(d, f"IVF{nlist},PQ{M}x8")
    index.cp_niter = args.kmeans_iters
    t0 = time.perf_counter()
    ntrain = min(nb, max(256000, 100 * nlist))            # bench_cpu_performance.py:77-90
    if rank == 0:
        xt = torch.cat([gen.chunk(SEED_TRAIN, i, min(1 << 20, ntrain - i * (1 << 20)))
                        for i in range((ntrain + (1 << 20) - 1) >> 20)])
        index.train(xt)
        del xt
        coarse, pq = index.quantizer.xb_tensor(), index.pq.centroids_tensor()
    else:
        coarse = torch.empty((nlist, d), dtype=torch.float32, device=device)
        pq = torch.empty((M, 256, d // M), dtype=torch.float32, device=device)
    if world > 1:
        dist.broadcast(coarse, 0)
        dist.broadcast(pq, 0)
        if rank != 0:
            index.set_codebooks(coarse, pq)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0

    xq = gen.chunk(SEED_QUERY, 0, max(nq, 1))
    ngt = min(args.gt_queries, xq.shape[0])
    gt_d = torch.full((ngt, 10), float("inf"), device=device)
    gt_i = torch.full((ngt, 10), -1, dtype=torch.int64, device=device)
    xq_gt = xq[:ngt]
    qn = (xq_gt * xq_gt).sum(1, keepdim=True)

    # Sharding by vector: chunk c (2M consecutive vectors) lives on rank c % world, so every inverted list is spread
    # over all GPUs and each rank only generates, encodes and ground-truths its own 1/world of the database.
    t0 = time.perf_counter()
    chunk = 1 << 21
    nchunks = (nb + chunk - 1) // chunk
    # With R replicas of S = world / R shards each (--replicas R; `--shard-mode replica` is R = world), a rank holds shard
    # `sh` of S: the chunks c % S == sh.  layout = (S, sh, shard_group); the ground truth is merged inside the shard group.
    S, sh, shard_group = layout or (world, rank, None)
    by_list = args.shard_mode == "list" and world > 1        # every rank encodes everything, then keeps its lists
    gS, gsh, ggroup = (world, rank, None) if by_list else (S, sh, shard_group)      # who ground-truths which chunk
    for ci in (range(nchunks) if by_list else range(sh, nchunks, S)):
        n = min(chunk, nb - ci * chunk)
        x = gen.chunk(SEED_BASE, ci, n)
        pos0 = ci * chunk
        if ngt and ci % gS == gsh:
            # exact brute force for recall: ||x||^2 - 2 q.x (+ ||q||^2), fp32 library GEMM (not on the timed path)
            dd = torch.addmm((x * x).sum(1).unsqueeze(0), xq_gt, x.t(), alpha=-2.0) + qn
            cd, cidx = torch.topk(dd, 10, dim=1, largest=False)
            alld = torch.cat([gt_d, cd], 1)
            alli = torch.cat([gt_i, cidx + pos0], 1)
            gt_d, sel = torch.topk(alld, 10, dim=1, largest=False)
            gt_i = torch.gather(alli, 1, sel)
            del dd
        index.add_with_ids(x, torch.arange(pos0, pos0 + n, device=device))
        del x
        if rank == 0 and ((ci // S) % 10 == 0 or ci + S >= nchunks):
            torch.cuda.synchronize()
            log(f"[build] chunk {ci + 1}/{nchunks}  {time.perf_counter() - t0:.1f} s")
    if gS > 1 and ngt:
        # merge the per-rank ground truth
        all_d = torch.empty((gS * ngt, 10), device=device)
        all_i = torch.empty((gS * ngt, 10), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(all_d, gt_d.contiguous(), group=ggroup)
        dist.all_gather_into_tensor(all_i, gt_i.contiguous(), group=ggroup)
        all_d = all_d.view(gS, ngt, 10).permute(1, 0, 2).reshape(ngt, gS * 10)
        all_i = all_i.view(gS, ngt, 10).permute(1, 0, 2).reshape(ngt, gS * 10)
        gt_d, sel = torch.topk(all_d, 10, dim=1, largest=False)
        gt_i = torch.gather(all_i, 1, sel)
    if by_list:
        # every rank encoded everything; keep the lists this rank owns (l % world == rank)
        index = faiss.shard_index_by_list(index, rank, world)
        index.nprobe = nprobe
    index._sync_lists()
    torch.cuda.synchronize()
    t_add = time.perf_counter() - t0
    index.nprobe = nprobe
    return index, xq[:nq].contiguous(), gt_i, {"train_s": t_train, "add_s": t_add}


def compare_results(Dg, Ig, Dr, Ir):
    """GPU result vs oracle: distance bits, ids, and ids up to the order inside runs of equal distance (what a merge of
    shards may legitimately change: BASELINE exempts ties)."""
    bits_equal = bool(np.array_equal(Dg.view(np.uint32), Dr.view(np.uint32)))
    ids_equal = bool(np.array_equal(Ig, Ir))
    ties_ok, bad = True, 0
    if bits_equal and not ids_equal:
        k = Dg.shape[1]
        for q in np.nonzero((Ig != Ir).any(1))[0]:
            d = Dr[q].view(np.uint32)
            i = 0
            while i < k:
                j = i
                while j + 1 < k and d[j + 1] == d[i]:
                    j += 1
                if sorted(Ig[q, i:j + 1].tolist()) != sorted(Ir[q, i:j + 1].tolist()) and j != k - 1:
                    ties_ok = False
                    bad += 1
                    break
                i = j + 1
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = float(np.nanmax(np.abs(Dg - Dr) / np.maximum(np.abs(Dr), 1e-30))) if Dg.size else 0.0
    return {"distances_bit_exact": bits_equal, "ids_identical": ids_equal,
            "ids_identical_modulo_ties": bool(bits_equal and ties_ok), "queries_differing_outside_ties": int(bad),
            "max_rel_dist_err": rel}


def gather_probed_lists(index, xq_s, nprobe, rank, world, dist, device):
    """The oracle's view of the WHOLE (unsharded) index restricted to the lists the sample queries probe: every rank
    sends the entries of those lists to rank 0, which concatenates them list by list in rank order (= the order a single
    index holding shard 0's entries first would have).  Unprobed lists are left empty, which no sample query can see."""
    import torch
    nlist = index.nlist
    _, probes = index.quantizer.search(xq_s, min(nprobe, nlist))      # bit-identical to the oracle's coarse stage (tested)
    need = torch.zeros(nlist, dtype=torch.bool, device=device)
    need[probes.flatten().clamp(min=0)] = True
    index._finalize_lists()
    off = torch.from_numpy(index._offsets).to(device)
    lists = torch.nonzero(need).flatten()
    lens = (off[1:] - off[:-1])[lists]
    tot = int(lens.sum())
    cum = torch.cumsum(lens, 0) - lens
    rows = torch.repeat_interleave(off[lists], lens) + (torch.arange(tot, device=device) - torch.repeat_interleave(cum, lens))
    codes, ids = index._codes[rows].contiguous(), index._ids[rows].contiguous()
    list_no = torch.repeat_interleave(lists, lens)
    del rows
    if world > 1:
        if rank == 0:
            parts = [(list_no, codes, ids)]
            for r in range(1, world):
                n = torch.zeros(1, dtype=torch.int64, device=device)
                dist.recv(n, src=r)
                ln = torch.empty(int(n), dtype=torch.int64, device=device)
                cd = torch.empty((int(n), codes.shape[1]), dtype=torch.uint8, device=device)
                ii = torch.empty(int(n), dtype=torch.int64, device=device)
                dist.recv(ln, src=r)
                dist.recv(cd, src=r)
                dist.recv(ii, src=r)
                parts.append((ln, cd, ii))
            list_no = torch.cat([p[0] for p in parts])
            codes = torch.cat([p[1] for p in parts])
            ids = torch.cat([p[2] for p in parts])
            del parts
        else:
            dist.send(torch.tensor([list_no.shape[0]], dtype=torch.int64, device=device), dst=0)
            dist.send(list_no.contiguous(), dst=0)
            dist.send(codes, dst=0)
            dist.send(ids, dst=0)
            return None
    _, order = torch.sort(list_no, stable=True)
    counts = torch.bincount(list_no, minlength=nlist).cpu().numpy()
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(counts)
    return (index.quantizer.xb_tensor().cpu().numpy(), index.pq.centroids_tensor().cpu().numpy(), offsets,
            codes[order].cpu().numpy(), ids[order].cpu().numpy())


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    import b200ivfpq as faiss
    from b200ivfpq.shards import DistributedIndexIVFPQ

    cfg = apply_overrides(CONFIGS[args.config], args)
    nb, d, nlist, M, nprobe, k, nq = cfg
    if rank == 0:
        log(f"[bench] {workload_name(cfg, args)} on {world} GPU(s)")
    # multi-GPU layout: R replicas x S shards by vector (default R = 1: the reference's co.shard = True over all GPUs)
    R = world if args.shard_mode == "replica" else max(1, args.replicas)
    if world % R or (R > 1 and args.shard_mode == "list"):
        raise SystemExit(f"--replicas {R} must divide the number of GPUs ({world}) and needs --shard-mode vector")
    layout = None
    if world > 1 and R == world:
        layout = (1, 0, None)                                 # every rank holds (and builds) the whole index
    elif world > 1 and R > 1:
        from b200ivfpq.shards import IndexReplicas, make_replica_groups, replica_layout
        S, rep, sh = replica_layout(world, rank, R)
        shard_group, cross_group = make_replica_groups(R)
        layout = (S, sh, shard_group)
    index, xq, gt, build_info = build_index(cfg, rank, world, device, dist, args, layout)
    if world == 1:
        searcher = index
    elif R == 1 or R == world:
        searcher = DistributedIndexIVFPQ(index, shard_mode="replica" if R == world else args.shard_mode)
    else:
        searcher = IndexReplicas(DistributedIndexIVFPQ(index, group=shard_group), R, rep, cross_group)
    index.set_stage_timing(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return searcher.search(xq, k)

    # ---- parity + recall (outside the timed region) -----------------------------------------------------
    D, I = step_device()
    torch.cuda.synchronize()
    stats = index.last_scan_stats()
    fstats = index.filter_stats(reset=True) if os.environ.get("B200_IVFPQ_QL_STATS") == "1" else None
    if fstats and rank == 0:
        log(f"[bench] filter: {fstats} for {stats['codes']} (query, code) pairs")
    recall = None
    if rank == 0 and gt.shape[0]:
        ng = gt.shape[0]
        Ic, gc = I[:ng, :10].cpu().numpy(), gt.cpu().numpy()
        recall = float(sum(np.intersect1d(a, b).shape[0] for a, b in zip(Ic, gc)) / gc.size)
        log(f"[bench] recall@10 on {ng} queries = {recall:.4f}; scan bytes/query = {stats['bytes'] / nq / 1e6:.2f} MB")

    # ---- timed region: queries resident in HBM, results left in HBM ------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = faiss.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms, stage_acc = [], {}
    barrier()
    profile_range = os.environ.get("B200_BENCH_PROFILE") == "1"   # ncu --profile-from-start off: timed steps only
    if profile_range:
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    if profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = faiss.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    qps = nq / (ms_per_step / 1e3)

    # per-stage device times (CUDA events recorded by the library on the launching stream); a few extra steps
    filter_ms = []
    for _ in range(min(args.steps, 5)):
        step_device()
        st = index.stage_ms()
        filter_ms.append(index.filter_ms())
        scan_ms.append(st["scan"])
        for kk, v in st.items():
            stage_acc.setdefault(kk, []).append(v)
    stages = {kk: float(np.mean(v)) for kk, v in stage_acc.items()}
    scan_t = float(np.mean(scan_ms))
    # every rank's own scan time and scanned bytes (rank 0's line reports them: shard balance)
    per_rank = torch.tensor([scan_t, float(stats["bytes"]), ms_total / args.steps], dtype=torch.float64, device=device)
    if world > 1:
        allr = torch.empty((world, 3), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allr, per_rank)
    else:
        allr = per_rank.view(1, 3)
    allr = allr.cpu().numpy()

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region ------------------
    xq_host = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
    xq_host.copy_(xq)
    xq_np = xq_host.numpy()

    D_host = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
    I_host = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)

    def step_host():
        if world == 1:
            return index.search(xq_np, k)                  # b200_ivfpq_search_host: H2D, kernels, D2H, sync
        # N > 1: pinned host buffers in, pinned host buffers out, one synchronisation at the end
        Dd, Id = searcher.search(xq_host.to(device, non_blocking=True), k)
        D_host.copy_(Dd, non_blocking=True)
        I_host.copy_(Id, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return D_host, I_host

    for _ in range(max(1, min(args.warmup, 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_t = torch.tensor([(time.perf_counter() - t0) * 1e3], device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t.item()) / args.steps
    e2e_qps = nq / (e2e_ms / 1e3)

    # ---- batch-1 latency (p50 over 200 single-query searches, host in / host out) ---------------------
    lat = []
    q1 = xq_np[:1].copy()
    for i in range(220):
        qi = xq_np[i % nq:i % nq + 1]
        t0 = time.perf_counter()
        if world == 1:
            index.search(qi, k)
        else:
            Dd, Id = searcher.search(torch.from_numpy(qi).to(device), k)
            Dd.cpu()
        if i >= 20:
            lat.append((time.perf_counter() - t0) * 1e3)
    lat_p50 = float(np.median(lat))
    del q1
    # ---- the same latency while a decode-like GEMM stream shares the GPU (BASELINE config 4: retrieval interleaved with
    # LLM decode, ralm_tiktok.py:129-192): a second CUDA stream keeps bf16 GEMMs of a decoder-layer shape in flight
    lat_decode_p50 = None
    if args.decode_interleave and world == 1:
        side = torch.cuda.Stream(device=device)
        wa = torch.randn((8, 8192), device=device, dtype=torch.bfloat16)        # 8 sequences x hidden 8192
        wb = torch.randn((8192, 28672), device=device, dtype=torch.bfloat16)    # an MLP up-projection
        stop = threading.Event()

        def decode_loop():
            torch.cuda.set_device(local)
            with torch.cuda.stream(side):
                while not stop.is_set():
                    for _ in range(16):
                        torch.matmul(wa, wb)
                    side.synchronize()

        th = threading.Thread(target=decode_loop, daemon=True)
        th.start()
        time.sleep(0.05)
        lat2 = []
        for i in range(220):
            qi = xq_np[i % nq:i % nq + 1]
            t0 = time.perf_counter()
            index.search(qi, k)
            if i >= 20:
                lat2.append((time.perf_counter() - t0) * 1e3)
        stop.set()
        th.join(timeout=5)
        lat_decode_p50 = float(np.median(lat2))

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    filt = [v for v in filter_ms if v is not None]
    if filt:
        kern_name, kern_ms = "st_filter_kernel (integer lower-bound filter over the probed lists)", float(np.mean(filt))
    else:
        kern_name, kern_ms = "scan stage (table set-up + ADC scan + top-k)", scan_t
    achieved = stats["bytes"] / (kern_ms / 1e3) / 1e9
    # DRAM traffic of that kernel: one `ncu --set full` capture of the same command, kept under profiles/ with the
    # kernel name, config and GPU count it belongs to (never reused for another kernel or shape)
    traffic, traffic_src, pipe_pct = None, None, None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        try:
            for tj in json.load(open(tp)).get("captures", []):
                if (tj.get("config") == args.config and tj.get("n_gpus", 1) == world and not args.nb and
                        tj.get("kernel", "").split("<")[0] in kern_name):
                    traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
                    pipe_pct = tj.get("l1_data_pipe_pct")
        except Exception:
            pass
    dram_frac = (traffic / (kern_ms / 1e3) / 1e9 / peak) if traffic else None
    hbm_bound = dram_frac is not None and dram_frac > 0.5
    roofline = {"bound": "hbm" if hbm_bound else "l1/shared-memory data pipe (codes are served from L2: each probed list "
                         "is scanned by the ~nq*nprobe/nlist queries that probe it; DRAM traffic << algorithmic bytes)",
                "kernel": kern_name, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": stats["bytes"], "launch_ms": kern_ms,
                "hbm_equivalent_frac": achieved / peak, "dram_frac": dram_frac, "l1_data_pipe_pct": pipe_pct,
                "note": "achieved = algorithmic bytes (sum over probed lists of list_size*M) / kernel time: an equivalent "
                        "streaming rate, it exceeds the DRAM rate by the L2 reuse factor and can exceed 1.0 of the copy "
                        "peak; dram_frac = measured DRAM bytes (ncu capture) / kernel time / peak; l1_data_pipe_pct = "
                        "ncu l1tex__data_pipe_lsu_wavefronts, the pipe that bounds the kernel"}

    # ---- CPU baseline + parity (any N): the oracle on the SAME index and queries -----------------------
    cpu_baseline, parity = None, None
    # layouts whose shards are disjoint parts of ONE index (the default by-vector split, whole lists per GPU) are
    # gathered from all ranks; a replicated index is checked on rank 0's copy; mixed R x S layouts are not checked
    gather_world = world if (R == 1 or world == 1) else 1 if R == world else 0
    if not args.no_cpu_baseline and gather_world and (gather_world > 1 or rank == 0):
        ns = min(nq, args.parity_queries or (256 if nb * M > 8e9 else nq))
        arrays = gather_probed_lists(index, xq[:ns], nprobe, rank, gather_world, dist, device)
        if rank == 0:
            from oracle import ivfpq_oracle as oracle
            oracle.build()
            cores = use_all_host_threads(oracle)
            cqps, n_used, dt, Dr, Ir = time_oracle(oracle, xq_np[:ns], arrays, nprobe, k, budget_s=args.cpu_budget_s)
            Dg, Ig = D[:n_used].cpu().numpy(), I[:n_used].cpu().numpy()
            parity = {"queries": int(n_used), **compare_results(Dg, Ig, Dr, Ir),
                      "against": f"oracle on the unsharded index (entries of all {world} shards, lists probed by the sample)",
                      "oracle": "CPU restatement (oracle/), itself pinned against the reference's own code run as a C "
                                "simulation (oracle/_ref: hnswlib cell selection + accelerator kernel); not the Faiss binary"}
            if gt.shape[0]:
                ng = min(gt.shape[0], n_used)
                gc = gt[:ng].cpu().numpy()
                parity["recall_at_10_oracle"] = float(
                    sum(np.intersect1d(x_, y_).shape[0] for x_, y_ in zip(Ir[:ng, :10], gc)) / gc.size)
                parity["recall_at_10_ours_same_queries"] = float(
                    sum(np.intersect1d(x_, y_).shape[0] for x_, y_ in zip(Ig[:ng, :10], gc)) / gc.size)
            cpu_baseline = {"value": cqps, "unit": "queries/s", "cores": cores, "kind": "port",
                            "sample": f"first {n_used} of {nq} queries, same index (all {world} shard(s)) and queries as the "
                                      f"GPU run, {dt:.1f} s; restated Faiss-CPU algorithm (oracle, OpenMP; residual look-up "
                                      f"tables per probe, no precomputed-table shortcut), not the Faiss binary"}
            log(f"[bench] cpu baseline {cqps:.1f} q/s on {cores} threads; parity {parity}")
            del arrays

    if rank == 0:
        sharding = (f"by list (NOT the reference's split): list l on GPU l % {world}, probes of other GPUs' lists masked"
                    if args.shard_mode == "list" and world > 1 else
                    f"replicated (Faiss IndexReplicas, NOT the sharded config): full index on each of "
                    f"{world} GPUs, batch sliced by query, results all-gathered"
                    if R == world and world > 1 else
                    f"{R} replicas (batch sliced by query) x {world // R} shards by vector (the reference's "
                    f"-R {R}); NOT the all-GPU sharded config" if R > 1 else
                    f"by vector: 2M-vector chunks round-robin over {world} GPU(s)")
        shard_merge = ("none" if world == 1 else
                       "none: NCCL all-gather of the per-slice results" if R == world else
                       "K5 reads every shard's top-k in place over NVLink (symmetric memory)"
                       if getattr(searcher, "peer_merge", False) else
                       "NCCL all-gather + K5" + (f" (peer memory unavailable: {searcher.peer_merge_error})"
                                                 if getattr(searcher, "peer_merge_error", None) else ""))
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32/u8", "data": "synthetic",
            "config": bench_config(cfg, args),
            "details": {"sharding": sharding, "shard_merge": shard_merge,
                        "l2_policy": "inputs larger than L2 (codes %.0f MB per GPU vs 126 MB L2)" % (index.ntotal * M / 1e6),
                        "scan_kernel": os.environ.get("B200_IVFPQ_SCAN", "auto"), "ntotal_per_gpu": index.ntotal,
                        "build": build_info},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_qps, "unit": "queries/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": nq * d * 4,
                    "d2h_bytes_per_step": nq * k * 12},
            "latency_batch1_ms_p50": lat_p50, "latency_batch1_with_decode_stream_ms_p50": lat_decode_p50,
            "recall_at_10": recall, "stages_ms": stages,
            "scan_ms_per_rank": [round(float(v), 3) for v in allr[:, 0]],
            "scan_gbytes_per_rank": [round(float(v) / 1e9, 2) for v in allr[:, 1]],
            "ms_per_step_per_rank": [round(float(v), 3) for v in allr[:, 2]],
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_vs_oracle": parity, "filter_stats": fstats,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--nb", type=int, default=0, help="override database size (exploration only)")
    ap.add_argument("--nq", type=int, default=0)
    ap.add_argument("--nprobe", type=int, default=0)
    ap.add_argument("--sigma", type=float, default=0.1)
    ap.add_argument("--sigma-iso", type=float, default=0.002)
    ap.add_argument("--latent-dim", type=int, default=12)
    ap.add_argument("--ncentres", type=int, default=0)
    ap.add_argument("--gt-queries", type=int, default=1000)
    ap.add_argument("--kmeans-iters", type=int, default=25, help="Lloyd iterations for index.train (Faiss default 25)")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--parity-queries", type=int, default=0,
                    help="queries (from the start of the batch) checked against the oracle and timed on the CPU; 0 = the whole "
                         "batch, or 256 when the index is too large to copy to the host (C3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-interleave", action="store_true",
                    help="also measure the batch-1 latency while a second stream runs decoder-shaped bf16 GEMMs (config c4)")
    ap.add_argument("--replicas", type=int, default=1,
                    help="N > 1: R replica groups of N/R vector shards each, the batch sliced by query between the "
                         "groups (the reference's -R, bench_gpu_performance_OSDI.py:613-626); default 1 = all GPUs shard")
    ap.add_argument("--shard-mode", default="vector", choices=["vector", "list", "replica"],
                    help="N > 1: 'vector' = the reference's / north_star's split (every list on every GPU); 'list' = "
                         "whole lists per GPU (l %% world); 'replica' = Faiss IndexReplicas (co.shard = False): full index "
                         "on every GPU, the batch sliced by query")
    return ap.parse_args(argv)


def main():
    args = parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: timing rules ask for >= 3 warm-up steps")
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
