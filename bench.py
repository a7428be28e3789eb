#!/usr/bin/env python
"""bench.py — batched exact-match k-mer COUNT throughput (BASELINE.json configs[1]), plus a located-hits/s leg
(configs[2] shape) and an opt-in derived-structures leg reported as extra objects on the same JSON line:
3.1 Gbp synthetic nucleotide index (seed k=12, SA ratio 8), 100 M random 20-mers per GPU, query-sharded.

    python bench.py --gpus N --steps K --warmup W            our arm  (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N ...             the reference's own OpenMP/AVX2 path on the host cores

One "step" = one pass of awFmParallelSearchCount's work over this rank's whole query batch.
  value : queries/s, whole job, inputs (packed queries + index) already resident in HBM, CUDA-event timed;
          for N>1 every step also gathers the per-rank count arrays onto rank 0 over NCCL/NVLink, overlapped
          with the search kernels chunk by chunk (the only collective on this path).
  e2e   : queries/s through the reference-facing drop-in call awFmParallelSearchCount(index, searchList, threads)
          on HOST memory: the 32-B AwFmKmerSearchData entries point at host strings; packing, H2D, kernels, D2H
          and the scatter of `count` back into the structs are all inside the timed region.
The index is built on the device by avxwindowfmindex_b200.build_index (byte-identical to awFmCreateIndex, see
tests/test_gpu_build_index.py) because the reference's CPU build of 3.1 Gbp takes ~20 min; the reference arm and
the cpu_baseline leg search that same index with the UNMODIFIED reference library (oracle/_ref/libawfm_ref.so).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kmer_count_queries_per_sec"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bp", type=int, default=3_100_000_000, help="text length (config: 3.1 Gbp)")
    ap.add_argument("--queries", type=int, default=100_000_000, help="queries per GPU (config: 100 M)")
    ap.add_argument("--kmer", type=int, default=20)
    ap.add_argument("--seed-k", type=int, default=12)
    ap.add_argument("--sa-ratio", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=10_000_000, help="queries per CPU-baseline pass")
    ap.add_argument("--count-path", default="auto", choices=["auto", "sweep", "tile"],
                    help="auto = the library's own choice (sweep for batches this large), tile = force the tile kernel")
    ap.add_argument("--locate-queries", type=int, default=10_000_000, help="cfg 3 leg: random 16-mers located per GPU")
    ap.add_argument("--locate-kmer", type=int, default=16)
    ap.add_argument("--derived-seed-depth", type=int, default=16, help="0 = skip the derived-structures leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region through NVML (the same counters the
    nvidia-smi line of B200_PROFILING.md prints), every 5 ms from a side thread."""

    def __init__(self, device):
        self.device = device
        self.rows = []
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device]) if visible and visible.split(",")[device].isdigit() else device
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _run(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                self.rows.append((n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM),
                                  n.nvmlDeviceGetPowerUsage(self._h) / 1000.0,
                                  n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)))
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self._nvml:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        n = self._nvml
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[2]
        names = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(self.sm_max),
                "power_w_max": max(r[1] for r in self.rows), "samples": len(self.rows),
                "reasons": [k for k, v in names.items() if bits & v]}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic_per_launch(path_name):
    """dram bytes of one count call over the bench batch from the committed ncu capture (profiles/), or None:
    "tile" = the single countKernelV1 launch, "sweep" = sum over the kernels of the sweep pipeline."""
    path = os.path.join(ROOT, "profiles", "ncu_count_traffic.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return d.get("dram_bytes_per_launch" if path_name == "tile" else "sweep_dram_bytes_per_call")
        except Exception:
            return None
    return None


def host_index_struct(arrays):
    """struct AwFmIndex over host arrays, as a C caller of the reference API holds it."""
    return arrays.as_awfm_index()


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_pass(ref, index_ptr, letters, n, length, threads, reps):
    """Times the unmodified reference's awFmParallelSearchCount on `n` host queries; returns best queries/s."""
    from avxwindowfmindex_b200 import KmerSearchList
    sl = KmerSearchList(ref.lib, n).fill(letters[: n * length], fixed_len=length)
    best = 0.0
    ref.lib.awFmParallelSearchCount(index_ptr, sl.ptr, threads)  # warm-up (page in index, spin up OpenMP)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ref.lib.awFmParallelSearchCount(index_ptr, sl.ptr, threads)
        times.append(time.perf_counter() - t0)
        best = max(best, n / times[-1])
    counts = sl.counts()
    sl.close()
    return best, times, counts


def build_on_device(args, device, lib):
    import torch
    from avxwindowfmindex_b200 import DeviceBuiltIndex, abi, capi, synth
    d_text = torch.empty(args.bp, dtype=torch.uint8, device=f"cuda:{device}")
    capi.check(lib.awfm_gpu_synth_letters(device, d_text.data_ptr(), args.bp, synth.TEXT_SEED + 2, 0, 0))
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), args.bp, abi.AwFmAlphabetDna, args.seed_k,
                                              args.sa_ratio, device=device)
    del d_text
    torch.cuda.synchronize()
    return built


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the reference arm
    import torch
    from avxwindowfmindex_b200 import capi, synth
    from oracle import harness
    lib = capi.load()
    threads = os.cpu_count()
    config = workload_config(args, 1)
    if not harness.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libawfm_ref.so was not built"}))
        return
    ref = harness.Reference()
    built = build_on_device(args, 0, lib)  # setup only: byte-identical to awFmCreateIndex, never timed
    arrays = built.to_host()
    built.close()
    torch.cuda.empty_cache()
    ix = host_index_struct(arrays)
    n = min(args.cpu_sample, args.queries)
    letters = synth.random_queries(n, args.kmer, seed=synth.QUERY_SEED + 2)
    sl_times = []
    from avxwindowfmindex_b200 import KmerSearchList
    sl = KmerSearchList(ref.lib, n).fill(letters, fixed_len=args.kmer)
    for _ in range(args.warmup):
        ref.lib.awFmParallelSearchCount(C.addressof(ix), sl.ptr, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        ref.lib.awFmParallelSearchCount(C.addressof(ix), sl.ptr, threads)
        sl_times.append(time.perf_counter() - t1)
    total = time.perf_counter() - t0
    sl.close()
    value = n * args.steps / total
    sample = f"{n} of the {args.queries} random {args.kmer}-mers per step, reference awFmParallelSearchCount, numThreads={threads}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    return {
        "workload": f"count: {args.bp} bp synthetic nucleotide index (seed k={args.seed_k}, SA ratio {args.sa_ratio}), "
                    f"{args.queries} random {args.kmer}-mers per GPU (BASELINE.json configs[1])",
        "text_bp": args.bp, "seed_k": args.seed_k, "sa_ratio": args.sa_ratio, "kmer": args.kmer,
        "queries_per_gpu": args.queries, "parallelism": f"query-sharded x{world}, index replicated per GPU" + ("; counts gathered to rank 0 over NCCL, the gather of step s overlapped with the search of step s+1" if world > 1 else ""),
        "l2_policy": "inputs larger than L2 (index 3.5 GB + packed queries 2 GB per step vs 126 MB L2)",
        "index_built_by": "device builder, byte-identical to the reference's awFmCreateIndex (tests/test_gpu_build_index.py)",
    }


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from avxwindowfmindex_b200 import KmerSearchList, abi, capi, synth
    from oracle import harness  # checker + cpu_baseline only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lib = capi.load()  # raises if the CUDA library is missing: there is no fallback
    dev = torch.device(f"cuda:{local}")

    # ---- index: built on this GPU, stays resident ----
    t0 = time.time()
    built = build_on_device(args, local, lib)
    gpu = built.gpu_index()
    build_s = time.time() - t0
    need_host = (not args.no_e2e) or rank == 0
    arrays = built.to_host() if need_host else None
    tie_suffixes, build_ms = built.tie_suffixes, built.build_ms
    built.close()
    torch.cuda.empty_cache()

    # ---- queries: this rank's shard of the random k-mer stream, resident in HBM ----
    n, L = args.queries, args.kmer
    d_letters = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(local, d_letters.data_ptr(), n * L, synth.QUERY_SEED + 2, rank * n * L, 0))
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    if args.count_path != "auto":
        gpu.set_tuning(sweep_min_queries=1 if args.count_path == "sweep" else -1)
    # N > 1: every step's counts are gathered onto rank 0 over NCCL/NVLink (the only collective on this path).  The
    # gather of step s runs on NCCL's stream while step s+1 searches into the other count buffer; the last gather is
    # drained inside the timed region.
    bufs = [d_counts, torch.zeros(n, dtype=torch.int32, device=dev)] if world > 1 else [d_counts]
    works = [None] * len(bufs)
    gathered = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step(s):
        b = s % len(bufs)
        if works[b] is not None:  # the gather that last read this buffer (stream dependency, the host does not block)
            works[b].wait()
            works[b] = None
        gpu.count_device(d_letters.data_ptr(), None, L, n, bufs[b].data_ptr(), None, stream.cuda_stream)
        if world > 1 and not os.environ.get("AWFM_BENCH_SKIP_GATHER"):  # (development switch: search only)
            works[b] = dist.gather(bufs[b], gathered if rank == 0 else None, dst=0, async_op=True)

    def drain():
        for b, w in enumerate(works):
            if w is not None:
                w.wait()
                works[b] = None

    for s in range(args.warmup):
        step(s)
    drain()
    torch.cuda.synchronize()
    launches_per_step = int(gpu.stats()["launches"])  # kernels of ours in one count call (same for every step)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local) as clocks:
        ev[0].record(stream)
        for s in range(args.steps):
            step(s)
        drain()
        ev[1].record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    launches = launches_per_step * args.steps
    total_ms = ev[0].elapsed_time(ev[1])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # device time of one count call over the whole batch for the roofline, CUDA events on the launching stream
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def count_call_ms(reps=5):
        out = []
        for _ in range(reps):
            ka.record(stream)
            gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream.cuda_stream)
            kb.record(stream)
            torch.cuda.synchronize()
            out.append(ka.elapsed_time(kb))
        return sum(out) / len(out)

    kernel_avg = count_call_ms()
    gpu.set_tuning(sweep_profile=1)
    gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream.cuda_stream)
    torch.cuda.synchronize()
    stage_ms = gpu.sweep_stage_ms()  # [] when the call took the tile kernel
    gpu.set_tuning(sweep_profile=0)
    tile_ms = None
    if stage_ms and args.count_path == "auto":  # the single-kernel path on the same batch, for the record
        gpu.set_tuning(sweep_min_queries=-1)
        gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream.cuda_stream)
        tile_ms = count_call_ms(reps=3)
        gpu.set_tuning(sweep_min_queries=0)
        gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream.cuda_stream)
        torch.cuda.synchronize()

    # ---- second half of BASELINE's metric: located hits/s (configs[2] shape: random 16-mers, this index's SA ratio),
    #      device-resident, ranges -> scan -> expand -> backtrace walk -> positions; outside the timed count steps ----
    result = {}
    nl, Ll = args.locate_queries, args.locate_kmer
    if nl > 0:
        d_lq = torch.empty(nl * Ll + 64, dtype=torch.uint8, device=dev)
        capi.check(lib.awfm_gpu_synth_letters(local, d_lq.data_ptr(), nl * Ll, synth.QUERY_SEED + 3, rank * nl * Ll, 0))
        d_lc = torch.zeros(nl, dtype=torch.int32, device=dev)
        d_lr = torch.zeros((nl, 2), dtype=torch.int64, device=dev)
        d_lh = torch.zeros(nl + 1, dtype=torch.int64, device=dev)
        gpu.count_device(d_lq.data_ptr(), None, Ll, nl, d_lc.data_ptr(), d_lr.data_ptr(), stream.cuda_stream)
        gpu.scan_ranges_device(d_lr.data_ptr(), nl, d_lh.data_ptr(), stream.cuda_stream)
        hits = int(d_lh[-1].item())
        d_lp = torch.zeros(max(hits, 1), dtype=torch.int64, device=dev)

        def locate_all():
            gpu.count_device(d_lq.data_ptr(), None, Ll, nl, d_lc.data_ptr(), d_lr.data_ptr(), stream.cuda_stream)
            gpu.scan_ranges_device(d_lr.data_ptr(), nl, d_lh.data_ptr(), stream.cuda_stream)
            gpu.locate_device(d_lr.data_ptr(), d_lh.data_ptr(), nl, 0, hits, d_lp.data_ptr(), stream.cuda_stream)

        def best_ms(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(reps):
                ka.record(stream)
                fn()
                kb.record(stream)
                torch.cuda.synchronize()
                best = min(best, ka.elapsed_time(kb))
            return best

        ms_all = best_ms(locate_all)
        ms_walk = best_ms(lambda: gpu.locate_device(d_lr.data_ptr(), d_lh.data_ptr(), nl, 0, hits, d_lp.data_ptr(),
                                                    stream.cuda_stream))
        loc = {"workload": f"{nl} random {Ll}-mers per GPU on the same index (SA ratio {args.sa_ratio}), BASELINE.json configs[2] shape",
               "hits": hits, "locate_ms": ms_all, "located_hits_per_s": hits / ms_all * 1e3,
               "locate_queries_per_s": nl / ms_all * 1e3, "walk_ms": ms_walk, "walk_hits_per_s": hits / ms_walk * 1e3}
        if arrays is not None:
            ls = min(nl, 200_000)
            o_hit, o_pos, lwork = harness.Oracle(arrays).locate(d_lq[: ls * Ll].cpu().numpy(), fixed_len=Ll,
                                                                 threads=os.cpu_count())
            nh = int(o_hit[-1])
            same = (np.array_equal(d_lh[: ls + 1].cpu().numpy().astype(np.uint64), o_hit)
                    and np.array_equal(d_lp[:nh].cpu().numpy().astype(np.uint64), o_pos))
            loc["parity_sample"] = {"queries": ls, "hits": nh, "bit_exact_vs_oracle": bool(same)}
            if nh:
                loc["backtrace_steps_per_hit"] = lwork["backtraceSteps"] / nh
                loc["algorithmic_bytes_per_hit"] = lwork["locateBytes"] / nh
                loc["walk_algorithmic_GBps"] = lwork["locateBytes"] / nh * hits / ms_walk / 1e6
            if not same:
                raise SystemExit("PARITY FAILURE: CUDA positions differ from the oracle on the bench's locate leg")
        result["locate"] = loc
        h_lq_cpu = d_lq[: min(nl, 1_000_000) * Ll].cpu().numpy()
        h_lq = None
        if not args.no_e2e:
            h_lq = torch.empty(nl * Ll, dtype=torch.uint8).pin_memory()
            h_lq.copy_(d_lq[: nl * Ll])

        # ---- derived structures (opt-in: HBM for fewer dependent DRAM round trips), same queries, same checks ----
        if args.derived_seed_depth > args.seed_k:
            try:
                derived = {}
                build_seed_ms = gpu.extend_seed_table(args.derived_seed_depth)
                dms = best_ms(lambda: gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None,
                                                       stream.cuda_stream), reps=3)
                derived["seed_table"] = {"depth": args.derived_seed_depth, "build_ms": build_seed_ms,
                                         "count_ms": dms, "count_queries_per_s": n / dms * 1e3}
                build_sa_ms = gpu.densify_suffix_array(1)
                wms = best_ms(lambda: gpu.locate_device(d_lr.data_ptr(), d_lh.data_ptr(), nl, 0, hits, d_lp.data_ptr(),
                                                        stream.cuda_stream))
                derived["suffix_array"] = {"sa_ratio": 1, "build_ms": build_sa_ms, "walk_ms": wms,
                                           "walk_hits_per_s": hits / wms * 1e3}
                derived["device_bytes_with_both"] = gpu.device_bytes()
                if arrays is not None:
                    derived["suffix_array"]["bit_exact_vs_oracle"] = bool(
                        np.array_equal(d_lp[:nh].cpu().numpy().astype(np.uint64), o_pos))
                    derived_counts = d_counts[:1_000_000].cpu().numpy().astype(np.uint32)
                gpu.extend_seed_table(0)
                gpu.densify_suffix_array(0)
                # d_counts again from the plain path (what the parity sample below checks)
                gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream.cuda_stream)
                torch.cuda.synchronize()
                result["derived_structures"] = derived
            except capi.AwfmGpuError as e:  # e.g. not enough free HBM for the depth asked for
                result["derived_structures"] = {"error": str(e)}
                derived_counts = None
        else:
            derived_counts = None
        del d_lq, d_lc, d_lr, d_lh, d_lp
    else:
        derived_counts = None
        h_lq = h_lq_cpu = None

    # ---- parity on a sample + exact algorithmic bytes from the oracle (checker, not the product) ----
    sample = min(n, 1_000_000)
    h_counts_sample = d_counts[:sample].cpu().numpy().astype(np.uint32)
    h_letters_sample = d_letters[: sample * L].cpu().numpy()
    if arrays is not None:
        o_counts, _, work = harness.Oracle(arrays).count(h_letters_sample, fixed_len=L, threads=os.cpu_count())
        parity = bool(np.array_equal(o_counts, h_counts_sample))
        bytes_per_query = work["countBytes"] / sample
        result["parity_sample"] = {"queries": sample, "bit_exact_vs_oracle": parity,
                                   "lf_steps_per_query": work["lfSteps"] / sample,
                                   "block_reads_per_query": work["lfBlockReads"] / sample,
                                   "algorithmic_bytes_per_query": bytes_per_query}
        if not parity:
            raise SystemExit("PARITY FAILURE: CUDA counts differ from the oracle on the bench workload")
        if derived_counts is not None:
            ok = bool(np.array_equal(derived_counts[:sample], o_counts[: len(derived_counts[:sample])]))
            result["derived_structures"]["seed_table"]["bit_exact_vs_oracle"] = ok
            if not ok:
                raise SystemExit("PARITY FAILURE: counts through the derived seed table differ from the oracle")
    else:
        bytes_per_query = None

    # ---- e2e: the reference-facing drop-in call on host memory ----
    e2e = None
    threads = max(1, (os.cpu_count() or 1) // world)
    if not args.no_e2e:
        # host memory of one rank's list: 32-B entry + one malloc'd 32-B position list (48 B with its header) + letters
        ne = n
        try:
            import psutil
            room = psutil.virtual_memory().available // max(world, 1) // 2
            if ne * (80 + L) > room:
                ne = max(1_000_000, int(room // (80 + L)))
        except Exception:
            pass
        h_letters = torch.empty(ne * L, dtype=torch.uint8).pin_memory()
        h_letters.copy_(d_letters[: ne * L])
        hl = h_letters.numpy()
        ix = host_index_struct(arrays)
        ip = C.addressof(ix)
        t0 = time.time()
        sl = KmerSearchList(lib, ne).fill(hl, fixed_len=L)  # awFmCreateKmerSearchList: one position list per query, as the reference
        list_s = time.time() - t0
        assert lib.awFmGpuPrepareIndex(ip) == abi.AwFmSuccess  # one-time upload, reported separately
        lib.awFmParallelSearchCount(ip, sl.ptr, threads)
        if world > 1:
            dist.barrier()
        times = []
        for _ in range(args.e2e_steps):
            t1 = time.perf_counter()
            lib.awFmParallelSearchCount(ip, sl.ptr, threads)
            times.append(time.perf_counter() - t1)
        assert lib.awFmGpuLastCountStatus() == abi.AwFmSuccess
        e2e_counts = sl.entries()["count"][: min(sample, ne)]
        if not np.array_equal(e2e_counts, h_counts_sample[: len(e2e_counts)]):
            raise SystemExit("PARITY FAILURE: drop-in counts differ from the device-resident path")
        t_step = sum(times) / len(times)
        if world > 1:
            t = torch.tensor([t_step], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_step = float(t.item())
        e2e = {"value": world * ne / t_step, "unit": UNIT, "h2d_bytes_per_step": ne * L, "d2h_bytes_per_step": ne * 4,
               "call": "awFmParallelSearchCount(index, searchList, numThreads) drop-in, host AwFmKmerSearchList",
               "host_threads": threads, "ms_per_step": 1e3 * t_step, "search_list_setup_s": round(list_s, 2),
               "queries_per_gpu": ne}
        if ne != n:
            e2e["note"] = f"host memory bounds the list to {ne} of the {n} queries per rank"
        sl.close()
        del h_letters
        # the other half of the metric through the same door: awFmParallelSearchLocate on a host list
        if h_lq is not None:
            sl = KmerSearchList(lib, nl).fill(h_lq.numpy(), fixed_len=Ll)
            rc = lib.awFmParallelSearchLocate(ip, sl.ptr, threads)  # warm-up: grows the position lists that need it
            if world > 1:
                dist.barrier()
            times = []
            for _ in range(args.e2e_steps):
                t1 = time.perf_counter()
                rc = lib.awFmParallelSearchLocate(ip, sl.ptr, threads)
                times.append(time.perf_counter() - t1)
            if rc != abi.AwFmSuccess:
                raise SystemExit(f"drop-in awFmParallelSearchLocate returned {rc}")
            l_counts = sl.entries()["count"][:nl]
            l_hits = int(l_counts.sum(dtype=np.uint64))
            if arrays is not None and "locate" in result and "parity_sample" in result["locate"]:
                ls = result["locate"]["parity_sample"]["queries"]
                mine = sl.positions_flat(ls)
                same = bool(np.array_equal(l_counts[:ls].astype(np.uint64), np.diff(o_hit)) and np.array_equal(mine, o_pos))
                if not same:
                    raise SystemExit("PARITY FAILURE: drop-in positions differ from the oracle")
            else:
                same = None
            t_step = sum(times) / len(times)
            if world > 1:
                t = torch.tensor([t_step], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_step = float(t.item())
            e2e["locate"] = {"call": "awFmParallelSearchLocate(index, searchList, numThreads) drop-in, host AwFmKmerSearchList",
                             "queries_per_gpu": nl, "hits_per_gpu": l_hits, "ms_per_step": 1e3 * t_step,
                             "located_hits_per_s": world * l_hits / t_step, "queries_per_s": world * nl / t_step,
                             "h2d_bytes_per_step": nl * Ll, "d2h_bytes_per_step": (nl + 1) * 8 + l_hits * 8,
                             "bit_exact_vs_oracle_sample": same}
            sl.close()
        lib.awFmGpuReleaseIndex(ip)

    # ---- cpu baseline (rank 0, N=1): the unmodified reference on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        ns = min(args.cpu_sample, n)
        hs = d_letters[: ns * L].cpu().numpy()
        if harness.have_reference():
            ref = harness.Reference()
            ix = host_index_struct(arrays)
            best, times, r_counts = cpu_reference_pass(ref, C.addressof(ix), hs, ns, L, cores, reps=3)
            ok = bool(np.array_equal(r_counts[:sample], h_counts_sample[: len(r_counts[:sample])]))
            cpu = {"value": best, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"first {ns} of the {n} queries, 1 warm-up + best of 3 passes of the reference's "
                             f"awFmParallelSearchCount (oracle/_ref), numThreads={cores}",
                   "bit_exact_vs_cuda": ok}
            if h_lq_cpu is not None:  # located hits/s of the reference on the locate leg's queries (bounded sample)
                nq = len(h_lq_cpu) // Ll
                rsl = KmerSearchList(ref.lib, nq).fill(h_lq_cpu, fixed_len=Ll)
                ref.lib.awFmParallelSearchLocate(C.addressof(ix), rsl.ptr, cores)
                best_t = 1e30
                for _ in range(3):
                    t1 = time.perf_counter()
                    ref.lib.awFmParallelSearchLocate(C.addressof(ix), rsl.ptr, cores)
                    best_t = min(best_t, time.perf_counter() - t1)
                r_hits = int(rsl.entries()["count"][:nq].sum(dtype=np.uint64))
                rsl.close()
                cpu["locate"] = {"located_hits_per_s": r_hits / best_t, "queries_per_s": nq / best_t, "queries": nq,
                                 "hits": r_hits, "sample": f"first {nq} of the locate leg's {nl} {Ll}-mers, 1 warm-up + best "
                                                           f"of 3 passes of the reference's awFmParallelSearchLocate"}
        else:
            t1 = time.perf_counter()
            harness.Oracle(arrays).count(hs[: 1_000_000 * L], fixed_len=L, threads=cores)
            dt = time.perf_counter() - t1
            cpu = {"value": 1_000_000 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first 1000000 queries, scalar C oracle with OpenMP over queries"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    achieved = (bytes_per_query * n / (kernel_avg * 1e-3) / 1e9) if bytes_per_query else None
    traffic = ncu_traffic_per_launch("sweep" if stage_ms else "tile")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": kernel_avg,
                "algorithmic_bytes_per_launch": bytes_per_query * n if bytes_per_query else None}
    if stage_ms:
        names = ["clear+sweepPack", "radix sort (CUB)", "sweepStep<first>"] + \
                [f"sweepStep pass {i + 2}" for i in range(len(stage_ms) - 4)] + ["sweepIrregular"]
        roofline["kernel"] = ("sweep pipeline, one count call over the rank's whole batch: pack -> radix sort on the seed "
                              "index -> one sweepStep pass per LF step (csrc/awfm_sweep.cuh)")
        roofline["stages_ms"] = {k: round(v, 3) for k, v in zip(names, stage_ms)}
        roofline["note"] = ("algorithmic bytes (SURVEY 8d) charge every rank its own 104-B block read; the sweep orders the "
                            "live queries by range start, so queries on the same 128-B line share one DRAM fetch and the "
                            "index is streamed once per pass: achieved/peak may exceed 1, `traffic` is what DRAM really "
                            "moved (sum over the pipeline's kernels, ncu)")
        if traffic:
            roofline["dram_GBps"] = traffic / (kernel_avg * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_GBps"] / peak
        if tile_ms and bytes_per_query:
            t_traffic = ncu_traffic_per_launch("tile")
            roofline["tile_kernel"] = {"kernel": "countKernelV1 (one launch, one random line per rank)", "kernel_ms": tile_ms,
                                       "queries_per_s": n / tile_ms * 1e3,
                                       "achieved": bytes_per_query * n / (tile_ms * 1e-3) / 1e9,
                                       "frac": bytes_per_query * n / (tile_ms * 1e-3) / 1e9 / peak, "traffic": t_traffic}
    else:
        roofline["kernel"] = "countKernelV1 (one launch over the rank's whole batch)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "index": {"device_bytes": gpu.device_bytes(), "build_s": round(build_s, 2), "build_gpu_ms": round(build_ms, 1),
                  "tie_suffixes_resolved_on_host": tie_suffixes},
    }
    line.update(result)
    print(json.dumps(line))
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
