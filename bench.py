#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on every BASELINE config, one JSON line.

    python bench.py --gpus N --steps K --warmup W            our arm  (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N ...             the reference's own OpenMP/AVX2 path on the host cores

Headline workload = BASELINE.json configs[1]: 3.1 Gbp synthetic nucleotide index (seed k=12, SA ratio 8), 100 M random
20-mers per GPU, query-sharded.  One "step" = one pass of awFmParallelSearchCount's work over this rank's whole batch.
  value       queries/s, whole job, packed queries + index resident in HBM, CUDA-event timed; for N>1 every step's counts
              are pushed into rank 0's buffer by a copy-engine peer write over NVLink (no SM involved).
  e2e         queries/s through the reference-facing drop-in call awFmParallelSearchCount(index, searchList, threads)
              on HOST memory (32-B AwFmKmerSearchData entries pointing at host strings): packing, H2D, kernels, D2H and
              the scatter of `count` back into the structs inside the timed region.
  e2e_packed  the same metric through the additive packed-batch call (awfm_gpu_group_count = awFmGpuCountPacked): 2-bit
              packed 20-mers in page-locked host memory in, u32 counts in page-locked host memory out, chunk-pipelined
              H2D / search / D2H inside the timed region.
Further objects on the same line: cfg1 (configs[0]), locate / cfg3 (configs[2], SA ratios 1/8/16), cfg4 (configs[3]),
cfg5 (configs[4], run at every N), single_process_fanout (N>1: ONE process driving all N GPUs through the library),
roofline (sweep compulsory-traffic model, random-access probe measured in the run), cpu_baseline, index_verification
(SHA-256 of the device-built index against the digests of the reference-built one, tests/golden/cfg2_index_sha256.json).
The 3.1 Gbp index is built on the device (byte-identical to awFmCreateIndex: those digests) because the reference's CPU
build takes ~40 min; cfg 1's index is built by the reference itself.  The reference arm and the cpu_baseline legs search
the same indexes with the UNMODIFIED reference library (oracle/_ref/libawfm_ref.so).
"""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
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
    ap.add_argument("--locate-queries", type=int, default=10_000_000, help="cfg 3: random 16-mers located per GPU")
    ap.add_argument("--locate-kmer", type=int, default=16)
    ap.add_argument("--derived-seed-depth", type=int, default=16, help="0 = skip the derived-structures leg")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl", "none"],
                    help="N>1: how every step's counts reach rank 0 (p2p = copy-engine peer writes over NVLink)")
    ap.add_argument("--amino-residues", type=int, default=1_000_000_000, help="cfg 4 text length")
    ap.add_argument("--amino-queries", type=int, default=50_000_000)
    ap.add_argument("--cfg5-records", type=int, default=10_000)
    ap.add_argument("--cfg5-queries", type=int, default=10_000_000, help="cfg 5: sampled 32-mers per GPU")
    ap.add_argument("--skip", default="", help="comma list of legs to skip: cfg1,cfg3,cfg4,cfg5,fanout,derived,variable,dropin,packed,cpu")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.skip = set(x for x in a.skip.split(",") if x)
    if a.no_e2e:
        a.skip |= {"dropin", "packed"}
    if a.no_cpu_baseline:
        a.skip.add("cpu")
    return a


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


def ncu_traffic_record():
    """DRAM bytes of one count call over the bench batch, measured by ncu in a gpurun of this round and committed under
    profiles/ (a run under ncu is never a bench value, so the capture is a separate call): {file, sweep, tile}."""
    path = os.path.join(ROOT, "profiles", "ncu_count_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


class Env:
    """What every leg needs: the library, torch, the rank's device and stream, rank/world, the distributed module."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from avxwindowfmindex_b200 import capi
        self.args = args
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            # a second, CPU-side group: a rank waiting in an NCCL barrier keeps a spinning kernel on its GPU, which
            # time-slices against any other process using that GPU (the single-process fan-out leg does)
            self.cpu_group = dist.new_group(backend="gloo")
        self.lib = capi.load()  # raises if the CUDA library is missing: there is no fallback
        self.stream = torch.cuda.current_stream()
        self.cores = os.cpu_count() or 1
        self.threads = max(1, self.cores // self.world)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def cpu_barrier(self):
        """waits on the host only: the waiting ranks' GPUs stay idle"""
        if self.world > 1:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([x], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t)
        return int(t.item())

    def min_over_ranks(self, x):
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([int(x)], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item())

    def event_ms(self, fn, reps=5, warm=1, best=True):
        """device time of fn() on the launching stream, CUDA events, synchronize on both sides"""
        torch = self.torch
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        out = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            fn()
            b.record(self.stream)
            torch.cuda.synchronize()
            out.append(a.elapsed_time(b))
        return min(out) if best else sum(out) / len(out)


def synth_device(env, count, seed, start=0, amino=False):
    from avxwindowfmindex_b200 import capi
    t = env.torch.empty(count + 64, dtype=env.torch.uint8, device=env.dev)
    capi.check(env.lib.awfm_gpu_synth_letters(env.local, t.data_ptr(), count, seed, start, int(amino)))
    return t


def build_on_device(env, bp, seed_k, sa_ratio, text_seed, amino=False, d_text=None):
    from avxwindowfmindex_b200 import DeviceBuiltIndex, abi
    own = d_text is None
    if own:
        d_text = synth_device(env, bp, text_seed, 0, amino)
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), bp, abi.AwFmAlphabetAmino if amino else abi.AwFmAlphabetDna,
                                              seed_k, sa_ratio, device=env.local)
    if own:
        del d_text
    env.torch.cuda.synchronize()
    return built


def pack_bits_device(env, d_letters, n, L, amino=False):
    """2-/5-bit packing on the GPU with torch (bench SETUP only, never timed; format: include/awfm_gpu.h)."""
    torch = env.torch
    bits = 5 if amino else 2
    nbytes = (L * bits + 7) // 8
    per = 56 // bits  # letters per 64-bit accumulator (whole bytes' worth of headroom is not needed: see below)
    span = per * bits
    out = torch.empty((n, nbytes), dtype=torch.uint8, device=env.dev)
    if amino:
        table = torch.full((256,), 20, dtype=torch.int64, device=env.dev)
        for i, ch in enumerate(b"ACDEFGHIKLMNPQRSTVWY"):
            table[ch] = i
    step = 1 << 22
    for a in range(0, n, step):
        m = min(step, n - a)
        w = d_letters[a * L:(a + m) * L].view(m, L).to(torch.int64)
        code = table[w] if amino else ((w >> 1) ^ (w >> 2)) & 3
        accs = []  # accumulator c holds letters [c*per, (c+1)*per) in its low `span` bits
        for c0 in range(0, L, per):
            c1 = min(L, c0 + per)
            shifts = bits * torch.arange(c1 - c0, device=env.dev, dtype=torch.int64)
            accs.append((code[:, c0:c1] << shifts).sum(dim=1))
        cols = []
        for i in range(nbytes):
            c, o = divmod(8 * i, span)
            b = accs[c] >> o
            if o + 8 > span and c + 1 < len(accs):
                b = b | (accs[c + 1] << (span - o))
            cols.append(b & 0xFF)
        out[a:a + m] = torch.stack(cols, dim=1).to(torch.uint8)
    return out.reshape(-1)


def pinned_copy(env, d_tensor, dtype=np.uint8):
    from avxwindowfmindex_b200 import PinnedArray
    p = PinnedArray(d_tensor.numel(), dtype)
    env.torch.from_numpy(p.array).copy_(d_tensor if d_tensor.dtype != env.torch.int32 else d_tensor)
    return p


def wall_times(fn, reps, warm=1, before=None):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        if before is not None:
            before()  # untimed
        t0 = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t0)
    return out


def fasta_index_struct(arrays, meta):
    """struct AwFmIndex + FastaVector (record table only) over host arrays, for the reference's contig mapping"""
    from avxwindowfmindex_b200 import abi
    ix = arrays.as_awfm_index()
    fv = abi.FastaVector()
    fv.metadata.data = meta.ctypes.data
    fv.metadata.count = fv.metadata.capacity = len(meta)
    ix.fastaVector = C.addressof(fv)
    ix.featureFlags = 1
    return ix, fv


class PeerGather:
    """Every rank's result shard into ONE buffer on rank 0's GPU by copy-engine peer writes over NVLink/NVSwitch
    (include/awfm_gpu.h: awfm_gpu_ipc_* / awfm_gpu_peer_copy_async).  The root exports a device buffer through CUDA IPC,
    the handle travels by one broadcast at setup, and from then on a rank's push is a cudaMemcpyAsync on its own stream:
    no collective and no SM on the data path."""

    def __init__(self, env, total_bytes):
        from avxwindowfmindex_b200 import capi
        self.env, self.lib = env, env.lib
        torch, dist = env.torch, env.dist
        self.base = C.c_void_p()
        handle = torch.zeros(64, dtype=torch.uint8)
        if env.rank == 0:
            capi.check(self.lib.awfm_gpu_device_malloc(env.local, C.byref(self.base), total_bytes))
            h = (C.c_uint8 * 64)()
            capi.check(self.lib.awfm_gpu_ipc_export(env.local, self.base, h))
            handle = torch.tensor(list(h), dtype=torch.uint8)
        d = handle.to(env.dev)
        dist.broadcast(d, src=0)
        if env.rank != 0:
            h = (C.c_uint8 * 64)(*d.cpu().tolist())
            capi.check(self.lib.awfm_gpu_ipc_open(env.local, h, C.byref(self.base)))
        self.total_bytes = total_bytes

    def push(self, byte_offset, src_ptr, nbytes, stream):
        from avxwindowfmindex_b200 import capi
        capi.check(self.lib.awfm_gpu_peer_copy_async(self.env.local, self.base.value + byte_offset, src_ptr, nbytes, stream))

    def read_root(self, byte_offset, nbytes):
        """rank 0: the gathered bytes as a numpy array (check only; cudaMemcpyDefault copies device -> host as well)"""
        from avxwindowfmindex_b200 import PinnedArray, capi
        p = PinnedArray(nbytes, np.uint8)
        capi.check(self.lib.awfm_gpu_peer_copy_async(self.env.local, p.ptr, self.base.value + byte_offset, nbytes, None))
        self.env.torch.cuda.synchronize()
        out = np.array(p.array)
        p.close()
        return out

    def close(self):
        if not self.base:
            return
        self.env.torch.cuda.synchronize()
        if self.env.rank == 0:
            self.lib.awfm_gpu_device_free(self.env.local, self.base)
        else:
            self.lib.awfm_gpu_ipc_close(self.env.local, self.base)
        self.base = C.c_void_p()


# ----------------------------------------------------------------------------------------------- reference arm
def workload_config(args, world):
    gather = {"p2p": "every step's counts pushed into rank 0's buffer by copy-engine peer writes over NVLink (CUDA IPC), "
                     "overlapped with the next step's search",
              "nccl": "counts gathered to rank 0 over NCCL, overlapped with the next step's search",
              "none": "no gather"}[args.gather]
    return {
        "workload": f"count: {args.bp} bp synthetic nucleotide index (seed k={args.seed_k}, SA ratio {args.sa_ratio}), "
                    f"{args.queries} random {args.kmer}-mers per GPU (BASELINE.json configs[1])",
        "text_bp": args.bp, "seed_k": args.seed_k, "sa_ratio": args.sa_ratio, "kmer": args.kmer,
        "queries_per_gpu": args.queries,
        "parallelism": f"query-sharded x{world}, index replicated per GPU" + (f"; {gather}" if world > 1 else ""),
        "l2_policy": "inputs larger than L2 (index 3.5 GB + packed queries 2 GB per step vs 126 MB L2)",
        "index_built_by": "device builder; SHA-256 of every section equal to the reference-built index's "
                          "(index_verification, tests/golden/cfg2_index_sha256.json)",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the reference arm
    from avxwindowfmindex_b200 import KmerSearchList, synth
    from oracle import harness
    if not harness.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libawfm_ref.so was not built"}))
        return
    os.environ.pop("WORLD_SIZE", None)  # this arm is one process whatever N is
    env = Env(args)
    threads = env.cores
    config = workload_config(args, args.gpus)
    ref = harness.Reference()
    built = build_on_device(env, args.bp, args.seed_k, args.sa_ratio, synth.TEXT_SEED + 2)  # setup only, never timed
    arrays = built.to_host()
    built.close()
    env.torch.cuda.empty_cache()
    ix = arrays.as_awfm_index()
    n = min(args.cpu_sample, args.queries)
    letters = synth.random_queries(n, args.kmer, seed=synth.QUERY_SEED + 2)
    sl = KmerSearchList(ref.lib, n).fill(letters, fixed_len=args.kmer)
    for _ in range(args.warmup):
        ref.lib.awFmParallelSearchCount(C.addressof(ix), sl.ptr, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.lib.awFmParallelSearchCount(C.addressof(ix), sl.ptr, threads)
    total = time.perf_counter() - t0
    sl.close()
    value = n * args.steps / total
    sample = (f"{n} of the {args.queries} random {args.kmer}-mers per step, reference awFmParallelSearchCount, "
              f"numThreads={threads}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------- legs
def verify_index(arrays, args):
    """SHA-256 of every section of the device-built index against the digests of the index the reference's
    awFmCreateIndex builds from the same text on the CPU (tools/ref_index_hashes.py)."""
    from avxwindowfmindex_b200 import synth
    from avxwindowfmindex_b200.index import section_digests
    path = os.path.join(ROOT, "tests", "golden", "cfg2_index_sha256.json")
    if not os.path.exists(path):
        return {"checked": False, "why": "tests/golden/cfg2_index_sha256.json missing"}
    want = json.load(open(path))
    t = want["text"]
    if (t["length"], want["seed_k"], want["sa_ratio"], t["seed"]) != (args.bp, args.seed_k, args.sa_ratio, synth.TEXT_SEED + 2):
        return {"checked": False, "why": "bench arguments differ from the configuration the digests were taken on"}
    t0 = time.time()
    got = section_digests(arrays, with_chunks=False)
    if any(got[k]["sha256"] != w["sha256"] for k, w in want["sections"].items() if k in got):
        got = section_digests(arrays)  # a section differs: per-chunk digests say where
    out = {"checked": True, "reference_digests": "tests/golden/cfg2_index_sha256.json (awFmCreateIndex + divsufsort64 on the CPU)",
           "seconds": round(time.time() - t0, 1), "sections": {}}
    ok = True
    for name, w in want["sections"].items():
        g = got.get(name)
        same = bool(g and g["sha256"] == w["sha256"] and g["bytes"] == w["bytes"])
        out["sections"][name] = {"bytes": w["bytes"], "sha256": g["sha256"] if g else None, "equal_to_reference_built": same}
        if not same and g:
            out["sections"][name]["first_differing_64MiB_chunk"] = next(
                (i for i, (x, y) in enumerate(zip(g.get("chunks", []), w["chunks"])) if x != y), None)
        ok &= same
    out["byte_identical_to_reference_built_index"] = ok
    return out


def leg_cfg1(env):
    """BASELINE configs[0]: 1 Mbp random nucleotide text, seed k=8, SA ratio 8, 100 k random 12-mers, count.  The index
    is built by the reference itself (awFmCreateIndex); the reference is timed at 1, 2, 4, ... all threads."""
    from avxwindowfmindex_b200 import GpuGroup, KmerSearchList, PinnedArray, abi, pack_queries_bits, synth
    from avxwindowfmindex_b200.search import QUERY_2BIT, GpuIndex
    from oracle import harness
    torch, lib = env.torch, env.lib
    n, L, bp = 100_000, 12, 1_000_000
    text = synth.random_text(bp, seed=synth.TEXT_SEED + 1)
    letters = synth.random_queries(n, L, seed=synth.QUERY_SEED + 1)
    out = {"workload": "1 Mbp synthetic random nucleotide index (seed k=8, SA ratio 8), awFmParallelSearchCount on 100 k "
                       "random 12-mers (BASELINE.json configs[0])", "queries": n}
    ref = harness.Reference() if harness.have_reference() else None
    tmp = tempfile.mkdtemp(prefix="awfm_cfg1_")
    if ref is not None:
        ptr = ref.create_index(text.tobytes(), os.path.join(tmp, "cfg1.awfmi"), abi.AwFmAlphabetDna, 8, 8)
        arrays = ref.arrays(ptr)
        out["index_built_by"] = "reference awFmCreateIndex (libdivsufsort) on the host"
    else:
        built = build_on_device(env, bp, 8, 8, synth.TEXT_SEED + 1)
        arrays = built.to_host()
        built.close()
        ptr = None
        out["index_built_by"] = "device builder (oracle/_ref not available on this box)"
    gpu = GpuIndex(arrays, device=env.local)
    # device-resident
    d_q = torch.from_numpy(letters).to(env.dev)
    d_c = torch.zeros(n, dtype=torch.int32, device=env.dev)
    call = lambda: gpu.count_device(d_q.data_ptr(), None, L, n, d_c.data_ptr(), None, env.stream.cuda_stream)  # noqa: E731
    ms = [env.event_ms(call, reps=1, warm=0) for _ in range(30)][10:]
    counts = d_c.cpu().numpy().astype(np.uint32)
    out["device"] = {"ms_best": min(ms), "ms_median": sorted(ms)[len(ms) // 2], "queries_per_s": n / min(ms) * 1e3,
                     "kernel": "countKernelV1 (tile kernel; the batch is below the sweep threshold)", "reps": len(ms)}
    out["total_hits"] = int(counts.sum())
    # drop-in e2e on a host list
    ix = arrays.as_awfm_index()
    ip = C.addressof(ix)
    sl = KmerSearchList(lib, n).fill(letters, fixed_len=L)
    assert lib.awFmGpuPrepareIndex(ip) == abi.AwFmSuccess
    best = {}
    for t in (1, 4, env.cores):
        ts = wall_times(lambda: lib.awFmParallelSearchCount(ip, sl.ptr, t), reps=20, warm=5)
        best[t] = {"ms_best": 1e3 * min(ts), "ms_median": 1e3 * sorted(ts)[len(ts) // 2], "queries_per_s": n / min(ts)}
    dropin_ok = bool(np.array_equal(sl.counts(), counts))
    out["e2e_dropin"] = {"call": "awFmParallelSearchCount drop-in on a host AwFmKmerSearchList", "by_num_threads": best,
                         "queries_per_s": max(v["queries_per_s"] for v in best.values()),
                         "bit_exact_vs_device_path": dropin_ok}
    sl.close()
    lib.awFmGpuReleaseIndex(ip)
    # packed e2e
    group = GpuGroup(indexes=[gpu])
    packed = pack_queries_bits(letters, L)
    pin, pout = PinnedArray(len(packed), np.uint8), PinnedArray(n, np.uint32)
    pin.array[:] = packed
    ts = wall_times(lambda: group.count(pin.array, QUERY_2BIT, fixed_len=L, out=pout.array), reps=30, warm=10)
    out["e2e_packed"] = {"call": "awfm_gpu_group_count, 2-bit packed 12-mers, page-locked in/out", "ms_best": 1e3 * min(ts),
                         "ms_median": 1e3 * sorted(ts)[len(ts) // 2], "queries_per_s": n / min(ts),
                         "bit_exact_vs_device_path": bool(np.array_equal(pout.array, counts))}
    pin.close(), pout.close(), group.close()
    # the reference at 1, 2, 4, ... all threads
    if ref is not None:
        rsl = KmerSearchList(ref.lib, n).fill(letters, fixed_len=L)
        rows = {}
        t = 1
        while True:
            ts = wall_times(lambda: ref.lib.awFmParallelSearchCount(ptr, rsl.ptr, t), reps=12, warm=5)
            rows[t] = {"ms_best": 1e3 * min(ts), "ms_median": 1e3 * sorted(ts)[len(ts) // 2], "queries_per_s": n / min(ts)}
            if t >= env.cores:
                break
            t = min(env.cores, t * 2)
        r_counts = rsl.counts()
        rsl.close()
        out["reference"] = {"by_num_threads": rows, "cores": env.cores,
                            "best_queries_per_s": max(v["queries_per_s"] for v in rows.values()),
                            "sample": "all 100000 queries, 5 warm-up + 12 timed calls per thread count"}
        out["bit_exact_vs_reference"] = bool(np.array_equal(r_counts, counts))
        ref.dealloc_index(ptr)
    else:
        o_counts, _, _ = harness.Oracle(arrays).count(letters, fixed_len=L, threads=env.cores)
        out["bit_exact_vs_oracle"] = bool(np.array_equal(o_counts, counts))
    gpu.close()
    if not (dropin_ok and out["e2e_packed"]["bit_exact_vs_device_path"] and
            out.get("bit_exact_vs_reference", out.get("bit_exact_vs_oracle"))):
        raise SystemExit("PARITY FAILURE in the cfg 1 leg: " + json.dumps(out))
    return out


def locate_leg(env, gpu, arrays, d_lq, nl, Ll, label, reference_sample=1_000_000, e2e=True, dropin=False):
    """Device-resident locate (ranges -> scan -> expand -> walk) + packed end to end + parity sample + the reference on a
    bounded sample.  Returns the leg's dict."""
    from avxwindowfmindex_b200 import GpuGroup, KmerSearchList, PinnedArray, abi
    from avxwindowfmindex_b200.search import QUERY_2BIT, QUERY_5BIT
    from oracle import harness
    torch, stream = env.torch, env.stream.cuda_stream
    amino = bool(arrays.amino) if arrays is not None else False
    d_lc = torch.zeros(nl, dtype=torch.int32, device=env.dev)
    d_lr = torch.zeros((nl, 2), dtype=torch.int64, device=env.dev)
    d_lh = torch.zeros(nl + 1, dtype=torch.int64, device=env.dev)
    gpu.count_device(d_lq.data_ptr(), None, Ll, nl, d_lc.data_ptr(), d_lr.data_ptr(), stream)
    gpu.scan_ranges_device(d_lr.data_ptr(), nl, d_lh.data_ptr(), stream)
    hits = int(d_lh[-1].item())
    d_lp = torch.zeros(max(hits, 1), dtype=torch.int64, device=env.dev)

    def prepare():  # search + ranges of the queries with hits + hit offsets scanned from the counts
        gpu.locate_prepare_device(d_lq.data_ptr(), 0, Ll, nl, d_lc.data_ptr(), d_lr.data_ptr(), d_lh.data_ptr(), stream)

    def locate_all():
        prepare()
        gpu.locate_device(d_lr.data_ptr(), d_lh.data_ptr(), nl, 0, hits, d_lp.data_ptr(), stream)

    ms_all = env.event_ms(locate_all)
    ms_prepare = env.event_ms(prepare)
    ms_walk = env.event_ms(lambda: gpu.locate_device(d_lr.data_ptr(), d_lh.data_ptr(), nl, 0, hits, d_lp.data_ptr(), stream))
    # the same through the three separate calls with full range output (every query's final range, as the reference
    # leaves it), for the record
    ms_count = env.event_ms(lambda: gpu.count_device(d_lq.data_ptr(), None, Ll, nl, d_lc.data_ptr(), d_lr.data_ptr(), stream), reps=3)
    ms_scan = env.event_ms(lambda: gpu.scan_ranges_device(d_lr.data_ptr(), nl, d_lh.data_ptr(), stream), reps=3)
    # the front end through the sweep whatever the batch size (the automatic choice for this batch is in stages_ms)
    auto = 0 if env.args.count_path == "auto" else (1 if env.args.count_path == "sweep" else -1)
    gpu.set_tuning(sweep_min_queries=1)
    ms_prepare_sweep = env.event_ms(prepare, reps=3)
    gpu.set_tuning(sweep_min_queries=-1)
    ms_prepare_tile = env.event_ms(prepare, reps=3)
    gpu.set_tuning(sweep_min_queries=auto)
    prepare()
    loc = {"workload": label, "queries": nl, "hits": hits, "locate_ms": ms_all,
           "prepare_ms_by_count_path": {"sweep": ms_prepare_sweep, "tile_kernel": ms_prepare_tile}, "located_hits_per_s": hits / ms_all * 1e3,
           "locate_queries_per_s": nl / ms_all * 1e3,
           "stages_ms": {"search+ranges_of_hits+scan (awfm_gpu_locate_prepare_device)": ms_prepare, "expand+walk": ms_walk},
           "separate_calls_ms": {"count_with_all_ranges": ms_count, "scan_of_ranges": ms_scan},
           "walk_ms": ms_walk, "walk_hits_per_s": hits / ms_walk * 1e3}
    o_hit = o_pos = None
    if arrays is not None:
        ls = min(nl, 200_000)
        o_hit, o_pos, lwork = harness.Oracle(arrays).locate(d_lq[: ls * Ll].cpu().numpy(), fixed_len=Ll, threads=env.cores)
        nh = int(o_hit[-1])
        same = (np.array_equal(d_lh[: ls + 1].cpu().numpy().astype(np.uint64), o_hit)
                and np.array_equal(d_lp[:nh].cpu().numpy().astype(np.uint64), o_pos))
        loc["parity_sample"] = {"queries": ls, "hits": nh, "bit_exact_vs_oracle": bool(same)}
        if nh:
            loc["backtrace_steps_per_hit"] = lwork["backtraceSteps"] / nh
            loc["algorithmic_bytes_per_hit"] = lwork["locateBytes"] / nh
            loc["walk_algorithmic_GBps"] = lwork["locateBytes"] / nh * hits / ms_walk / 1e6
        if not same:
            raise SystemExit(f"PARITY FAILURE: CUDA positions differ from the oracle ({label})")
    if e2e and "packed" not in env.args.skip:
        group = GpuGroup(indexes=[gpu])
        fmt = QUERY_5BIT if amino else QUERY_2BIT
        d_bits = pack_bits_device(env, d_lq, nl, Ll, amino)
        pin = pinned_copy(env, d_bits)
        del d_bits
        ph, pp = PinnedArray(nl + 1, np.uint64), PinnedArray(max(hits, 1), np.uint64)
        call = lambda: group.locate(pin.array, fmt, fixed_len=Ll, out=(ph.array, pp.array))  # noqa: E731
        ts = wall_times(call, reps=env.args.e2e_steps, warm=2)
        t_step = env.max_over_ranks(sum(ts) / len(ts))
        same = bool(np.array_equal(ph.array, d_lh.cpu().numpy().astype(np.uint64)) and
                    np.array_equal(pp.array[:hits], d_lp[:hits].cpu().numpy().astype(np.uint64)))
        loc["e2e_packed"] = {"call": f"awfm_gpu_group_locate ({'5' if amino else '2'}-bit packed queries, page-locked in/out, CSR out)",
                             "ms_per_step": 1e3 * t_step, "located_hits_per_s": env.world * hits / t_step,
                             "queries_per_s": env.world * nl / t_step, "h2d_bytes_per_step": int(pin.array.nbytes),
                             "d2h_bytes_per_step": (nl + 1) * 8 + hits * 8, "bit_exact_vs_device_path": same}
        pin.close(), ph.close(), pp.close(), group.close()
        if not same:
            raise SystemExit(f"PARITY FAILURE: packed locate differs from the device-resident path ({label})")
    if dropin and arrays is not None and "dropin" not in env.args.skip:
        lib = env.lib
        h_lq = d_lq[: nl * Ll].cpu().numpy()
        ix = arrays.as_awfm_index()
        ip = C.addressof(ix)
        assert lib.awFmGpuPrepareIndex(ip) == abi.AwFmSuccess
        sl = KmerSearchList(lib, nl).fill(h_lq, fixed_len=Ll)
        rc = lib.awFmParallelSearchLocate(ip, sl.ptr, env.threads)  # warm-up: grows the position lists that need it
        env.barrier()
        ts = wall_times(lambda: lib.awFmParallelSearchLocate(ip, sl.ptr, env.threads), reps=env.args.e2e_steps, warm=0)
        l_counts = sl.entries()["count"][:nl]
        l_hits = int(l_counts.sum(dtype=np.uint64))
        same = None
        if o_hit is not None:
            ls = len(o_hit) - 1
            same = bool(np.array_equal(l_counts[:ls].astype(np.uint64), np.diff(o_hit)) and
                        np.array_equal(sl.positions_flat(ls), o_pos))
            if not same:
                raise SystemExit("PARITY FAILURE: drop-in positions differ from the oracle")
        t_step = env.max_over_ranks(sum(ts) / len(ts))
        loc["e2e_dropin"] = {"call": "awFmParallelSearchLocate(index, searchList, numThreads) drop-in, host AwFmKmerSearchList",
                             "return_code": int(rc), "hits_per_gpu": l_hits, "ms_per_step": 1e3 * t_step,
                             "located_hits_per_s": env.world * l_hits / t_step, "queries_per_s": env.world * nl / t_step,
                             "h2d_bytes_per_step": nl * Ll, "d2h_bytes_per_step": (nl + 1) * 8 + l_hits * 8,
                             "host_threads": env.threads, "bit_exact_vs_oracle_sample": same}
        sl.close()
        lib.awFmGpuReleaseIndex(ip)
    if (arrays is not None and env.rank == 0 and env.world == 1 and "cpu" not in env.args.skip and
            harness.have_reference() and reference_sample):
        ref = harness.Reference()
        ix = arrays.as_awfm_index()
        nq = min(nl, reference_sample)
        hq = d_lq[: nq * Ll].cpu().numpy()
        rsl = KmerSearchList(ref.lib, nq).fill(hq, fixed_len=Ll)
        ts = wall_times(lambda: ref.lib.awFmParallelSearchLocate(C.addressof(ix), rsl.ptr, env.cores), reps=3, warm=1)
        r_counts = rsl.entries()["count"][:nq]
        r_hits = int(r_counts.sum(dtype=np.uint64))
        same = bool(np.array_equal(np.diff(d_lh[: nq + 1].cpu().numpy()).astype(np.uint32), r_counts) and
                    np.array_equal(rsl.positions_flat(min(nq, 100_000)),
                                   d_lp[: int(d_lh[min(nq, 100_000)].item())].cpu().numpy().astype(np.uint64)))
        rsl.close()
        loc["reference"] = {"located_hits_per_s": r_hits / min(ts), "queries_per_s": nq / min(ts), "queries": nq,
                            "hits": r_hits, "cores": env.cores, "bit_exact_vs_cuda": same,
                            "sample": f"first {nq} queries, 1 warm-up + best of 3 calls of the reference's awFmParallelSearchLocate"}
        if not same:
            raise SystemExit(f"PARITY FAILURE: CUDA positions differ from the reference ({label})")
    return loc


def leg_cfg3_ratio(env, ratio):
    """BASELINE configs[2] at SA ratio 1 / 16: same text, index rebuilt on the device with that ratio."""
    from avxwindowfmindex_b200 import synth
    a = env.args
    built = build_on_device(env, a.bp, a.seed_k, ratio, synth.TEXT_SEED + 2)
    gpu = built.gpu_index()
    arrays = built.to_host()
    built.close()
    env.torch.cuda.empty_cache()
    d_lq = synth_device(env, a.locate_queries * a.locate_kmer, synth.QUERY_SEED + 3, 0)
    out = locate_leg(env, gpu, arrays, d_lq, a.locate_queries, a.locate_kmer,
                     f"{a.locate_queries} random {a.locate_kmer}-mers, {a.bp} bp index, SA ratio {ratio} (BASELINE.json configs[2])")
    out["sa_ratio"] = ratio
    out["device_bytes"] = gpu.device_bytes()
    gpu.close()
    del d_lq, arrays
    env.torch.cuda.empty_cache()
    return out


def leg_cfg4(env):
    """BASELINE configs[3]: 1 G-residue amino index (seed k=5; SA ratio 8), count + locate of 50 M random 8-mers."""
    from avxwindowfmindex_b200 import GpuGroup, KmerSearchList, PinnedArray, synth
    from avxwindowfmindex_b200.search import QUERY_5BIT
    from oracle import harness
    a, torch, stream = env.args, env.torch, env.stream.cuda_stream
    bp, n, L = a.amino_residues, a.amino_queries, 8
    built = build_on_device(env, bp, 5, 8, synth.TEXT_SEED + 4, amino=True)
    gpu = built.gpu_index()
    arrays = built.to_host()
    out = {"workload": f"{bp}-residue synthetic amino index (seed k=5, SA ratio 8), count + locate of {n} random 8-mers "
                       "(BASELINE.json configs[3])", "index_build_gpu_ms": built.build_ms, "device_bytes": gpu.device_bytes()}
    built.close()
    torch.cuda.empty_cache()
    d_q = synth_device(env, n * L, synth.QUERY_SEED + 4, 0, amino=True)
    d_c = torch.zeros(n, dtype=torch.int32, device=env.dev)
    call = lambda: gpu.count_device(d_q.data_ptr(), None, L, n, d_c.data_ptr(), None, stream)  # noqa: E731
    ms = env.event_ms(call)
    gpu.set_tuning(sweep_min_queries=-1)
    ms_tile = env.event_ms(call, reps=3)
    gpu.set_tuning(sweep_min_queries=0)
    call()
    torch.cuda.synchronize()
    counts_sample = d_c[:1_000_000].cpu().numpy().astype(np.uint32)
    h_sample = d_q[: 1_000_000 * L].cpu().numpy()
    o_counts, _, work = harness.Oracle(arrays).count(h_sample, fixed_len=L, threads=env.cores)
    ok = bool(np.array_equal(o_counts, counts_sample))
    out["count"] = {"ms": ms, "queries_per_s": n / ms * 1e3, "path": "sweep", "tile_kernel_ms": ms_tile,
                    "tile_kernel_queries_per_s": n / ms_tile * 1e3, "bit_exact_vs_oracle_sample": ok,
                    "lf_steps_per_query": work["lfSteps"] / len(o_counts),
                    "algorithmic_bytes_per_query": work["countBytes"] / len(o_counts)}
    if not ok:
        raise SystemExit("PARITY FAILURE: cfg 4 counts differ from the oracle")
    if "packed" not in a.skip:
        group = GpuGroup(indexes=[gpu])
        d_bits = pack_bits_device(env, d_q, n, L, amino=True)
        pin = pinned_copy(env, d_bits)
        del d_bits
        pout = PinnedArray(n, np.uint32)
        ts = wall_times(lambda: group.count(pin.array, QUERY_5BIT, fixed_len=L, out=pout.array), reps=a.e2e_steps, warm=2)
        same = bool(np.array_equal(pout.array[:1_000_000], counts_sample))
        out["count"]["e2e_packed"] = {"call": "awfm_gpu_group_count, 5-bit packed 8-mers, page-locked in/out",
                                      "ms_per_step": 1e3 * min(ts), "queries_per_s": n / min(ts),
                                      "h2d_bytes_per_step": int(pin.array.nbytes), "d2h_bytes_per_step": 4 * n,
                                      "bit_exact_vs_device_path_sample": same}
        pin.close(), pout.close(), group.close()
        if not same:
            raise SystemExit("PARITY FAILURE: cfg 4 packed counts differ from the device-resident path")
    if harness.have_reference() and "cpu" not in a.skip:
        ref = harness.Reference()
        ix = arrays.as_awfm_index()
        ns = min(n, 5_000_000)
        hs = d_q[: ns * L].cpu().numpy()
        rsl = KmerSearchList(ref.lib, ns).fill(hs, fixed_len=L)
        ts = wall_times(lambda: ref.lib.awFmParallelSearchCount(C.addressof(ix), rsl.ptr, env.cores), reps=3, warm=1)
        r_counts = rsl.counts()
        rsl.close()
        out["count"]["reference"] = {"queries_per_s": ns / min(ts), "queries": ns, "cores": env.cores,
                                     "bit_exact_vs_cuda": bool(np.array_equal(r_counts[:1_000_000], counts_sample))}
    out["locate"] = locate_leg(env, gpu, arrays, d_q, n, L, f"locate of the same {n} amino 8-mers", reference_sample=2_000_000)
    gpu.close()
    return out


def leg_cfg5(env):
    """BASELINE configs[4]: multi-sequence FASTA (10 k contigs, ~1 Gbp), locate sampled 32-mers with the sampled SA in
    HBM and map every hit to (contig, offset) on the device; every rank searches its own shard of the query stream and
    pushes its (position, contig, offset) rows into rank 0's buffer by copy-engine peer writes.  EVERY query is checked
    against the place it was cut from."""
    from avxwindowfmindex_b200 import GpuGroup, KmerSearchList, PinnedArray, abi, synth
    from avxwindowfmindex_b200.search import QUERY_2BIT
    from oracle import harness
    a, torch, dist, stream = env.args, env.torch, env.dist, env.stream.cuda_stream
    dev, rank, world = env.dev, env.rank, env.world
    records, L, n = a.cfg5_records, 32, a.cfg5_queries
    lengths = synth.multi_fasta_lengths(records, 50_000, 150_000, seed=synth.TEXT_SEED + 5)
    ends = np.cumsum(lengths + 1)
    total = int(ends[-1])
    header_ends = np.cumsum([len(b"contig%d" % i) + 1 for i in range(records)])
    meta = np.stack([header_ends.astype(np.uint64), ends.astype(np.uint64)], axis=1)
    d_text = synth_device(env, total, synth.TEXT_SEED + 5, 0)
    d_text[torch.from_numpy(ends - 1).to(dev)] = 0
    built = build_on_device(env, total, 12, 8, 0, d_text=d_text)
    gpu = built.gpu_index()
    gpu.set_sequences(meta)
    arrays = built.to_host() if rank == 0 else None
    build_ms = built.build_ms
    built.close()
    starts = np.concatenate([[0], ends[:-1]])
    z = synth.splitmix64(synth.QUERY_SEED + 5, rank * 2 * n, 2 * n)
    rec = (z[:n] % np.uint64(records)).astype(np.int64)
    off = (z[n:] % (lengths[rec] - L + 1).astype(np.uint64)).astype(np.int64)
    g = starts[rec] + off
    d_g, d_rec, d_off = torch.from_numpy(g).to(dev), torch.from_numpy(rec).to(dev), torch.from_numpy(off).to(dev)
    d_q = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
    ar = torch.arange(L, device=dev)
    for s in range(0, n, 1 << 20):
        e = min(n, s + (1 << 20))
        d_q[s * L:e * L] = d_text[(d_g[s:e, None] + ar[None, :]).reshape(-1)]
    del d_text
    torch.cuda.empty_cache()
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ranges = torch.zeros((n, 2), dtype=torch.int64, device=dev)
    d_hit = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), stream)
    hits = int(d_hit[-1].item())
    d_pos, d_seq, d_loc = (torch.zeros(hits, dtype=torch.int64, device=dev) for _ in range(3))
    # hit totals of every rank (deterministic per rank): where each rank's rows go in rank 0's buffer
    totals = [hits]
    if world > 1:
        t = torch.zeros(world, dtype=torch.int64, device=dev)
        t[rank] = hits
        dist.all_reduce(t)
        totals = [int(x) for x in t.tolist()]
    all_hits = sum(totals)
    base = sum(totals[:rank])
    gather = PeerGather(env, 3 * 8 * all_hits) if world > 1 and a.gather == "p2p" else None

    def step():
        gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
        gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), stream)
        gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, 0, hits, d_pos.data_ptr(), stream)
        gpu.map_positions_device(d_pos.data_ptr(), hits, d_seq.data_ptr(), d_loc.data_ptr(), stream)
        if gather is not None:  # rows land at this rank's offset of the three global arrays on rank 0
            for j, src in enumerate((d_pos, d_seq, d_loc)):
                gather.push(8 * (j * all_hits + base), src.data_ptr(), 8 * hits, stream)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    env.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(3, min(a.steps, 10))
    ev0.record(env.stream)
    for _ in range(steps):
        step()
    ev1.record(env.stream)
    torch.cuda.synchronize()
    env.barrier()
    ms = env.max_over_ranks(ev0.elapsed_time(ev1) / steps)
    ms_count = env.event_ms(lambda: gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream), reps=3)
    # which path the count took (automatic choice; 32-mers on a k = 12 table: 20 letters left of the seed, sweepRefill)
    # and the tile kernel on the same batch beside it
    gpu.set_tuning(sweep_profile=1)
    gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    torch.cuda.synchronize()
    count_stages = gpu.sweep_stage_ms()
    gpu.set_tuning(sweep_profile=0, sweep_ordered_emit=0)  # survivors' counts and ranges scattered straight from the last pass
    ms_count_scatter = env.event_ms(lambda: gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream), reps=3)
    gpu.set_tuning(sweep_ordered_emit=1, sweep_min_queries=-1)
    d_ranges_tile = torch.zeros_like(d_ranges)
    ms_count_tile = env.event_ms(lambda: gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges_tile.data_ptr(), stream), reps=3)
    tile_same = bool(torch.equal(d_ranges_tile, d_ranges))
    del d_ranges_tile
    gpu.set_tuning(sweep_min_queries=0)
    ms_walk = env.event_ms(lambda: gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, 0, hits, d_pos.data_ptr(), stream), reps=3)
    ms_map = env.event_ms(lambda: gpu.map_positions_device(d_pos.data_ptr(), hits, d_seq.data_ptr(), d_loc.data_ptr(), stream), reps=3)
    # by-construction check of EVERY query of this rank: the (contig, offset) it was cut from is among its hits
    cnt = d_hit[1:] - d_hit[:-1]
    q_of_hit = torch.repeat_interleave(torch.arange(n, device=dev), cnt)
    match = (d_seq == d_rec[q_of_hit]) & (d_loc == d_off[q_of_hit]) & (d_pos == d_g[q_of_hit])
    found = torch.zeros(n, dtype=torch.int32, device=dev).index_add_(0, q_of_hit, match.to(torch.int32))
    ok = bool((found >= 1).all().item()) and bool((cnt >= 1).all().item())
    ok = ok and bool(((d_loc + L) <= torch.from_numpy(lengths).to(dev)[d_seq]).all().item())
    gathered_ok = None
    if gather is not None:
        env.barrier()
        if rank == 0:  # the gathered arrays hold rank 0's own rows at their head
            got = gather.read_root(0, 8 * hits).view(np.int64)
            gathered_ok = bool(np.array_equal(got, d_pos.cpu().numpy()))
            tail = gather.read_root(8 * (all_hits - 1), 8).view(np.int64)  # the last rank's last row arrived
            gathered_ok = gathered_ok and bool(tail[0] != 0 or all_hits == hits)
    ok_all = bool(env.min_over_ranks(int(ok)))
    out = {"workload": f"multi-sequence FASTA ({records} contigs, {total} bp incl. separators, seed k=12, SA ratio 8), locate "
                       f"{n} sampled 32-mers per GPU + contig mapping (BASELINE.json configs[4])",
           "n_gpus": world, "queries_per_gpu": n, "hits_total": all_hits, "ms_per_step": ms, "steps": steps,
           "locate_queries_per_s": world * n / ms * 1e3, "located_and_mapped_hits_per_s": all_hits / ms * 1e3,
           "kernel_ms_rank0": {"count_with_ranges": ms_count, "expand+walk": ms_walk, "contig_map": ms_map},
           "count_path": "sweep (20 passes, letters 17-20 through sweepRefill, survivors through sweepEmit)" if count_stages else "tile kernel",
           "count_stages_ms": [round(x, 3) for x in count_stages], "count_tile_kernel_ms": ms_count_tile,
           "count_without_ordered_emit_ms": ms_count_scatter,
           "count_ranges_equal_to_tile_kernel_all_queries": tile_same,
           "every_query_found_at_its_origin": ok_all, "index_build_gpu_ms": build_ms, "scaling": "weak",
           "gather": ("copy-engine peer writes of (position, contig, offset) rows into rank 0's buffer, inside the timed step"
                      if gather is not None else None), "gathered_rows_check": gathered_ok}
    # packed end to end: host queries in, CSR + (contig, offset) per hit out
    if "packed" not in a.skip:
        group = GpuGroup(indexes=[gpu])
        d_bits = pack_bits_device(env, d_q, n, L)
        pin = pinned_copy(env, d_bits)
        del d_bits
        ph = PinnedArray(n + 1, np.uint64)
        pp, ps, pl = (PinnedArray(max(hits, 1), np.uint64) for _ in range(3))
        call = lambda: group.locate(pin.array, QUERY_2BIT, fixed_len=L, mapped=True, out=(ph.array, pp.array, ps.array, pl.array))  # noqa: E731
        ts = wall_times(call, reps=a.e2e_steps, warm=2)
        t_step = env.max_over_ranks(sum(ts) / len(ts))
        same = bool(np.array_equal(pp.array[:hits], d_pos.cpu().numpy().astype(np.uint64)) and
                    np.array_equal(ps.array[:hits], d_seq.cpu().numpy().astype(np.uint64)) and
                    np.array_equal(pl.array[:hits], d_loc.cpu().numpy().astype(np.uint64)))
        same = bool(env.min_over_ranks(int(same)))
        out["e2e_packed"] = {"call": "awfm_gpu_group_locate with contig mapping (2-bit packed queries, page-locked in/out)",
                             "ms_per_step": 1e3 * t_step, "located_and_mapped_hits_per_s": all_hits / t_step,
                             "queries_per_s": world * n / t_step, "h2d_bytes_per_step": int(pin.array.nbytes),
                             "d2h_bytes_per_step": (n + 1) * 8 + 24 * hits, "bit_exact_vs_device_path": same}
        if world == 1:  # hits per walk window of the pipeline (default 2^22): smaller windows shorten fill and drain
            sweep = {}
            for wh in (1 << 20, 1 << 21, 1 << 22):
                group.set_tuning(packed_window_hits=wh)
                sweep[str(wh)] = 1e3 * min(wall_times(call, reps=3, warm=1))
            out["e2e_packed"]["ms_by_window_hits"] = sweep
        for p in (pin, ph, pp, ps, pl):
            p.close()
        group.close()
    # the unmodified reference on a bounded sample (rank 0)
    if rank == 0 and harness.have_reference() and "cpu" not in a.skip:
        ref = harness.Reference()
        ix, fv = fasta_index_struct(arrays, meta)
        ip = C.addressof(ix)
        ns = min(n, 200_000)
        hq = d_q[: ns * L].cpu().numpy()
        h_hit = d_hit[: ns + 1].cpu().numpy().astype(np.uint64)
        nh = int(h_hit[-1])
        h_pos, h_seq, h_loc = (x[:nh].cpu().numpy().astype(np.uint64) for x in (d_pos, d_seq, d_loc))
        sl = KmerSearchList(ref.lib, ns).fill(hq, fixed_len=L)
        ts = wall_times(lambda: ref.lib.awFmParallelSearchLocate(ip, sl.ptr, env.cores), reps=2, warm=1)
        same_pos = bool(np.array_equal(sl.positions_flat(), h_pos))
        m = min(nh, 20_000)
        same_map = all(ref.contig_of(ip, int(h_pos[i])) == (abi.AwFmSuccess, int(h_seq[i]), int(h_loc[i])) for i in range(m))
        sl.close()
        out["reference"] = {"queries": ns, "cores": env.cores, "locate_queries_per_s": ns / min(ts),
                            "located_hits_per_s": nh / min(ts), "positions_bit_exact": same_pos,
                            "contig_mapping_checked_hits": m, "contig_mapping_identical": bool(same_map)}
        if not (same_pos and same_map):
            raise SystemExit("PARITY FAILURE: cfg 5 positions / contig mapping differ from the reference")
    if gather is not None:
        gather.close()
    gpu.close()
    if not ok_all:
        raise SystemExit("PARITY FAILURE: cfg 5: a query was not found at the place it was cut from")
    return out


def leg_hostlink(env):
    """What the box's host<->device links deliver with all N GPUs copying at once from / to page-locked memory: H2D
    alone, D2H alone, both directions together (aggregate GB/s, max-over-ranks time).  The end-to-end numbers are bound
    by these: a packed count query costs 5 B in and 4 B out."""
    torch = env.torch
    nbytes = 1 << 28
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out.fill_(2)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    s1, s2 = torch.cuda.Stream(device=env.dev), torch.cuda.Stream(device=env.dev)
    out = {}
    for name in ("h2d", "d2h", "both"):
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            env.cpu_barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                if name in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_in, non_blocking=True)
                if name in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = env.max_over_ranks(time.perf_counter() - t0)
            best = max(best, 4 * nbytes * env.world * (2 if name == "both" else 1) / dt / 1e9)
        out[name + "_GBps"] = best
    out["what"] = (f"aggregate over {env.world} GPU(s) copying at once, 4 x 256 MiB per direction and GPU, page-locked host "
                   "memory, best of 3; `both` counts the bytes of both directions")
    return out


def leg_fanout(env, arrays, h_bits_rank0, want_counts_rank0):
    """N>1, rank 0 alone (the other ranks wait): ONE process drives all N GPUs through the library's own fan-out
    (awfm_gpu_group_create over all devices).  N x 100 M 2-bit packed 20-mers in page-locked host memory,
    counts land in one host array; plus the unchanged awFmParallelSearchCount with AWFM_GPU_DEVICES=all."""
    from avxwindowfmindex_b200 import GpuGroup, KmerSearchList, PinnedArray, abi, synth
    from avxwindowfmindex_b200.search import QUERY_2BIT
    a = env.args
    G, n, L = env.world, a.queries, a.kmer
    t0 = time.time()
    group = GpuGroup(arrays, devices=list(range(G)))
    upload_s = time.time() - t0
    qb = (L + 3) // 4
    pin = PinnedArray(G * n * qb, np.uint8)
    for g in range(G):  # the same batch on every shard: every shard's answer is known
        pin.array[g * n * qb:(g + 1) * n * qb] = h_bits_rank0
    pout = PinnedArray(G * n, np.uint32)
    call = lambda: group.count(pin.array, QUERY_2BIT, fixed_len=L, out=pout.array)  # noqa: E731
    ts = wall_times(call, reps=a.e2e_steps, warm=2)
    same = all(np.array_equal(pout.array[g * n:(g + 1) * n], want_counts_rank0) for g in range(G))
    out = {"what": "one process, all GPUs: awfm_gpu_group_count over a device group", "devices": G,
           "queries": G * n, "ms_per_call": 1e3 * min(ts), "queries_per_s": G * n / min(ts),
           "index_replication_s": round(upload_s, 2), "bit_exact": bool(same), "launches": group.stats()["launches"]}
    group.close()
    pin.close(), pout.close()
    # the unchanged entry point on all GPUs
    if "dropin" not in a.skip:
        lib = env.lib
        os.environ["AWFM_GPU_DEVICES"] = "all"
        try:
            ix = arrays.as_awfm_index()
            ip = C.addressof(ix)
            nl = min(n, 50_000_000)
            letters = synth.random_queries(nl, L, seed=synth.QUERY_SEED + 2)
            sl = KmerSearchList(lib, nl).fill(letters, fixed_len=L)
            assert lib.awFmGpuPrepareIndex(ip) == abi.AwFmSuccess
            devices = lib.awFmGpuNumDevices(ip)
            ts = wall_times(lambda: lib.awFmParallelSearchCount(ip, sl.ptr, env.cores), reps=a.e2e_steps, warm=1)
            ok = bool(np.array_equal(sl.counts()[:1_000_000], want_counts_rank0[:1_000_000]))
            out["dropin_all_gpus"] = {"call": "awFmParallelSearchCount with AWFM_GPU_DEVICES=all (query strings back to back in PAGEABLE memory: "
                                              "the engine copies them to page-locked staging)", "devices": int(devices),
                                      "queries": nl, "ms_per_call": 1e3 * min(ts), "queries_per_s": nl / min(ts),
                                      "host_threads": env.cores, "bit_exact_sample": ok}
            sl.close()
            lib.awFmGpuReleaseIndex(ip)
        finally:
            del os.environ["AWFM_GPU_DEVICES"]
    return out


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    from avxwindowfmindex_b200 import GpuGroup, KmerSearchList, PinnedArray, abi, capi, synth
    from avxwindowfmindex_b200.search import QUERY_2BIT
    from oracle import harness  # checker + cpu_baseline only

    env = Env(args)
    torch, dist, lib = env.torch, env.dist, env.lib
    world, rank, local, dev, stream = env.world, env.rank, env.local, env.dev, env.stream
    result = {}

    # ---- index: built on this GPU, stays resident ----
    t0 = time.time()
    built = build_on_device(env, args.bp, args.seed_k, args.sa_ratio, synth.TEXT_SEED + 2)
    gpu = built.gpu_index()
    build_s = time.time() - t0
    need_host = rank == 0 or "dropin" not in args.skip
    arrays = built.to_host() if need_host else None
    tie_suffixes, tie_rounds, build_ms = built.tie_suffixes, built.tie_rounds, built.build_ms
    built.close()
    torch.cuda.empty_cache()
    if rank == 0:
        result["index_verification"] = verify_index(arrays, args)

    # ---- queries: this rank's shard of the random k-mer stream, resident in HBM ----
    n, L = args.queries, args.kmer
    d_letters = synth_device(env, n * L, synth.QUERY_SEED + 2, rank * n * L)
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    if args.count_path != "auto":
        gpu.set_tuning(sweep_min_queries=1 if args.count_path == "sweep" else -1)
    # N > 1: every step's counts go to rank 0.  p2p: a copy-engine peer write on this rank's stream right behind the
    # search (two count buffers, so step s+1 searches while step s's counts travel); nccl: torch.distributed.gather.
    bufs = [d_counts, torch.zeros(n, dtype=torch.int32, device=dev)] if world > 1 else [d_counts]
    gather_mode = args.gather if world > 1 else "none"
    peer = None
    if gather_mode == "p2p":
        try:
            peer = PeerGather(env, world * n * 4)
        except Exception as e:  # e.g. IPC not permitted in this container: say so and use the collective
            gather_mode = "nccl"
            result["gather_fallback"] = f"p2p unavailable ({e}); NCCL gather used"
        gather_mode = "p2p" if env.min_over_ranks(int(gather_mode == "p2p")) else "nccl"
    works = [None] * len(bufs)
    gathered = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(world)] if (gather_mode == "nccl" and rank == 0) else None
    copy_stream = torch.cuda.Stream(device=dev) if gather_mode == "p2p" else None
    searched = [torch.cuda.Event() for _ in bufs]
    pushed = [torch.cuda.Event() for _ in bufs]

    def step(s):
        b = s % len(bufs)
        if gather_mode == "nccl" and works[b] is not None:
            works[b].wait()
            works[b] = None
        if gather_mode == "p2p":
            stream.wait_event(pushed[b])  # the push that last read this buffer
        gpu.count_device(d_letters.data_ptr(), None, L, n, bufs[b].data_ptr(), None, stream.cuda_stream)
        if gather_mode == "p2p":
            searched[b].record(stream)
            copy_stream.wait_event(searched[b])
            peer.push(rank * n * 4, bufs[b].data_ptr(), n * 4, copy_stream.cuda_stream)
            pushed[b].record(copy_stream)
        elif gather_mode == "nccl":
            works[b] = dist.gather(bufs[b], gathered if rank == 0 else None, dst=0, async_op=True)

    def drain():
        for b, w in enumerate(works):
            if w is not None:
                w.wait()
                works[b] = None
        if gather_mode == "p2p":
            for e in pushed:
                stream.wait_event(e)

    for s in range(args.warmup):
        step(s)
    drain()
    torch.cuda.synchronize()
    launches_per_step = int(gpu.stats()["launches"])  # kernels of ours in one count call (same for every step)
    env.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local) as clocks:
        ev[0].record(stream)
        for s in range(args.steps):
            step(s)
        drain()
        ev[1].record(stream)
        torch.cuda.synchronize()
        env.barrier()
        torch.cuda.synchronize()
    launches = launches_per_step * args.steps
    ms_per_step = env.max_over_ranks(ev[0].elapsed_time(ev[1])) / args.steps
    value = world * n / (ms_per_step * 1e-3)
    if peer is not None:
        env.barrier()
        if rank == 0:  # what arrived: rank 0's own shard at the head, the last rank's at the tail
            got = peer.read_root(0, 4_000_000).view(np.int32)
            result["gather_check"] = {"mode": "p2p", "rank0_shard_head_equal": bool(np.array_equal(got, bufs[(args.steps - 1) % len(bufs)][:1_000_000].cpu().numpy()))}
        env.barrier()
        peer.close()
        peer = None

    # ---- device time of one count call over the whole batch, stage times, live counts (roofline inputs) ----
    call = lambda: gpu.count_device(d_letters.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream.cuda_stream)  # noqa: E731
    kernel_avg = env.event_ms(call, reps=5, best=False)
    gpu.set_tuning(sweep_profile=1)
    call()
    torch.cuda.synchronize()
    stage_ms = gpu.sweep_stage_ms()  # [] when the call took the tile kernel
    live, irregular = gpu.sweep_live() if stage_ms else ([], 0)
    gpu.set_tuning(sweep_profile=0)
    tile_ms = None
    if stage_ms and args.count_path == "auto":  # the single-kernel path on the same batch, for the record
        gpu.set_tuning(sweep_min_queries=-1)
        tile_ms = env.event_ms(call, reps=3, best=False)
        gpu.set_tuning(sweep_min_queries=0)
        call()
        torch.cuda.synchronize()
    # random-access roofline, measured now on this GPU: independent random 32-B sector reads over a 4 GiB array, one
    # lane per read (the access shape of one nucleotide rank)
    random_access = None
    if rank == 0:
        try:
            gb = C.c_double()
            capi.check(lib.awfm_gpu_gather_bandwidth(local, 4 << 30, 32, 1 << 28, 1, C.byref(gb)))
            lines = gb.value * 1e9 / 32
            gb128 = C.c_double()
            capi.check(lib.awfm_gpu_gather_bandwidth(local, 4 << 30, 128, 1 << 27, 4, C.byref(gb128)))
            random_access = {"lines_per_s": lines, "GBps_at_128B_per_line": lines * 128 / 1e9,
                             "probe": "awfm_gpu_gather_bandwidth: 2^28 independent random 32-B reads over 4 GiB, one lane per read, best of 3",
                             "lines_per_s_128B_reads_4_lanes": gb128.value * 1e9 / 128}
        except capi.AwfmGpuError as e:
            random_access = {"error": str(e)}

    # ---- cfg 3 at this index's SA ratio: located hits/s (device-resident, packed e2e, drop-in e2e, reference) ----
    nl, Ll = args.locate_queries, args.locate_kmer
    if nl > 0:
        d_lq = synth_device(env, nl * Ll, synth.QUERY_SEED + 3, rank * nl * Ll)
        result["locate"] = locate_leg(env, gpu, arrays if rank == 0 or "dropin" not in args.skip else None, d_lq, nl, Ll,
                                      f"{nl} random {Ll}-mers per GPU on the same index (SA ratio {args.sa_ratio}), BASELINE.json configs[2]",
                                      dropin=True)
        result["locate"]["sa_ratio"] = args.sa_ratio
        # ---- derived structures (opt-in: HBM for fewer dependent DRAM round trips), same queries, same checks ----
        if args.derived_seed_depth > args.seed_k and world == 1 and "derived" not in args.skip:
            try:
                derived = {}
                d_lc = torch.zeros(nl, dtype=torch.int32, device=dev)
                d_lr = torch.zeros((nl, 2), dtype=torch.int64, device=dev)
                d_lh = torch.zeros(nl + 1, dtype=torch.int64, device=dev)
                gpu.count_device(d_lq.data_ptr(), None, Ll, nl, d_lc.data_ptr(), d_lr.data_ptr(), stream.cuda_stream)
                gpu.scan_ranges_device(d_lr.data_ptr(), nl, d_lh.data_ptr(), stream.cuda_stream)
                hits = int(d_lh[-1].item())
                d_lp = torch.zeros(max(hits, 1), dtype=torch.int64, device=dev)
                walk = lambda: gpu.locate_device(d_lr.data_ptr(), d_lh.data_ptr(), nl, 0, hits, d_lp.data_ptr(), stream.cuda_stream)  # noqa: E731
                walk()
                plain = d_lp.clone()
                if args.derived_seed_depth > args.seed_k + 2:  # a modest table first: 4^(k+2) x 8 B (2.1 GB at k = 12)
                    build14_ms = gpu.extend_seed_table(args.seed_k + 2)
                    bytes14 = gpu.device_bytes()
                    dms14 = env.event_ms(call, reps=3)
                    counts14 = d_counts[:1_000_000].cpu().numpy().astype(np.uint32)
                build_seed_ms = gpu.extend_seed_table(args.derived_seed_depth)
                dms = env.event_ms(call, reps=3)
                derived_counts = d_counts[:1_000_000].cpu().numpy().astype(np.uint32)
                gpu.set_tuning(sweep_min_queries=-1)
                dms_tile = env.event_ms(call, reps=3)
                gpu.set_tuning(sweep_min_queries=0 if args.count_path == "auto" else (1 if args.count_path == "sweep" else -1))
                derived["seed_table"] = {"depth": args.derived_seed_depth, "build_ms": build_seed_ms, "count_ms": dms,
                                         "count_queries_per_s": n / dms * 1e3, "tile_kernel_ms": dms_tile,
                                         "tile_kernel_queries_per_s": n / dms_tile * 1e3}
                if args.derived_seed_depth > args.seed_k + 2:
                    derived["seed_table_modest"] = {"depth": args.seed_k + 2, "build_ms": build14_ms, "count_ms": dms14,
                                                    "count_queries_per_s": n / dms14 * 1e3, "device_bytes_with_it": bytes14,
                                                    "_counts": counts14}
                build_sa_ms = gpu.densify_suffix_array(1)
                wms = env.event_ms(walk)
                derived["suffix_array"] = {"sa_ratio": 1, "build_ms": build_sa_ms, "walk_ms": wms,
                                           "walk_hits_per_s": hits / wms * 1e3,
                                           "bit_exact_vs_plain_walk": bool(torch.equal(plain, d_lp))}
                derived["device_bytes_with_both"] = gpu.device_bytes()
                gpu.extend_seed_table(0)
                gpu.densify_suffix_array(0)
                call()  # d_counts again from the plain path (what the parity sample below checks)
                torch.cuda.synchronize()
                plain_counts = d_counts[:1_000_000].cpu().numpy().astype(np.uint32)
                derived["seed_table"]["bit_exact_vs_plain_path"] = bool(np.array_equal(derived_counts, plain_counts))
                modest_ok = True
                if "seed_table_modest" in derived:
                    modest_ok = bool(np.array_equal(derived["seed_table_modest"].pop("_counts"), plain_counts))
                    derived["seed_table_modest"]["bit_exact_vs_plain_path"] = modest_ok
                result["derived_structures"] = derived
                if not (derived["seed_table"]["bit_exact_vs_plain_path"] and modest_ok and
                        derived["suffix_array"]["bit_exact_vs_plain_walk"]):
                    raise SystemExit("PARITY FAILURE: derived structures change results")
                del d_lc, d_lr, d_lh, d_lp, plain
            except capi.AwfmGpuError as e:  # e.g. not enough free HBM for the depth asked for
                result["derived_structures"] = {"error": str(e)}
        del d_lq
        torch.cuda.empty_cache()

    # ---- variable-length batch (letters + offsets, lengths uniform in kmer-4 .. kmer+7): sweep with marker-bit payloads
    #      vs the tile kernel, sample checked against the oracle ----
    if world == 1 and "variable" not in args.skip:
        nv = min(n, 50_000_000)
        lo_len, hi_len = max(args.seed_k, L - 4), L + 7
        g = torch.Generator(device=dev).manual_seed(7)
        lengths = torch.randint(lo_len, hi_len + 1, (nv,), device=dev, generator=g, dtype=torch.int64)
        d_voff = torch.zeros(nv + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lengths, 0, out=d_voff[1:])
        del lengths
        total_letters = int(d_voff[-1].item())
        d_var = synth_device(env, total_letters, synth.QUERY_SEED + 5)
        d_vc = torch.zeros(nv, dtype=torch.int32, device=dev)
        vcall = lambda: gpu.count_device(d_var.data_ptr(), d_voff.data_ptr(), 0, nv, d_vc.data_ptr(), None, stream.cuda_stream)  # noqa: E731
        v_ms = env.event_ms(vcall, reps=3)
        gpu.set_tuning(sweep_profile=1)
        vcall()
        torch.cuda.synchronize()
        v_stages = gpu.sweep_stage_ms()
        gpu.set_tuning(sweep_profile=0)
        sweep_counts = d_vc.clone()
        gpu.set_tuning(sweep_min_queries=-1)
        v_tile_ms = env.event_ms(vcall, reps=3)
        gpu.set_tuning(sweep_min_queries=0 if args.count_path == "auto" else (1 if args.count_path == "sweep" else -1))
        var = {"workload": f"{nv} queries of {lo_len}..{hi_len} letters (uniform), {total_letters} letters + {nv + 1} offsets, device-resident",
               "ms": v_ms, "queries_per_s": nv / v_ms * 1e3, "path": "sweep (marker-bit payloads)" if v_stages else "tile kernel",
               "stages_ms": [round(x, 3) for x in v_stages], "tile_kernel_ms": v_tile_ms,
               "tile_kernel_queries_per_s": nv / v_tile_ms * 1e3,
               "sweep_equal_to_tile_kernel_all_queries": bool(torch.equal(sweep_counts, d_vc))}
        if arrays is not None:
            vs = min(nv, 500_000)
            h_off = d_voff[: vs + 1].cpu().numpy().astype(np.uint64)
            h_let = d_var[: int(h_off[-1])].cpu().numpy()
            o_counts, _, _ = harness.Oracle(arrays).count(h_let, h_off, threads=env.cores)
            var["parity_sample"] = {"queries": vs, "bit_exact_vs_oracle":
                                    bool(np.array_equal(o_counts, sweep_counts[:vs].cpu().numpy().astype(np.uint32)))}
            if not var["parity_sample"]["bit_exact_vs_oracle"]:
                raise SystemExit("PARITY FAILURE: variable-length counts differ from the oracle")
        if not var["sweep_equal_to_tile_kernel_all_queries"]:
            raise SystemExit("PARITY FAILURE: variable-length sweep and tile kernel disagree")
        result["variable_length"] = var
        del d_var, d_voff, d_vc, sweep_counts
        torch.cuda.empty_cache()

    # ---- parity on a sample + exact algorithmic bytes from the oracle (checker, not the product) ----
    sample = min(n, 1_000_000)
    h_counts_sample = d_counts[:sample].cpu().numpy().astype(np.uint32)
    h_letters_sample = d_letters[: sample * L].cpu().numpy()
    bytes_per_query = lines_per_query = None
    if arrays is not None:
        o_counts, _, work = harness.Oracle(arrays).count(h_letters_sample, fixed_len=L, threads=env.cores)
        parity = bool(np.array_equal(o_counts, h_counts_sample))
        bytes_per_query = work["countBytes"] / sample
        lines_per_query = (work["seeded"] + work["lfBlockReads"]) / sample
        result["parity_sample"] = {"queries": sample, "bit_exact_vs_oracle": parity,
                                   "lf_steps_per_query": work["lfSteps"] / sample,
                                   "block_reads_per_query": work["lfBlockReads"] / sample,
                                   "algorithmic_bytes_per_query": bytes_per_query}
        if not parity:
            raise SystemExit("PARITY FAILURE: CUDA counts differ from the oracle on the bench workload")

    # ---- e2e: the reference-facing drop-in call on host memory ----
    e2e = None
    threads = env.threads
    if "dropin" not in args.skip:
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
        ix = arrays.as_awfm_index()
        ip = C.addressof(ix)
        t0 = time.time()
        sl = KmerSearchList(lib, ne).fill(hl, fixed_len=L)  # awFmCreateKmerSearchList: one position list per query, as the reference
        list_s = time.time() - t0
        assert lib.awFmGpuPrepareIndex(ip) == abi.AwFmSuccess  # one-time upload, reported separately
        lib.awFmParallelSearchCount(ip, sl.ptr, threads)
        env.barrier()
        # Every timed call starts from the state awFmCreateKmerSearchList leaves (count == 0 in every entry,
        # src/AwFmParallelSearch.c:64); the engine does not rewrite an entry that already holds its count, so the lines of
        # queries without hits stay clean.  The same call with every count poisoned first (all 32-B entries written
        # back) is reported beside it.
        ent_counts = sl.entries()["count"]

        def reset_counts(v=0):
            ent_counts[:ne] = v

        times = wall_times(lambda: lib.awFmParallelSearchCount(ip, sl.ptr, threads), reps=max(args.e2e_steps, 5), warm=0, before=reset_counts)
        assert lib.awFmGpuLastCountStatus() == abi.AwFmSuccess
        times_poisoned = wall_times(lambda: lib.awFmParallelSearchCount(ip, sl.ptr, threads), reps=2, warm=0,
                                    before=lambda: reset_counts(0xFFFFFFFF))
        e2e_counts = sl.entries()["count"][: min(sample, ne)]
        if not np.array_equal(e2e_counts, h_counts_sample[: len(e2e_counts)]):
            raise SystemExit("PARITY FAILURE: drop-in counts differ from the device-resident path")
        t_step = env.max_over_ranks(sum(times) / len(times))
        e2e = {"value": world * ne / t_step, "unit": UNIT, "h2d_bytes_per_step": ne * L, "d2h_bytes_per_step": ne * 4,
               "call": "awFmParallelSearchCount(index, searchList, numThreads) drop-in, host AwFmKmerSearchList "
                       "(32-B entries pointing at query strings laid back to back in page-locked memory)",
               "host_threads": threads, "ms_per_step": 1e3 * t_step, "ms_each_call_this_rank": [round(1e3 * t, 2) for t in times],
               "search_list_setup_s": round(list_s, 2),
               "queries_per_gpu": ne,
               "list_state": "count == 0 in every entry before each timed call (a fresh list); entries whose count is already "
                             "right are not rewritten",
               "every_count_rewritten": {"what": "same call, every entry's count set to 0xFFFFFFFF before each call (untimed)",
                                         "ms_per_step": 1e3 * env.max_over_ranks(min(times_poisoned)),
                                         "queries_per_s": world * ne / env.max_over_ranks(min(times_poisoned))}}
        if ne != n:
            e2e["note"] = f"host memory bounds the list to {ne} of the {n} queries per rank"
        if world == 1:  # the unfriendly layout: pageable memory, every query string in its own 32-byte heap-like slot
            m = min(ne, 20_000_000)
            hp = np.zeros((m, 32), dtype=np.uint8)
            hp[:, :L] = hl[: m * L].reshape(m, L)
            sl2 = KmerSearchList(lib, m).fill(hp.reshape(-1)[: m * L], fixed_len=L)
            ent = sl2.entries()
            ent["kmerString"][:m] = np.uint64(hp.ctypes.data) + np.arange(m, dtype=np.uint64) * np.uint64(32)
            sl2._letters = hp
            ts = wall_times(lambda: lib.awFmParallelSearchCount(ip, sl2.ptr, threads), reps=2, warm=1)
            e2e["pageable_scattered_strings"] = {
                "layout": "pageable memory, one query string per 32-byte slot (not back to back)", "queries": m,
                "ms_per_step": 1e3 * min(ts), "queries_per_s": m / min(ts),
                "bit_exact": bool(np.array_equal(sl2.counts()[:sample], h_counts_sample[:m]))}
            sl2.close()
        sl.close()
        del h_letters
        lib.awFmGpuReleaseIndex(ip)

    # ---- e2e_packed: the additive packed-batch call, 2-bit queries in page-locked host memory ----
    e2e_packed = None
    h_bits = None
    if "packed" not in args.skip:
        d_bits = pack_bits_device(env, d_letters, n, L)
        h_bits = pinned_copy(env, d_bits)
        del d_bits
        pout = PinnedArray(n, np.uint32)
        group = GpuGroup(indexes=[gpu])
        pcall = lambda: group.count(h_bits.array, QUERY_2BIT, fixed_len=L, out=pout.array)  # noqa: E731
        pcall()
        env.barrier()
        times = wall_times(pcall, reps=max(args.e2e_steps, 5), warm=1)
        p_launches = group.stats()["launches"]
        same = bool(np.array_equal(pout.array[:sample], h_counts_sample))
        if not same:
            raise SystemExit("PARITY FAILURE: packed-batch counts differ from the device-resident path")
        t_step = env.max_over_ranks(sum(times) / len(times))
        e2e_packed = {"value": world * n / t_step, "unit": UNIT, "h2d_bytes_per_step": int(h_bits.array.nbytes),
                      "d2h_bytes_per_step": n * 4, "ms_per_step": 1e3 * t_step, "best_ms": 1e3 * min(times),
                      "call": "awfm_gpu_group_count / awFmGpuCountPacked: 2-bit packed 20-mers in page-locked host memory, u32 "
                              "counts into page-locked host memory, chunk-pipelined on 3 streams per GPU",
                      "gpu_launches_per_step": int(p_launches), "bit_exact_vs_device_path_sample": same,
                      "queries_per_gpu": n}
        # the same call on ASCII letters (20 B per query over the bus instead of 5)
        if world == 1:
            ha = pinned_copy(env, d_letters[: n * L])
            ts = wall_times(lambda: group.count(ha.array, 0, fixed_len=L, out=pout.array), reps=2, warm=1)
            e2e_packed["ascii_letters"] = {"ms_per_step": 1e3 * min(ts), "queries_per_s": n / min(ts),
                                           "h2d_bytes_per_step": n * L,
                                           "bit_exact": bool(np.array_equal(pout.array[:sample], h_counts_sample))}
            ha.close()
        pout.close()
        group.close()

    if e2e_packed is not None:
        link = leg_hostlink(env)
        result["host_link"] = link
        bytes_per_step = world * (e2e_packed["h2d_bytes_per_step"] + e2e_packed["d2h_bytes_per_step"])
        e2e_packed["host_link_GBps"] = bytes_per_step / (e2e_packed["ms_per_step"] * 1e-3) / 1e9
        e2e_packed["frac_of_host_link_both_directions"] = e2e_packed["host_link_GBps"] / link["both_GBps"]

    # ---- cpu baseline (rank 0, N=1): the unmodified reference on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and "cpu" not in args.skip:
        cores = env.cores
        ns = min(args.cpu_sample, n)
        hs = d_letters[: ns * L].cpu().numpy()
        if harness.have_reference():
            ref = harness.Reference()
            ix = arrays.as_awfm_index()
            rsl = KmerSearchList(ref.lib, ns).fill(hs, fixed_len=L)
            ts = wall_times(lambda: ref.lib.awFmParallelSearchCount(C.addressof(ix), rsl.ptr, cores), reps=3, warm=1)
            r_counts = rsl.counts()
            rsl.close()
            ok = bool(np.array_equal(r_counts[:sample], h_counts_sample[: len(r_counts[:sample])]))
            cpu = {"value": ns / min(ts), "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"first {ns} of the {n} queries, 1 warm-up + best of 3 passes of the reference's "
                             f"awFmParallelSearchCount (oracle/_ref), numThreads={cores}",
                   "bit_exact_vs_cuda": ok}
            if "locate" in result and "reference" in result["locate"]:
                cpu["locate"] = result["locate"]["reference"]
        else:
            t1 = time.perf_counter()
            harness.Oracle(arrays).count(hs[: 1_000_000 * L], fixed_len=L, threads=cores)
            dt = time.perf_counter() - t1
            cpu = {"value": 1_000_000 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first 1000000 queries, scalar C oracle with OpenMP over queries"}

    device_bytes = gpu.device_bytes()
    # ---- N>1: one process, all GPUs (rank 0; the other ranks' GPUs are idle meanwhile) ----
    if world > 1 and "fanout" not in args.skip and h_bits is not None:
        env.cpu_barrier()
        if rank == 0:
            full = d_counts.cpu().numpy().astype(np.uint32)
            result["single_process_fanout"] = leg_fanout(env, arrays, h_bits.array, full)
        env.cpu_barrier()
    if h_bits is not None:
        h_bits.close()
    gpu.close()
    del d_letters, d_counts, bufs
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs ----
    def guarded(fn, *a):
        """a secondary leg that breaks (out of memory on a small box, ...) is reported, it does not take the headline
        with it; a parity failure (SystemExit) is never swallowed"""
        if world > 1:
            return fn(*a)  # ranks must fail together
        try:
            return fn(*a)
        except Exception as e:  # noqa: BLE001
            import traceback
            torch.cuda.empty_cache()
            return {"error": repr(e), "traceback": traceback.format_exc()[-1500:]}

    if world == 1:
        if "cfg1" not in args.skip:
            result["cfg1"] = guarded(leg_cfg1, env)
        if "cfg3" not in args.skip and nl > 0:
            result["cfg3"] = {"ratio_%d" % args.sa_ratio: "see `locate`"}
            for ratio in (1, 16):
                if ratio != args.sa_ratio:
                    result["cfg3"]["ratio_%d" % ratio] = guarded(leg_cfg3_ratio, env, ratio)
        if "cfg4" not in args.skip:
            result["cfg4"] = guarded(leg_cfg4, env)
    if "cfg5" not in args.skip:
        result["cfg5"] = guarded(leg_cfg5, env)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline ----
    peak, peak_src = measured_peak()
    traffic = ncu_traffic_record()
    roofline = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "kernel_ms": kernel_avg,
                "random_access": random_access}
    if stage_ms:
        # compulsory DRAM traffic of one sweep call: what any implementation of this algorithm has
        # to move — query bytes in, counts out, per pass the live records in and out (16 B each) and the index once
        # while the live set still covers it (else two 32-B sectors per record), the seed entries once.  The
        # re-ordering that makes the passes streamable (pack output, radix sort, the first pass's re-read of the sorted
        # pairs, the output clear) is overhead, not compulsory.
        index_bytes = (args.bp + 1) / 2.0  # sectors: 0.5 B per BWT position
        seed_bytes = 16.0 * 4 ** args.seed_k
        comp = n * L + 4.0 * n + min(seed_bytes, 16.0 * n)
        per_pass = []
        for p, alive in enumerate(live):  # pass p (0-based) takes live[p] records in and leaves live[p+1]
            nxt = live[p + 1] if p + 1 < len(live) else 0
            b = (0 if p == 0 else 16.0 * alive) + min(index_bytes, 64.0 * alive) + 16.0 * nxt
            per_pass.append(b)
            comp += b
        achieved = comp / (kernel_avg * 1e-3) / 1e9
        names = ["clear+pack", "radix sort", "sweepStep<first>"] + \
                [f"sweepStep pass {i + 2}" for i in range(len(stage_ms) - 4)] + ["sweepIrregular"]
        roofline.update({
            "kernel": "sweep pipeline, one count call over the rank's whole batch: pack -> radix sort on the seed index -> "
                      "one sweepStep pass per LF step (csrc/awfm_sweep.cuh)",
            "achieved": achieved, "frac": achieved / peak,
            "model": "compulsory DRAM bytes of the sweep (letters in + counts out + seed entries + per pass: live records "
                     "in/out at 16 B + the index once while the live set covers it) / device time of the call; "
                     "re-ordering traffic (pack output, radix sort, re-read of sorted pairs, output clear) is not counted",
            "compulsory_bytes_per_launch": comp, "compulsory_bytes_per_pass": per_pass, "live_records_per_pass": live,
            "irregular_queries": irregular,
            "stages_ms": {k: round(v, 3) for k, v in zip(names, stage_ms)},
            "traffic": traffic.get("sweep_dram_bytes_per_call"), "traffic_source": traffic.get("sweep_file"),
        })
        if traffic.get("sweep_dram_bytes_per_call"):
            roofline["dram_GBps"] = traffic["sweep_dram_bytes_per_call"] / (kernel_avg * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_GBps"] / peak
        if bytes_per_query:
            roofline["per_query_gather_model"] = {
                "algorithmic_bytes_per_launch": bytes_per_query * n,
                "GBps": bytes_per_query * n / (kernel_avg * 1e-3) / 1e9,
                "note": "SURVEY 8d's per-query model charges every rank its own block read; the sweep shares line fetches "
                        "between sorted queries, so this figure is not a bound for it (reported for continuity with round 1)"}
        if tile_ms and bytes_per_query:
            tk = {"kernel": "countKernelV1 (one launch, one random 128-B line per rank)", "kernel_ms": tile_ms,
                  "queries_per_s": n / tile_ms * 1e3, "achieved": bytes_per_query * n / (tile_ms * 1e-3) / 1e9,
                  "frac": bytes_per_query * n / (tile_ms * 1e-3) / 1e9 / peak,
                  "traffic": traffic.get("dram_bytes_per_launch"), "traffic_source": traffic.get("tile_file")}
            if lines_per_query and random_access and "lines_per_s" in random_access:
                tk["line_misses_per_query"] = lines_per_query
                tk["lines_per_s"] = lines_per_query * n / (tile_ms * 1e-3)
                tk["frac_of_random_access_roofline"] = tk["lines_per_s"] / random_access["lines_per_s"]
            roofline["tile_kernel"] = tk
    else:
        achieved = (bytes_per_query * n / (kernel_avg * 1e-3) / 1e9) if bytes_per_query else None
        roofline.update({"kernel": "countKernelV1 (one launch over the rank's whole batch)", "achieved": achieved,
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic.get("dram_bytes_per_launch"), "traffic_source": traffic.get("tile_file")})
        if lines_per_query and random_access and "lines_per_s" in random_access:
            roofline["lines_per_s"] = lines_per_query * n / (kernel_avg * 1e-3)
            roofline["frac_of_random_access_roofline"] = roofline["lines_per_s"] / random_access["lines_per_s"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": e2e, "e2e_packed": e2e_packed, "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "index": {"device_bytes": device_bytes, "build_s": round(build_s, 2), "build_gpu_ms": round(build_ms, 1),
                  "suffixes_tied_after_radix_pass": tie_suffixes, "prefix_doubling_rounds_on_device": tie_rounds},
        "gather": gather_mode,
    }
    if e2e is None and e2e_packed is not None:  # --skip dropin: the packed call is the only end-to-end number
        line["e2e"] = e2e_packed
    line.update(result)
    print(json.dumps(line))
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
