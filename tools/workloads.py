"""Secondary BASELINE.json configurations on one B200 (run under gpurun): cfg 3 (locate, 10 M random 16-mers on the
3.1 Gbp index, SA ratios 1/8/16) and cfg 4 (1 G-residue amino index, seed k=5, count + locate of 50 M random 8-mers).
For each: device-resident kernel throughput (CUDA events), the drop-in call on a host AwFmKmerSearchList, the
unmodified reference on a bounded sample with all host cores, and bit-exact parity of the sample.  One JSON line per
configuration, appended to gpurun_out/workloads.jsonl.  Not the bench contract (bench.py is); evidence for DESIGN.md."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from avxwindowfmindex_b200 import DeviceBuiltIndex, KmerSearchList, abi, capi, synth  # noqa: E402
from oracle import harness  # noqa: E402


def run(name, amino, bp, seed_k, ratio, nq, L, sample_q, e2e_q, reps=3):
    lib = capi.load()
    dev = torch.device("cuda:0")
    alphabet = abi.AwFmAlphabetAmino if amino else abi.AwFmAlphabetDna
    t0 = time.time()
    d_text = torch.empty(bp, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), bp, synth.TEXT_SEED + 3, 0, int(amino)))
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), bp, alphabet, seed_k, ratio)
    del d_text
    gpu = built.gpu_index()
    arrays = built.to_host()
    out = {"config": name, "alphabet": "amino" if amino else "dna", "text": bp, "seed_k": seed_k, "sa_ratio": ratio,
           "queries": nq, "kmer": L, "index_build_gpu_ms": built.build_ms, "tie_suffixes": built.tie_suffixes,
           "device_bytes": gpu.device_bytes(), "setup_s": round(time.time() - t0, 1)}
    built.close()
    torch.cuda.empty_cache()
    stream = torch.cuda.current_stream().cuda_stream
    d_q = torch.empty(nq * L + 64, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_q.data_ptr(), nq * L, synth.QUERY_SEED + 3, 0, int(amino)))
    d_counts = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_ranges = torch.zeros((nq, 2), dtype=torch.int64, device=dev)
    d_hit = torch.zeros(nq + 1, dtype=torch.int64, device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    # ---- count ----
    ms = timed(lambda: gpu.count_device(d_q.data_ptr(), None, L, nq, d_counts.data_ptr(), None, stream))
    out["count_ms"] = ms
    out["count_queries_per_s"] = nq / ms * 1e3
    # ---- locate: ranges -> scan -> backtrace ----
    gpu.count_device(d_q.data_ptr(), None, L, nq, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    gpu.scan_ranges_device(d_ranges.data_ptr(), nq, d_hit.data_ptr(), stream)
    total = int(d_hit[-1].item())
    d_pos = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)

    def locate_all():
        gpu.count_device(d_q.data_ptr(), None, L, nq, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
        gpu.scan_ranges_device(d_ranges.data_ptr(), nq, d_hit.data_ptr(), stream)
        gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), nq, 0, total, d_pos.data_ptr(), stream)

    ms_all = timed(locate_all)
    ms_bt = timed(lambda: gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), nq, 0, total, d_pos.data_ptr(), stream))
    out.update({"hits": total, "locate_total_ms": ms_all, "backtrace_ms": ms_bt,
                "locate_queries_per_s": nq / ms_all * 1e3, "located_hits_per_s": total / ms_all * 1e3,
                "backtrace_hits_per_s": total / ms_bt * 1e3})
    # ---- parity + exact algorithmic bytes on a sample (oracle = checker) ----
    hs = d_q[: sample_q * L].cpu().numpy()
    oracle = harness.Oracle(arrays)
    o_counts, _, wc = oracle.count(hs, fixed_len=L, threads=os.cpu_count())
    o_hit, o_pos, wl = oracle.locate(hs, fixed_len=L, threads=os.cpu_count())
    nh = int(o_hit[-1])
    ok = (np.array_equal(d_counts[:sample_q].cpu().numpy().astype(np.uint32), o_counts)
          and np.array_equal(d_hit[: sample_q + 1].cpu().numpy().astype(np.uint64), o_hit)
          and np.array_equal(d_pos[:nh].cpu().numpy().astype(np.uint64), o_pos))
    out["parity_sample"] = {"queries": sample_q, "hits": nh, "bit_exact_vs_oracle": bool(ok)}
    out["count_alg_bytes_per_query"] = wc["countBytes"] / sample_q
    out["count_alg_GBps"] = wc["countBytes"] / sample_q * nq / ms / 1e6
    if nh:
        out["backtrace_steps_per_hit"] = wl["backtraceSteps"] / nh
        out["locate_alg_bytes_per_hit"] = wl["locateBytes"] / nh
        out["backtrace_alg_GBps"] = wl["locateBytes"] / nh * total / ms_bt / 1e6
    # ---- drop-in on host memory ----
    ix = arrays.as_awfm_index()
    ip = C.addressof(ix)
    ne = min(e2e_q, nq)
    h_letters = torch.empty(ne * L, dtype=torch.uint8).pin_memory()
    h_letters.copy_(d_q[: ne * L])
    sl = KmerSearchList(lib, ne).fill(h_letters.numpy(), fixed_len=L)
    threads = os.cpu_count()
    assert lib.awFmGpuPrepareIndex(ip) == abi.AwFmSuccess
    lib.awFmParallelSearchCount(ip, sl.ptr, threads)
    best_c = best_l = 1e30
    for _ in range(reps):
        t1 = time.perf_counter()
        lib.awFmParallelSearchCount(ip, sl.ptr, threads)
        best_c = min(best_c, time.perf_counter() - t1)
    assert lib.awFmParallelSearchLocate(ip, sl.ptr, threads) == abi.AwFmSuccess
    for _ in range(reps):
        t1 = time.perf_counter()
        rc = lib.awFmParallelSearchLocate(ip, sl.ptr, threads)
        best_l = min(best_l, time.perf_counter() - t1)
    e = sl.entries()
    e2e_hits = int(e["count"][:ne].astype(np.uint64).sum())
    ok_e2e = rc == abi.AwFmSuccess and np.array_equal(e["count"][:sample_q], o_counts[: min(sample_q, ne)])
    first_pos = np.concatenate([p for p in sl.positions()[:1000]] or [np.zeros(0, np.uint64)]) if ne >= 1000 else None
    if first_pos is not None:
        ok_e2e = ok_e2e and np.array_equal(first_pos, o_pos[: int(o_hit[1000])])
    out["e2e_dropin"] = {"queries": ne, "host_threads": threads, "count_queries_per_s": ne / best_c,
                         "locate_queries_per_s": ne / best_l, "located_hits_per_s": e2e_hits / best_l,
                         "bit_exact": bool(ok_e2e)}
    # ---- the unmodified reference on the host cores, bounded sample ----
    if harness.have_reference():
        ref = harness.Reference()
        ns = min(sample_q, ne)
        rsl = KmerSearchList(ref.lib, ns).fill(hs[: ns * L], fixed_len=L)
        ref.lib.awFmParallelSearchCount(ip, rsl.ptr, threads)
        bc = bl = 1e30
        for _ in range(2):
            t1 = time.perf_counter()
            ref.lib.awFmParallelSearchCount(ip, rsl.ptr, threads)
            bc = min(bc, time.perf_counter() - t1)
        for _ in range(2):
            t1 = time.perf_counter()
            ref.lib.awFmParallelSearchLocate(ip, rsl.ptr, threads)
            bl = min(bl, time.perf_counter() - t1)
        r_counts = rsl.counts()
        r_pos = np.concatenate(rsl.positions()[:1000]) if ns >= 1000 else None
        same = np.array_equal(r_counts, o_counts[:ns]) and (r_pos is None or np.array_equal(r_pos, o_pos[: int(o_hit[1000])]))
        out["cpu_reference"] = {"queries": ns, "cores": threads, "count_queries_per_s": ns / bc,
                                "locate_queries_per_s": ns / bl, "located_hits_per_s": int(o_hit[ns]) / bl,
                                "bit_exact_vs_oracle": bool(same)}
        rsl.close()
    sl.close()
    lib.awFmGpuReleaseIndex(ip)
    gpu.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "workloads.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")
    print(json.dumps(out), flush=True)
    del d_q, d_counts, d_ranges, d_hit, d_pos
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="cfg3,cfg4")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink texts/queries for a quick check")
    a = ap.parse_args()
    s = a.scale
    if "cfg3" in a.which:
        for ratio in (1, 8, 16):
            run(f"cfg3 locate 16-mers ratio {ratio}", False, int(3_100_000_000 * s), 12, ratio, int(10_000_000 * s), 16,
                int(1_000_000 * min(1, s * 10)), int(10_000_000 * s))
    if "cfg4" in a.which:
        run("cfg4 amino 8-mers", True, int(1_000_000_000 * s), 5, 8, int(50_000_000 * s), 8,
            int(1_000_000 * min(1, s * 10)), int(50_000_000 * s))


if __name__ == "__main__":
    main()
