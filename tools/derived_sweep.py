"""Derived structures at BASELINE sizes: what a deeper seed table (count, cfg 2) and denser SA samples (locate, cfg 3)
buy on one B200, what they cost in HBM and build time, and that results stay bit-exact against the oracle on a sample.
Appends JSON lines to gpurun_out/derived_sweep.jsonl.  Measurement tool, not product."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from avxwindowfmindex_b200 import DeviceBuiltIndex, abi, capi, synth  # noqa: E402
from oracle import harness  # noqa: E402


def timed(fn, reps=4):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=3_100_000_000)
    ap.add_argument("--amino", action="store_true")
    ap.add_argument("--seed-k", type=int, default=12)
    ap.add_argument("--ratio", type=int, default=8)
    ap.add_argument("--count-queries", type=int, default=100_000_000)
    ap.add_argument("--count-kmer", type=int, default=20)
    ap.add_argument("--locate-queries", type=int, default=10_000_000)
    ap.add_argument("--locate-kmer", type=int, default=16)
    ap.add_argument("--depths", default="0,13,14,15,16")
    ap.add_argument("--ratios", default="0,4,2,1")
    ap.add_argument("--sample", type=int, default=500_000)
    ap.add_argument("--count-variants", default="1")
    ap.add_argument("--count-lpqs", default="2")
    a = ap.parse_args()
    lib = capi.load()
    dev = torch.device("cuda:0")
    alphabet = abi.AwFmAlphabetAmino if a.amino else abi.AwFmAlphabetDna
    d_text = torch.empty(a.bp, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_text.data_ptr(), a.bp, synth.TEXT_SEED + 2, 0, int(a.amino)))
    built = DeviceBuiltIndex.from_device_text(d_text.data_ptr(), a.bp, alphabet, a.seed_k, a.ratio)
    del d_text
    gpu = built.gpu_index()
    arrays = built.to_host()
    built.close()
    torch.cuda.empty_cache()
    oracle = harness.Oracle(arrays)
    stream = torch.cuda.current_stream().cuda_stream
    out = {"text": a.bp, "alphabet": "amino" if a.amino else "dna", "seed_k": a.seed_k, "sa_ratio": a.ratio,
           "base_device_bytes": gpu.device_bytes(), "count": [], "locate": []}

    # ---- count vs seed-table depth ----
    n, L = a.count_queries, a.count_kmer
    d_q = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_q.data_ptr(), n * L, synth.QUERY_SEED + 2, 0, int(a.amino)))
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ranges = torch.zeros((n, 2), dtype=torch.int64, device=dev)
    ns = min(a.sample, n)
    hs = d_q[: ns * L].cpu().numpy()
    o_counts, o_ranges, work = oracle.count(hs, fixed_len=L, threads=os.cpu_count())
    for depth in [int(x) for x in a.depths.split(",")]:
        build_ms = gpu.extend_seed_table(depth)
        for variant in [int(x) for x in a.count_variants.split(",")]:
            for lpq in [int(x) for x in a.count_lpqs.split(",")]:
                gpu.set_tuning(count_variant=variant, count_lpq=lpq)
                ms = timed(lambda: gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), None, stream))
                gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
                torch.cuda.synchronize()
                ok = (np.array_equal(d_counts[:ns].cpu().numpy().astype(np.uint32), o_counts)
                      and np.array_equal(d_ranges[:ns].cpu().numpy().astype(np.uint64), o_ranges))
                row = {"depth": depth or a.seed_k, "derived": bool(depth), "count_variant": variant, "count_lpq": lpq,
                       "build_ms": round(build_ms, 1), "device_bytes": gpu.device_bytes(), "ms": round(ms, 3),
                       "Gq_per_s": round(n / ms / 1e6, 3), "ranges_and_counts_bit_exact_on_sample": bool(ok)}
                out["count"].append(row)
                print(json.dumps(row), flush=True)
    gpu.extend_seed_table(0)
    del d_q, d_counts, d_ranges
    torch.cuda.empty_cache()

    # ---- locate vs SA sampling ratio ----
    n, L = a.locate_queries, a.locate_kmer
    d_q = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
    capi.check(lib.awfm_gpu_synth_letters(0, d_q.data_ptr(), n * L, synth.QUERY_SEED + 3, 0, int(a.amino)))
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ranges = torch.zeros((n, 2), dtype=torch.int64, device=dev)
    d_hit = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    gpu.count_device(d_q.data_ptr(), None, L, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), stream)
    total = int(d_hit[-1].item())
    d_pos = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)
    ns = min(a.sample, n)
    hs = d_q[: ns * L].cpu().numpy()
    o_hit, o_pos, _ = oracle.locate(hs, fixed_len=L, threads=os.cpu_count())
    nh = int(o_hit[-1])
    for ratio in [int(x) for x in a.ratios.split(",")]:
        build_ms = gpu.densify_suffix_array(ratio)
        ms = timed(lambda: gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, 0, total, d_pos.data_ptr(), stream))
        ok = np.array_equal(d_pos[:nh].cpu().numpy().astype(np.uint64), o_pos)
        row = {"sa_ratio": ratio or a.ratio, "derived": bool(ratio), "build_ms": round(build_ms, 1),
               "device_bytes": gpu.device_bytes(), "hits": total, "backtrace_ms": round(ms, 3),
               "Ghits_per_s": round(total / ms / 1e6, 3), "positions_bit_exact_on_sample": bool(ok)}
        out["locate"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "derived_sweep.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")
    gpu.close()


if __name__ == "__main__":
    main()
